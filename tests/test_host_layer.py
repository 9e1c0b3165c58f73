"""CPU tests of the C++ host layer (include/fwhost.h): the reference's parser / vwmap / cmdline / cache known answers
through the PRODUCT code (tests/test_oracle_goldens.py runs the same vectors through the oracle), plus agreement of
the two independent implementations on a real-shape file."""
import json
import os

import numpy as np
import pytest

from fwumious_wabbit_b200 import Optimizer, host, synth
from oracle import fw_oracle as fo

M31, NOF, ONE = 0x7FFFFFFF, 0x80000000, 1065353216
REF_TRAIN = "/root/reference/examples/basic/datasets/train.vw"


def nd(a, b):
    return (a << 16) + b


def bits(x):
    return int(np.float32(x).view(np.uint32))


def test_vwmap_csv():  # vwmap.rs:159-222
    vw = host.VwNamespaceMap.new("\nA,featureA\nB,featureB\nC,featureC\n")
    assert vw.num_namespaces == 3
    assert [e["namespace_vwname"] for e in vw.source["entries"]] == ["A", "B", "C"]
    assert [e["namespace_index"] for e in vw.source["entries"]] == [0, 1, 2]
    vw = host.VwNamespaceMap.new("A,featureA\nB,featureB,f32\n_namespace_skip_prefix,2\n")
    assert vw.source["namespace_skip_prefix"] == 2 and vw.ns_is_f32() == [0, 1]
    with pytest.raises(ValueError, match='Unknown type used for the feature in vw_namespace_map.csv: "blah"'):
        host.VwNamespaceMap.new("A,featureA,blah\n")


def test_parser_known_answers():  # parser.rs:475-857 (same vectors as the oracle test)
    p = host.VowpalParser(host.VwNamespaceMap.new("A,featureA\nB,featureB\nC,featureC\n"))
    a = 2988156968 & M31
    for line in ("1 |A a\n", "1 |A a \n", "1  |A a\n", "1 |A  a\n", "1 |A:1.0 a\n"):
        assert p.next_vowpal(line).tolist() == [6, 1, ONE, a, NOF, NOF]
    assert p.next_vowpal("-1 |B b\n").tolist() == [6, 0, ONE, NOF, 2422381320 & M31, NOF]
    assert p.next_vowpal("1 |A a b\n").tolist() == [10, 1, ONE, nd(6, 10) | NOF, NOF, NOF, a, ONE, 3529656005 & M31, ONE]
    assert p.next_vowpal("1 |A:3 a:2.0\n").tolist() == [8, 1, ONE, nd(6, 8) | NOF, NOF, NOF, a, bits(6.0)]
    assert p.next_vowpal("1 |A a b:2.0 c:3.0\n").tolist() == [12, 1, ONE, nd(6, 12) | NOF, NOF, NOF, a, bits(1.0), 3529656005 & M31,
                                                             bits(2.0), 906509 & M31, bits(3.0)]
    assert p.next_vowpal("|A a\n").tolist() == [6, 0xFF, ONE, a, NOF, NOF]
    assert p.next_vowpal("1 0.1 |A a\n").tolist() == [6, 1, bits(0.1), a, NOF, NOF]
    assert p.next_vowpal("").tolist() == []
    for bad, msg in [("1 |UNDECLARED_NAMESPACE a\n", "Feature name was not predeclared in vw_namespace_map.csv: UNDECLARED_NAMESPACE"),
                     ("1 |A:not_a_parsable_number a\n", "Failed parsing namespace weight: not_a_parsable_number"),
                     ("1 |A a:2x0\n", "Failed parsing feature weight: 2x0"), ("$1", "Cannot parse an example"),
                     ("1 -0.1 |A a\n", "Example importance cannot be negative: -0.1! "),
                     ("1 fdsa |A a\n", "Failed parsing example importance: fdsa"), ("hogwild_load ", "Cannot parse an example")]:
        with pytest.raises(ValueError, match=msg.replace("(", r"\(")):
            p.next_vowpal(bad)
    with pytest.raises(host.FlushCommand):
        p.next_vowpal("flush")
    with pytest.raises(host.HogwildLoadCommand):
        p.next_vowpal("hogwild_load   /path/to/filename  ")
    pf = host.VowpalParser(host.VwNamespaceMap.new("A,featureA\nB,featureB,f32\nC,featureC\n_namespace_skip_prefix,1\n"))
    assert pf.next_vowpal("-1 |B B3\n").tolist() == [8, 0, ONE, NOF, nd(6, 8) | NOF, NOF, 1416737454 & M31, bits(3.0)]
    with pytest.raises(ValueError, match="Namespaces that are f32 can not have weight attached"):
        pf.next_vowpal("-1 |B:3 B3\n")


def test_cmdline_to_model_instance():  # model_instance.rs:567-696 and the flag semantics of :296-495
    vw = host.VwNamespaceMap.new("A,featureA\nB,featureB\nC,featureC\n")
    mi = host.new_model_instance_from_cmdline(["--keep", "A", "--keep", "B", "--interactions", "AB", "--linear", "featureC:3",
                                               "--ffm_k", "4", "--ffm_field", "A", "--ffm_field", "BC", "--ffm_bit_precision", "20",
                                               "-b", "18", "-l", "0.1", "--power_t", "0.4", "--adaptive", "--sgd"], vw)
    assert mi.feature_combo_descs == [([0], 1.0), ([1], 1.0), ([0, 1], 1.0), ([2], 3.0)]
    assert mi.ffm_fields == [[0], [1, 2]] and mi.ffm_k == 4 and mi.ffm_bit_precision == 20 and mi.bit_precision == 18
    assert mi.optimizer == Optimizer.AdagradLUT                      # --adaptive + fastmath (model_instance.rs:481-492)
    assert mi.ffm_learning_rate == np.float32(0.1) and mi.ffm_power_t == np.float32(0.4)  # default to the LR values (:418-428)
    assert mi.init_acc_gradient == 1.0 and mi.ffm_init_acc_gradient == 1.0
    mi = host.new_model_instance_from_cmdline(["--keep", "A", "--vwcompat", "--hash", "all", "--sgd", "--adaptive", "--noconstant"], vw)
    assert mi.optimizer == Optimizer.AdagradFlex and mi.init_acc_gradient == 0.0 and not mi.add_constant_feature
    with pytest.raises(ValueError, match="--vwcompat requires use of --sgd"):
        host.new_model_instance_from_cmdline(["--keep", "A", "--vwcompat", "--hash", "all"], vw)
    with pytest.raises(ValueError, match="Maximum ffm_k is: 128, passed: 200"):
        host.new_model_instance_from_cmdline(["--ffm_k", "200"], vw)
    with pytest.raises(ValueError, match="--l2 can only be 0.0"):
        host.new_model_instance_from_cmdline(["--l2", "0.5"], vw)
    with pytest.raises(ValueError, match="wasn't expected"):
        host.new_model_instance_from_cmdline(["--no_such_flag"], vw)
    # the JSON is the reference's: field order of model_instance.rs:47-97, enums as strings, pretty-printed
    js = host.model_instance_json_from_cmdline(["--keep", "A", "-l", "0.025"], vw)
    j = json.loads(js)
    assert list(j)[:6] == ["learning_rate", "minimum_learning_rate", "power_t", "bit_precision", "add_constant_feature", "feature_combo_descs"]
    assert j["optimizer"] == "SGD" and j["nn_config"] == {"layers": [], "topology": "one"} and j["transform_namespaces"] == {"v": []}
    assert '"learning_rate": 0.025,' in js and '"power_t": 0.5,' in js and js.startswith("{\n  \"learning_rate\"")
    assert j["feature_combo_descs"][0]["namespace_descriptors"][0] == {"namespace_index": 0, "namespace_type": "Primitive", "namespace_format": "Categorical"}


def test_cache_roundtrip_and_header(tmp_path):  # cache.rs:12-26, 133-182
    w = synth.workload("c2")
    vw = host.VwNamespaceMap.new("".join(f"{c},feature{c}\n" for c in w.ns_names))
    recs = w.records(1000).reshape(-1)
    path = str(tmp_path / "train.vw.fwcache")
    host.cache_write(path, vw, recs)
    raw = open(path, "rb").read()
    assert raw[:4] == b"FWCA" and int.from_bytes(raw[4:8], "little") == 11
    blob_len = int.from_bytes(raw[8:16], "little")
    assert json.loads(raw[16:16 + blob_len]) == vw.source and len(raw) == 16 + blob_len + recs.nbytes
    r, off, js = host.cache_read(path, vw)
    assert np.array_equal(r, recs) and off.tolist() == list(range(0, 11001, 11)) and json.loads(js) == vw.source
    other = host.VwNamespaceMap.new("A,featureA\n")
    with pytest.raises(IOError, match="vw_namespace_map.csv and the one from cache file differ"):
        host.cache_read(path, other)
    open(path, "r+b").write(b"XXXX")
    with pytest.raises(IOError, match="Cache header does not begin with magic bytes"):
        host.cache_read(path, vw)


def test_cache_of_gz_input_is_an_lz4_frame(tmp_path):  # cache.rs:68-71, 89-125
    """The cache of a `*.gz` input is the same byte stream inside an LZ4 frame.  Our codec (csrc/host/lz4frame.hpp, written from
    the format description) is checked against an independent implementation of the frame format (Arrow's): Arrow reads what
    we write, we read what Arrow writes, and damaged frames are rejected."""
    pa = pytest.importorskip("pyarrow")
    w = synth.workload("c2")
    vw = host.VwNamespaceMap.new("".join(f"{c},feature{c}\n" for c in w.ns_names))
    recs = w.records(30000).reshape(-1)
    plain, packed = str(tmp_path / "train.vw.fwcache"), str(tmp_path / "train.vw.gz.fwcache")
    host.cache_write(plain, vw, recs)
    host.cache_write(packed, vw, recs)
    image, z = open(plain, "rb").read(), open(packed, "rb").read()
    assert z[:4] == bytes([0x04, 0x22, 0x4D, 0x18]) and len(z) < 0.8 * len(image)   # a frame, and it does compress
    assert pa.Codec("lz4").decompress(z, decompressed_size=len(image)).to_pybytes() == image
    r, off, js = host.cache_read(packed, vw)
    assert np.array_equal(r, recs) and off.tolist() == list(range(0, 330001, 11)) and json.loads(js) == vw.source
    # a frame from the other implementation (its own block size / flags / checksums)
    foreign = str(tmp_path / "other.vw.gz.fwcache")
    open(foreign, "wb").write(pa.Codec("lz4").compress(image).to_pybytes())
    r2, off2, _ = host.cache_read(foreign, vw)
    assert np.array_equal(r2, recs) and np.array_equal(off2, off)
    # more than one 4 MiB block, and Arrow's multi-block (linked) frames of the same image
    big = w.records(250000).reshape(-1)   # 11 MB
    bigp, bigz = str(tmp_path / "big.vw.fwcache"), str(tmp_path / "big.vw.gz.fwcache")
    host.cache_write(bigp, vw, big)
    host.cache_write(bigz, vw, big)
    big_image = open(bigp, "rb").read()
    assert pa.Codec("lz4").decompress(open(bigz, "rb").read(), decompressed_size=len(big_image)).to_pybytes() == big_image
    assert np.array_equal(host.cache_read(bigz, vw)[0], big)
    open(bigz, "wb").write(pa.Codec("lz4").compress(big_image).to_pybytes())
    assert np.array_equal(host.cache_read(bigz, vw)[0], big)
    # edge cases: empty cache body, incompressible payload, damaged frames
    empty = str(tmp_path / "empty.vw.gz.fwcache")
    host.cache_write(empty, vw, np.zeros(0, np.uint32))
    assert host.cache_read(empty, vw)[0].size == 0
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 1 << 32, size=22 * 1000, dtype=np.uint64).astype(np.uint32)
    noise[::22] = 22  # valid record lengths, random payload
    hard = str(tmp_path / "noise.vw.gz.fwcache")
    host.cache_write(hard, vw, noise)
    assert np.array_equal(host.cache_read(hard, vw)[0], noise)
    bad = bytearray(z); bad[5] ^= 0x10   # descriptor byte: header checksum no longer matches
    open(packed, "wb").write(bytes(bad))
    with pytest.raises(IOError, match="lz4"):
        host.cache_read(packed, vw)
    open(packed, "wb").write(z[: len(z) // 2])
    with pytest.raises(IOError, match="lz4"):
        host.cache_read(packed, vw)


def test_batch_parser_equals_line_parser_and_oracle():
    """Multi-threaded whole-buffer parse == line-by-line parse == the oracle's independent parser, on synthetic lines and,
    when the reference tree is present, on its real-shape example file (58 namespaces, weights, multi-valued namespaces)."""
    w = synth.workload("c2")
    vw = host.VwNamespaceMap.new("".join(f"{c},feature{c}\n" for c in w.ns_names))
    p = host.VowpalParser(vw)
    text = "".join(w.line(i) + "\n" for i in range(5000))
    recs, off = p.parse_text(text, threads=4)
    assert np.array_equal(recs.reshape(-1, 11), w.records(5000))
    if os.path.exists(REF_TRAIN):
        csv = open(os.path.join(os.path.dirname(REF_TRAIN), "vw_namespace_map.csv")).read()
        vw = host.VwNamespaceMap.new(csv)
        p = host.VowpalParser(vw)
        text = open(REF_TRAIN).read()
        recs, off = p.parse_text(text, threads=3)
        names = [None] * vw.num_namespaces
        for e in vw.source["entries"]:
            names[e["namespace_index"]] = e["namespace_vwname"]
        op = fo.Parser([n or "\x00unused" for n in names], ns_is_f32=vw.ns_is_f32(), namespace_skip_prefix=vw.source["namespace_skip_prefix"])
        lines = text.splitlines(keepends=True)
        assert len(off) - 1 == len(lines) == 100
        for i, line in enumerate(lines):
            one = p.next_vowpal(line)
            assert np.array_equal(one, recs[off[i]:off[i + 1]])
            assert np.array_equal(one, op.parse(line))


def test_regressor_file_layout(tmp_path):  # persistence.rs:55-97 (no GPU needed: raw writer/reader)
    import ctypes as C

    L = host._L()
    vw = host.VwNamespaceMap.new("A,featureA\nB,featureB\n")
    mi_json = host.model_instance_json_from_cmdline(["--keep", "A", "--ffm_k", "1", "--ffm_field", "A", "--ffm_field", "B", "--adaptive"], vw)
    lr = np.arange(8, dtype=np.float32)
    ffm = np.arange(6, dtype=np.float32) * 0.5
    ptrs = (C.c_void_p * 2)(lr.ctypes.data_as(C.c_void_p), ffm.ctypes.data_as(C.c_void_p))
    sizes = (C.c_uint64 * 2)(lr.nbytes, ffm.nbytes)
    path = str(tmp_path / "model.fw")
    err = C.create_string_buffer(512)
    assert L.fwhost_regressor_write(path.encode(), vw.source_json.encode(), mi_json.encode(), 7, ptrs, sizes, 2, err, 512) == 0
    raw = open(path, "rb").read()
    assert raw[:4] == b"FWRE" and int.from_bytes(raw[4:8], "little") == 6
    l1 = int.from_bytes(raw[8:16], "little")
    assert json.loads(raw[16:16 + l1]) == vw.source
    o = 16 + l1
    l2 = int.from_bytes(raw[o:o + 8], "little")
    assert json.loads(raw[o + 8:o + 8 + l2])["optimizer"] == "AdagradLUT"
    o += 8 + l2
    assert int.from_bytes(raw[o:o + 8], "little") == 7 and raw[o + 8:] == lr.tobytes() + ffm.tobytes()
    r = L.fwhost_regressor_open(path.encode(), err, 512)
    assert r and json.loads(L.fwhost_regressor_mi_json(r).decode())["ffm_k"] == 1 and L.fwhost_regressor_weights_len(r) == 7
    back = np.empty(8, np.float32)
    assert L.fwhost_regressor_read(r, back.ctypes.data_as(C.c_void_p), 32) == 0 and np.array_equal(back, lr)
    L.fwhost_regressor_close(r)


def test_regressor_reader_reports_optimizer_and_dequantizes(tmp_path):
    """fwhost_regressor_optimizer / _dequantize come from the parsed ModelInstance (any key order / whitespace), and
    fwhost_regressor_read_quantized restates quantization.rs:77-95 (half-float bucket numbers incl. subnormals)."""
    import ctypes as C

    L = host._L()
    vw = host.VwNamespaceMap.new("A,a\nB,b\n")
    from fwumious_wabbit_b200 import ModelInstance

    mi = ModelInstance.new_empty()
    mi.num_namespaces, mi.optimizer = 2, Optimizer.SGD
    j = json.loads(host.model_instance_to_json(mi, vw))
    j["dequantize_weights"] = True
    compact = json.dumps(j, separators=(",", ":"))       # no space after the colon: a substring search would miss "SGD"
    halves = np.array([0, 1, 2, 1023, 1024, 2049, 40000, 65025, 6e-8, 3e-5], dtype=np.float16)
    payload = np.concatenate([np.array([0.25, -3.0], np.float32).view(np.uint8), halves.view(np.uint8)])
    ptrs = (C.c_void_p * 1)(payload.ctypes.data_as(C.c_void_p))
    sizes = (C.c_uint64 * 1)(payload.nbytes)
    err = C.create_string_buffer(1024)
    path = str(tmp_path / "q.fw").encode()
    assert L.fwhost_regressor_write(path, vw.source_json.encode(), compact.encode(), len(halves), ptrs, sizes, 1, err, 1024) == 0, err.value
    r = L.fwhost_regressor_open(path, err, 1024)
    assert r, err.value
    assert L.fwhost_regressor_optimizer(r) == Optimizer.SGD and L.fwhost_regressor_dequantize(r) == 1
    out = np.empty(len(halves), np.float32)
    assert L.fwhost_regressor_read_quantized(r, out.ctypes.data_as(C.c_void_p), len(halves)) == 0
    L.fwhost_regressor_close(r)
    want = (np.float32(-3.0) + halves.astype(np.float32) * np.float32(0.25)).astype(np.float32)
    assert np.array_equal(out, want)
