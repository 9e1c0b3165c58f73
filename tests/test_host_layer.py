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


def _quantize_restated(w):
    """quantization.rs:19-75 in numpy: stats rounded to 4 decimals, 65 025 buckets, round half away from zero, f16 (RNE)."""
    w = np.asarray(w, np.float32)
    rnd = lambda x: (np.sign(x) * np.floor(np.abs(x) + np.float32(0.5))).astype(np.float32)  # f32::round
    lo = np.float32(rnd(w.min() * np.float32(10000.0)) / np.float32(10000.0))
    hi = np.float32(rnd(w.max() * np.float32(10000.0)) / np.float32(10000.0))
    inc = np.float32(np.float32(hi - lo) / np.float32(65025.0))
    with np.errstate(over="ignore"):
        buckets = rnd(((w - lo) / inc).astype(np.float32)).astype(np.float16)
    return inc, lo, buckets


def test_quantize_ffm_weights_follows_the_reference_writer():  # quantization.rs:41-75 and its tests :98-150
    import ctypes as C

    L = host._L()
    # the reference's own test vector (quantization.rs:101-116): statistics and the length of the output
    w = np.array([0.51, 0.12, 0.11, 0.1232, 0.6123, 0.23], np.float32)
    mean = C.c_float(0)
    out = np.empty(8 + 2 * w.size, np.uint8)
    assert L.fwhost_quantize_ffm_weights(w.ctypes.data_as(C.c_void_p), w.size, out.ctypes.data_as(C.c_void_p), C.byref(mean)) == 0
    inc, lo = out[:8].view(np.float32)
    assert out.size // 2 == 10                                    # test_quantize: 4 header pairs + 6 weights
    assert mean.value == np.float32(0.51) and lo == np.float32(0.11)  # test_emit_statistics (mean samples every 10th weight)
    assert inc == np.float32(np.float32(np.float32(0.6123) - np.float32(0.11)) / np.float32(65025.0))
    # round trip (test_dequantize, quantization.rs:118-150): some weights come back exactly, all of them within 1e-4
    back = lo + out[8:].view(np.float16).astype(np.float32) * inc
    assert np.max(np.abs(back - w)) < 1e-4 and (back == w).sum() != 0
    # every bucket number 0 .. 65 025: increment is exactly 1, the half-float conversion sees every integer incl. all its ties
    ramp = np.arange(65026, dtype=np.float32)
    out = host.quantize_ffm_weights(ramp)
    assert np.array_equal(out[:8].view(np.float32), np.array([1.0, 0.0], np.float32))
    assert np.array_equal(out[8:].view(np.uint16), ramp.astype(np.float16).view(np.uint16))
    # a trained-looking block: bit-identical with the numpy restatement, header and buckets
    rng = np.random.default_rng(11)
    for scale, n in ((0.02, 200_001), (3.0, 50_000), (1e-4, 1000)):
        w = rng.normal(0, scale, n).astype(np.float32)
        w[::97] = 0.0
        inc, lo, buckets = _quantize_restated(w)
        out = host.quantize_ffm_weights(w)
        assert np.array_equal(out[:8].view(np.float32), np.array([inc, lo], np.float32))
        assert np.array_equal(out[8:].view(np.uint16), buckets.view(np.uint16))
    with pytest.raises(ValueError):
        host.quantize_ffm_weights(np.empty(0, np.float32))


def test_quantized_file_written_by_the_host_layer_reads_back(tmp_path):  # block_ffm.rs:835-857, main.rs:140-147
    import ctypes as C

    L = host._L()
    vw = host.VwNamespaceMap.new("A,a\nB,b\n")
    mi_json = host.model_instance_json_from_cmdline(["--keep", "A", "--ffm_k", "2", "--ffm_field", "A", "--ffm_field", "B", "--adaptive",
                                                     "--ffm_bit_precision", "10"], vw)
    err = C.create_string_buffer(1024)
    p = L.fwhost_model_instance_for_save(mi_json.encode(), 1, 1, err, 1024)
    assert p, err.value
    saved = json.loads(C.string_at(p).decode())
    L.fwhost_free(p)
    before = json.loads(mi_json)
    assert before["optimizer"] == "AdagradLUT" and before["dequantize_weights"] is False
    assert saved["optimizer"] == "SGD" and saved["dequantize_weights"] is True
    assert {k: v for k, v in saved.items() if k not in ("optimizer", "dequantize_weights")} == \
           {k: v for k, v in before.items() if k not in ("optimizer", "dequantize_weights")}
    p = L.fwhost_model_instance_for_save(mi_json.encode(), 0, 0, err, 1024)
    assert json.loads(C.string_at(p).decode()) == before
    L.fwhost_free(p)
    assert not L.fwhost_model_instance_for_save(b"{not json", 1, 1, err, 1024) and err.value
    rng = np.random.default_rng(5)
    lr_w = rng.normal(0, 0.1, 1 << 18).astype(np.float32)
    w = rng.normal(0, 0.05, (1 << 10) + 4).astype(np.float32)
    blocks = [lr_w, host.quantize_ffm_weights(w)]
    ptrs = (C.c_void_p * 2)(*[b.ctypes.data_as(C.c_void_p) for b in blocks])
    sizes = (C.c_uint64 * 2)(*[b.nbytes for b in blocks])
    path = str(tmp_path / "q.fw").encode()
    assert L.fwhost_regressor_write(path, vw.source_json.encode(), json.dumps(saved).encode(), lr_w.size + w.size, ptrs, sizes, 2, err, 1024) == 0
    r = L.fwhost_regressor_open(path, err, 1024)
    assert r and L.fwhost_regressor_optimizer(r) == Optimizer.SGD and L.fwhost_regressor_dequantize(r) == 1
    got_lr, got = np.empty_like(lr_w), np.empty_like(w)
    assert L.fwhost_regressor_read(r, got_lr.ctypes.data_as(C.c_void_p), lr_w.nbytes) == 0 and np.array_equal(got_lr, lr_w)
    assert L.fwhost_regressor_read_quantized(r, got.ctypes.data_as(C.c_void_p), w.size) == 0
    L.fwhost_regressor_close(r)
    inc, lo, buckets = _quantize_restated(w)
    assert np.array_equal(got, (lo + buckets.astype(np.float32) * inc).astype(np.float32))
    assert np.max(np.abs(got - w)) <= 17 * inc    # buckets above 32 768 are 32 apart in half precision


class _TablesOnly:
    """Stands in for a regressor where there is no GPU: the two accessors save_regressor_to_filename uses."""

    def __init__(self, blocks, immutable):
        self.blocks, self.immutable = blocks, immutable

    def block_len(self, b):
        n, payload = self.blocks[b]
        return n, payload.nbytes

    def export_block(self, b):
        return self.blocks[b][1]


@pytest.mark.parametrize("immutable", [True, False])
def test_save_regressor_with_weight_quantization(tmp_path, immutable):  # persistence.rs:55-97, block_ffm.rs:835-848, main.rs:136-148
    import ctypes as C
    from fwumious_wabbit_b200 import ModelInstance, _lib

    vw = host.VwNamespaceMap.new("A,a\nB,b\n")
    mi = ModelInstance.new_empty()
    mi.bit_precision, mi.ffm_k, mi.ffm_bit_precision, mi.ffm_fields, mi.num_namespaces = 10, 2, 10, [[0], [1]], 2
    mi.optimizer = Optimizer.SGD if immutable else Optimizer.AdagradLUT
    rng = np.random.default_rng(8)
    n_lr, n_f = 1 << 10, (1 << 10) + 4
    per = 1 if immutable else 2
    lr = rng.normal(0, 0.1, n_lr * per).astype(np.float32)
    ffm = np.concatenate([rng.normal(0, 0.05, n_f), rng.random(n_f) if not immutable else []]).astype(np.float32)
    re = _TablesOnly({_lib.BLOCK_LR: (n_lr, lr), _lib.BLOCK_FFM: (n_f, ffm)}, immutable)
    path = str(tmp_path / "q.fw")
    host.save_regressor_to_filename(path, mi, vw, re, quantize_weights=True)
    L = host._L()
    err = C.create_string_buffer(512)
    r = L.fwhost_regressor_open(path.encode(), err, 512)
    assert r, err.value
    # only the conversion to an inference regressor records the quantization in the file (main.rs:143-145)
    assert L.fwhost_regressor_dequantize(r) == (1 if immutable else 0)
    assert L.fwhost_regressor_optimizer(r) == mi.optimizer and L.fwhost_regressor_weights_len(r) == n_lr + n_f
    got_lr, got_w, got_acc = np.empty_like(lr), np.empty(n_f, np.float32), np.empty(n_f, np.float32)
    assert L.fwhost_regressor_read(r, got_lr.ctypes.data_as(C.c_void_p), lr.nbytes) == 0 and np.array_equal(got_lr, lr)
    assert L.fwhost_regressor_read_quantized(r, got_w.ctypes.data_as(C.c_void_p), n_f) == 0
    inc, lo, buckets = _quantize_restated(ffm[:n_f])
    assert np.array_equal(got_w, (lo + buckets.astype(np.float32) * inc).astype(np.float32))
    if not immutable:  # the accumulators follow the buckets untouched
        assert L.fwhost_regressor_read(r, got_acc.ctypes.data_as(C.c_void_p), 4 * n_f) == 0 and np.array_equal(got_acc, ffm[n_f:])
    assert L.fwhost_regressor_read(r, got_acc.ctypes.data_as(C.c_void_p), 1) != 0  # nothing after the last block
    L.fwhost_regressor_close(r)
    # and without the flag the file is the plain one
    host.save_regressor_to_filename(path, mi, vw, re)
    assert open(path, "rb").read().endswith(lr.tobytes() + ffm.tobytes())


def test_cache_reader_edge_cases(tmp_path):  # cache.rs:133-182, 187-232: header checks, ragged records, an empty cache
    vw = host.VwNamespaceMap.new("A,a\nB,b\n")
    p = host.VowpalParser(vw)
    lines = ["1 |A x |B y", "-1 2.5 |A x:2 z |B y", "|B q", "1 |A a b c d |B e:0.5"]
    recs, offs = p.parse_text("\n".join(lines) + "\n")
    assert np.diff(offs).tolist() == [5, 9, 5, 15]                     # ragged: inline slots and (hash, value) lists
    path = str(tmp_path / "ragged.vw.fwcache")
    host.cache_write(path, vw, recs)
    r, off, _ = host.cache_read(path, vw)
    assert np.array_equal(r, recs) and np.array_equal(off, offs)
    good = open(path, "rb").read()
    blob_len = int.from_bytes(good[8:16], "little")
    body = 16 + blob_len

    def damaged(data):
        q = str(tmp_path / "damaged.vw.fwcache")
        open(q, "wb").write(data)
        return q

    with pytest.raises(IOError, match="version of the cache file: 10"):
        host.cache_read(damaged(good[:4] + (10).to_bytes(4, "little") + good[8:]), vw)
    with pytest.raises(IOError, match="truncated cache header"):
        host.cache_read(damaged(good[:16 + blob_len // 2]), vw)
    with pytest.raises(IOError, match="corrupt record length"):      # a record that claims more words than the file holds
        host.cache_read(damaged(good[:body] + (99).to_bytes(4, "little") + good[body + 4:]), vw)
    with pytest.raises(IOError, match="corrupt record length"):      # a record shorter than its own header
        host.cache_read(damaged(good[:body] + (2).to_bytes(4, "little") + good[body + 4:]), vw)
    r, off, _ = host.cache_read(damaged(good[:body]), vw)               # header only: no records
    assert r.size == 0 and off.tolist() == [0]
    r, off, _ = host.cache_read(damaged(good + b"\x01\x02"), vw)        # trailing bytes short of a word are ignored
    assert np.array_equal(r, recs)
    with pytest.raises(IOError, match="cannot open"):
        host.cache_read(str(tmp_path / "absent.fwcache"), vw)


def test_parse_text_threads_agree(tmp_path):
    """The batch parser gives the same records and offsets whatever the thread count (slabs are placed in line order), with
    one-byte and longer namespace names and a last line without its newline; a blank line is an error, as in the reference
    (parser.rs:226-258: anything that starts with neither a label nor '|' must be a command)."""
    vw = host.VwNamespaceMap.new("A,a\nBB,b\nC,c:f32\n")
    rnd = np.random.default_rng(2)
    lines = []
    for i in range(20_000):
        parts = ["1" if rnd.random() < 0.5 else "-1"]
        if rnd.random() < 0.2:
            parts.append("%.2f" % (rnd.random() * 3))
        parts.append("|A " + " ".join("f%d" % rnd.integers(0, 50) for _ in range(rnd.integers(1, 4))))
        if rnd.random() < 0.7:
            parts.append("|BB:%s w%d:%.1f" % ("2" if rnd.random() < 0.3 else "1", rnd.integers(0, 9), rnd.random()))
        if rnd.random() < 0.5:
            parts.append("|C %.3f" % rnd.normal())
        lines.append(" ".join(parts))
    text = "\n".join(lines)                                             # no newline after the last line
    p = host.VowpalParser(vw)
    base_r, base_o = p.parse_text(text, threads=1)
    assert len(base_o) - 1 == sum(1 for l in lines if l)
    for th in (2, 3, 8):
        r, o = p.parse_text(text, threads=th)
        assert np.array_equal(r, base_r) and np.array_equal(o, base_o)
    one = np.concatenate([p.next_vowpal(l + "\n") for l in lines if l])
    assert np.array_equal(one, base_r)
    with pytest.raises(ValueError, match=r"Cannot parse an example \(line 3\)"):
        p.parse_text("1 |A x\n1 |A y\n\n1 |A z\n")
    with pytest.raises(ValueError, match=r"not predeclared in vw_namespace_map.csv: B \(line 2\)"):
        p.parse_text("1 |A x\n1 |B y\n" * 2000, threads=4)


def test_parser_fuzz_against_the_oracle_parser():
    """Two independent restatements of parser.rs:214-461 — the host layer's (C++, cursor + record writer) and the oracle's (C, the
    reference's loop structure) — on 60 000 generated lines: labels, importances, tags, unknown / repeated / weighted / multi-byte
    / f32 namespaces, weighted and empty features, runs of spaces, commands and garbage.  Same records bit for bit, or the same
    kind of failure with the same message."""
    import random

    vw = host.VwNamespaceMap.new("A,a\nBB,b\nC,c:f32\nD,d\n")
    names = [None] * vw.num_namespaces
    for e in vw.source["entries"]:
        names[e["namespace_index"]] = e["namespace_vwname"]
    hp = host.VowpalParser(vw)
    op = fo.Parser(names, ns_is_f32=vw.ns_is_f32(), namespace_skip_prefix=vw.source["namespace_skip_prefix"])
    rnd = random.Random(7)
    good_floats = ["1", "1.0", "2", "0.5", "1e3", "1e-3", ".5", "5.", "NONE", "-0.0", "+2", "inf", "nan"]
    bad_floats = ["abc", "", "1.5e", "0x10", "1_0", "١", "-1"]

    def number(p_bad):
        return rnd.choice(bad_floats if rnd.random() < p_bad else good_floats)

    def feature(f32):
        if f32:
            w = rnd.choice(["3.5", "-2", "1e2", "0", "NONE"]) if rnd.random() < 0.93 else rnd.choice(["x", "", "1.2.3"])
        else:
            w = rnd.choice(["x", "y1", "feat", "", "Z" * rnd.randint(1, 12), "3.5", "é", "|", "q\t", "a-b_c"])
        if rnd.random() < (0.03 if f32 else 0.25):
            w += ":" + number(0.05)
        return w

    def line():
        r = rnd.random()
        parts = ["1" if r < 0.42 else "-1" if r < 0.84 else "" if r < 0.94 else rnd.choice(["0", "2", "1.0", "-", "x", "flush", "hogwild_load f", "hogwild_load", " 1"])]
        if rnd.random() < 0.2:
            parts.append(number(0.1))
        if rnd.random() < 0.1:
            parts.append("'tag")
        for _ in range(rnd.randint(0, 5)):
            ns = rnd.choice(["A", "BB", "C", "D", "A", "D"]) if rnd.random() < 0.97 else rnd.choice(["E", "", "B"])
            parts.append("|" + ns + (":" + number(0.05) if rnd.random() < (0.02 if ns == "C" else 0.15) else ""))
            parts.extend(feature(ns == "C") for _ in range(rnd.randint(0, 4)))
        s = "".join(p + " " * (1 if rnd.random() < 0.9 else rnd.randint(0, 3)) for p in parts)
        return (s.rstrip(" ") if rnd.random() < 0.5 else s) + "\n"

    def outcome(parse, flush, hogwild, l):
        try:
            return "ok", parse(l).tolist()
        except ValueError as e:
            return "error", str(e)[:40]
        except flush:
            return "flush", None
        except hogwild:
            return "hogwild_load", None

    kinds = {}
    for _ in range(60_000):
        l = line()
        a = outcome(op.parse, fo.FlushCommand, fo.HogwildLoadCommand, l)
        b = outcome(hp.next_vowpal, host.FlushCommand, host.HogwildLoadCommand, l)
        assert a == b, (l, a, b)
        kinds[a[0]] = kinds.get(a[0], 0) + 1
    assert kinds["ok"] > 20_000 and kinds["error"] > 5_000 and kinds["flush"] > 10 and kinds["hogwild_load"] > 10, kinds
