"""Regressor files on the GPU path: the reference's save/load tests (persistence.rs:250-421) through
host.save_regressor_to_filename / new_regressor_from_filename, and interchange with the oracle's tables."""
import json

import numpy as np
import pytest

import fwumious_wabbit_b200 as fw
from fwumious_wabbit_b200 import FeatureBuffer, HashAndValue, HashAndValueAndSeq, ModelInstance, Optimizer, _lib, host, synth
from tests import util

pytestmark = pytest.mark.gpu


def close(a, b, tol=1e-5):
    assert abs(float(a) - float(b)) <= tol, (float(a), float(b))


VW = "A,featureA\nB,featureB\n"


def test_save_load_and_test_mode_lr(tmp_path):  # persistence.rs:250-313
    vw = host.VwNamespaceMap.new(VW)
    mi = ModelInstance.new_empty()
    mi.learning_rate, mi.power_t, mi.bit_precision, mi.optimizer, mi.init_acc_gradient = 0.1, 0.5, 18, Optimizer.AdagradFlex, 0.0
    mi.num_namespaces = 2
    re = fw.Regressor(mi)
    fbuf = FeatureBuffer(label=0.0, lr_buffer=[HashAndValue(1, 1.0, 0), HashAndValue(2, 1.0, 0)])
    close(re.learn(fbuf, True), 0.5)
    close(re.learn(fbuf, True), 0.45016602)
    close(re.learn(fbuf, False), 0.41731137)
    path = str(tmp_path / "test_regressor2.fw")
    host.save_regressor_to_filename(path, mi, vw, re)
    mi2, vw2, re2 = host.new_regressor_from_filename(path, immutable=False)
    assert mi2.optimizer == Optimizer.AdagradFlex and vw2.source == vw.source
    close(re2.learn(fbuf, False), 0.41731137)
    close(re2.predict(fbuf), 0.41731137)
    assert np.array_equal(re2.get_lr_table(), re.get_lr_table())       # weights AND accumulators came back
    close(re2.learn(fbuf, True), 0.41731137)                           # training resumes from the saved state (--save_resume)
    mi3, _, re3 = host.new_regressor_from_filename(path, immutable=True)
    assert re3.get_name() == 'Regressor with optimizer "SGD"'
    close(re3.predict(fbuf), 0.41731137)
    with pytest.raises(_lib.FwgpuError):
        re3.learn(fbuf, True)
    # -l / --power_t on the command line override the stored hyper-parameters (model_instance.rs:497-550)
    mi4, _, _ = host.new_regressor_from_filename(path, immutable=False, cmd_arguments=["-l", "0.25", "--ffm_power_t", "0.3"])
    assert mi4.learning_rate == 0.25 and abs(mi4.ffm_power_t - 0.3) < 1e-7 and abs(mi4.power_t - 0.5) < 1e-7


def test_save_load_and_test_mode_ffm_and_inference_file(tmp_path):  # persistence.rs:341-421 + main.rs:136-149
    vw = host.VwNamespaceMap.new(VW)
    mi = ModelInstance.new_empty()
    mi.learning_rate, mi.power_t, mi.bit_precision = 0.1, 0.0, 18
    mi.ffm_k, mi.ffm_bit_precision, mi.ffm_power_t, mi.ffm_learning_rate = 1, 18, 0.0, 0.1
    mi.ffm_fields, mi.optimizer, mi.num_namespaces = [[], []], Optimizer.AdagradFlex, 2
    re = fw.Regressor(mi)
    n, _ = re.block_len(_lib.BLOCK_FFM)
    re.set_ffm(np.ones(n, np.float32), np.zeros(n, np.float32))
    fbuf = FeatureBuffer(label=0.0, ffm_buffer=[HashAndValueAndSeq(1, 1.0, 0), HashAndValueAndSeq(3000, 1.0, 0), HashAndValueAndSeq(100, 2.0, 1)])
    close(re.learn(fbuf, True), 0.9933072)
    close(re.learn(fbuf, False), 0.9395168)
    path = str(tmp_path / "m.fw")
    host.save_regressor_to_filename(path, mi, vw, re)
    _, _, re2 = host.new_regressor_from_filename(path, False)
    assert re2.get_name() == 'Regressor with optimizer "AdagradFlex"'
    close(re2.learn(fbuf, False), 0.9395168)
    close(re2.predict(fbuf), 0.9395168)
    _, _, re3 = host.new_regressor_from_filename(path, True)
    assert re3.get_name() == 'Regressor with optimizer "SGD"'
    close(re3.predict(fbuf), 0.9395168)
    # --convert_inference_regressor: write the immutable regressor (weights only), reload, same predictions
    # (the assertion of examples/ffm/run_fw_with_prediction_tests.sh:130-137)
    inf = str(tmp_path / "inference.fw")
    host.save_regressor_to_filename(inf, mi, vw, re3)
    full_size, inf_size = (tmp_path / "m.fw").stat().st_size, (tmp_path / "inference.fw").stat().st_size
    assert inf_size < full_size * 0.6
    mi5, _, re5 = host.new_regressor_from_filename(inf, True)
    assert mi5.optimizer == Optimizer.SGD
    close(re5.predict(fbuf), 0.9395168)


def test_file_interchange_with_oracle_tables(tmp_path):
    """A model trained on the GPU, saved, and its payload loaded into the CPU oracle predicts the same (and the other
    way round): the block byte layouts are the reference's (block_helpers.rs:23-28, block_ffm.rs:835-848)."""
    w = synth.workload("c2")
    vw = host.VwNamespaceMap.new("".join(f"{c},feature{c}\n" for c in w.ns_names))
    recs = w.records(50_000)
    re = fw.Regressor(w.mi)
    re.learn_records(recs.reshape(-1), n_examples=50_000, update=True)
    path = str(tmp_path / "c2.fw")
    host.save_regressor_to_filename(path, w.mi, vw, re)
    raw = open(path, "rb").read()
    l1 = int.from_bytes(raw[8:16], "little")
    l2 = int.from_bytes(raw[16 + l1:24 + l1], "little")
    body = 24 + l1 + l2
    assert int.from_bytes(raw[body:body + 8], "little") == (1 << 18) + (1 << 20) + 32
    payload = np.frombuffer(raw, dtype=np.float32, offset=body + 8)
    ora = util.oracle_regressor(w.mi)
    n_lr, n_f = 1 << 18, (1 << 20) + 32
    ora.lr_table[:] = payload[: 2 * n_lr].reshape(n_lr, 2)
    ora.ffm_weights[:] = payload[2 * n_lr: 2 * n_lr + n_f]
    ora.ffm_acc[:] = payload[2 * n_lr + n_f: 2 * n_lr + 2 * n_f]
    m = 3000
    d = util.oracle_translate_batch(util.oracle_spec(w.mi), recs[:m], fixed_len=w.record_len)
    want = ora.learn_batch(d, update=False)
    _, _, re2 = host.new_regressor_from_filename(path, True)
    got = re2.learn_records(recs[:m].reshape(-1), n_examples=m, update=False)
    assert np.max(np.abs(got - want)) <= 1e-5


def test_quantized_inference_file_is_dequantized_on_load(tmp_path):
    """A file written with --weight_quantization (main.rs:140-147, quantization.rs:41-95) stores the FFM block as an 8-byte
    header {increment, min} and one half-float bucket number per weight; loading it dequantizes (persistence.rs:144-161,
    block_ffm.rs:853-857) instead of mis-reading the block as f32."""
    import ctypes as C

    vw = host.VwNamespaceMap.new(VW)
    mi = ModelInstance.new_empty()
    mi.learning_rate, mi.power_t, mi.bit_precision = 0.1, 0.0, 10
    mi.ffm_k, mi.ffm_bit_precision, mi.ffm_power_t, mi.ffm_learning_rate = 2, 10, 0.0, 0.1
    mi.ffm_fields, mi.optimizer, mi.num_namespaces = [[0], [1]], Optimizer.SGD, 2
    rng = np.random.default_rng(4)
    n_lr, n_ffm = 1 << 10, (1 << 10) + 4
    lr_w = rng.normal(0, 0.1, n_lr).astype(np.float32)
    w = rng.normal(0, 0.2, n_ffm).astype(np.float32)
    # quantize_ffm_weights: min/max rounded to 1e-4, 65025 buckets, bucket number stored as f16
    wmin = np.float32(np.round(w.min() * np.float32(10000.0)) / np.float32(10000.0))
    wmax = np.float32(np.round(w.max() * np.float32(10000.0)) / np.float32(10000.0))
    inc = np.float32((wmax - wmin) / np.float32(65025.0))
    buckets = np.round((w - wmin) / inc).astype(np.float16)
    payload_ffm = np.concatenate([np.array([inc, wmin], np.float32).view(np.uint8), buckets.view(np.uint8)])
    j = json.loads(host.model_instance_to_json(mi, vw))
    j["dequantize_weights"] = True
    path = str(tmp_path / "q.fw")
    L = host._L()
    blocks = [lr_w.view(np.uint8), payload_ffm]
    ptrs = (C.c_void_p * 2)(*[b.ctypes.data_as(C.c_void_p) for b in blocks])
    sizes = (C.c_uint64 * 2)(*[b.nbytes for b in blocks])
    err = C.create_string_buffer(1024)
    assert L.fwhost_regressor_write(path.encode(), vw.source_json.encode(), json.dumps(j).encode(), n_lr + n_ffm, ptrs, sizes, 2, err, 1024) == 0, err.value
    _, _, re = host.new_regressor_from_filename(path, immutable=True)
    got, _ = re.get_ffm()
    want = (wmin + buckets.astype(np.float32) * inc).astype(np.float32)
    assert np.array_equal(got, want)
    assert np.max(np.abs(got - w)) < 0.01            # half-float bucket numbers: coarse above bucket 2048, as in the reference
    assert np.array_equal(re.get_lr_table()[:, 0], lr_w)


def test_weight_quantization_written_by_the_gpu_path(tmp_path):
    """--convert_inference_regressor --weight_quantization (main.rs:136-148): the inference regressor's FFM block goes to the file
    as the 8-byte header and one half-float bucket per weight (quantization.rs:41-75), its ModelInstance says SGD and
    dequantize_weights = true, and loading the file gives back exactly what the reference's dequantizer computes from it."""
    w = synth.workload("c2")
    vw = host.VwNamespaceMap.new("".join(f"{c},feature{c}\n" for c in w.ns_names))
    recs = w.records(60_000)
    re = fw.Regressor(w.mi)
    re.learn_records(recs[:50_000].reshape(-1), n_examples=50_000, update=True)
    full, inf, q = (str(tmp_path / n) for n in ("full.fw", "inf.fw", "q.fw"))
    host.save_regressor_to_filename(full, w.mi, vw, re)
    mi_i, _, re_i = host.new_regressor_from_filename(full, immutable=True)
    host.save_regressor_to_filename(inf, mi_i, vw, re_i)
    host.save_regressor_to_filename(q, mi_i, vw, re_i, quantize_weights=True)
    n_lr, n_f = 1 << 18, (1 << 20) + 32
    import os
    assert os.path.getsize(inf) - os.path.getsize(q) == 2 * n_f - 8 + len('"dequantize_weights": false') - len('"dequantize_weights": true')
    mi_q, _, re_q = host.new_regressor_from_filename(q, immutable=True)
    assert mi_q.optimizer == Optimizer.SGD
    raw = open(q, "rb").read()
    l1 = int.from_bytes(raw[8:16], "little")
    l2 = int.from_bytes(raw[16 + l1:24 + l1], "little")
    assert json.loads(raw[24 + l1:24 + l1 + l2])["dequantize_weights"] is True
    body = 24 + l1 + l2 + 8
    assert np.array_equal(np.frombuffer(raw, np.float32, n_lr, body), re_i.get_lr_table()[:, 0])   # the LR block is not quantized
    inc, lo = np.frombuffer(raw, np.float32, 2, body + 4 * n_lr)
    buckets = np.frombuffer(raw, np.float16, n_f, body + 4 * n_lr + 8)
    assert len(raw) == body + 4 * n_lr + 8 + 2 * n_f
    w_full, _ = re_i.get_ffm()
    rnd = lambda x: (np.sign(x) * np.floor(np.abs(x) + np.float32(0.5))).astype(np.float32)
    assert lo == np.float32(rnd(w_full.min() * np.float32(1e4)) / np.float32(1e4))
    assert np.array_equal(buckets.view(np.uint16), rnd(((w_full - lo) / inc).astype(np.float32)).astype(np.float16).view(np.uint16))
    w_q, _ = re_q.get_ffm()
    assert np.array_equal(w_q, (lo + buckets.astype(np.float32) * inc).astype(np.float32))
    assert np.max(np.abs(w_q - w_full)) <= 17 * inc
    # what the quantization costs on held-out records: small against the spread of the predictions
    held = recs[50_000:].reshape(-1)
    p_i = re_i.learn_records(held, n_examples=10_000, update=False)
    p_q = re_q.learn_records(held, n_examples=10_000, update=False)
    print("quantized vs plain: max |dp| %.2e, mean |dp| %.2e, std p %.3f, max |dw| / increment %.2f" % (np.max(np.abs(p_i - p_q)), np.mean(np.abs(p_i - p_q)), np.std(p_i), np.max(np.abs(w_q - w_full)) / inc))
    assert np.max(np.abs(p_i - p_q)) < 0.02 and np.mean(np.abs(p_i - p_q)) < 2e-3
