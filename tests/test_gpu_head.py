"""GPU parity tests of the dense head (BASELINE config 5: LR + FFM + triangle -> copy -> hidden layers with ReLU ->
join [h, x] -> one neuron -> sigmoid; regressor.rs:191-320, block_neural.rs:196-341, block_relu.rs:79-111,
block_misc.rs:435-519).  Everything goes through the C ABI and is compared with the CPU oracle on identical weights.

Tolerance: 1e-5 per prediction (north_star).  Bit-exactness is not a target here: the reference's own forward is an
MKL sgemv whose summation order is unspecified (pinned to 5e-6 only by block_neural.rs:536-537)."""
import os

import numpy as np
import pytest

import fwumious_wabbit_b200 as fw
from fwumious_wabbit_b200 import ModelInstance, Optimizer, _lib, synth
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-5
SEQUENTIAL = 0x7FFFFFFF


def head_mi(ffm_k=4, F=8, combos=3, widths=(16, 12), relu=(True, True), opt=Optimizer.AdagradLUT, bits=12, ffm_bits=12, init="hu"):
    mi = ModelInstance.new_empty()
    mi.learning_rate, mi.power_t = 0.1, 0.5
    mi.ffm_learning_rate, mi.ffm_power_t = 0.05, 0.5
    mi.nn_learning_rate, mi.nn_power_t, mi.nn_init_acc_gradient = 0.02, 0.45, 0.0
    mi.bit_precision, mi.ffm_k, mi.ffm_bit_precision, mi.optimizer = bits, ffm_k, ffm_bits, opt
    mi.feature_combo_descs = [([j], 1.0) for j in range(combos)]
    mi.ffm_fields = [[j] for j in range(F)] if ffm_k else []
    mi.num_namespaces = max(combos, F)
    mi.nn_layers = [{"width": str(w), "activation": "relu" if r else "none", "init": init} for w, r in zip(widths, relu)]
    return mi


def randomise(ora, rng):
    ora.lr_table[:, 0] = rng.normal(0, 0.2, ora.lr_table.shape[0]).astype(np.float32)
    if ora.ffm_weights is not None and len(ora.ffm_weights):
        ora.ffm_weights[:] = rng.normal(0, 0.3, ora.ffm_weights.shape[0]).astype(np.float32)
    for l in range(ora.nn_layer_count):
        w = ora.nn_weights(l)
        w[:] = rng.normal(0, 0.25, w.shape[0]).astype(np.float32)


def test_head_init_and_block_layout():
    """Layer shapes follow the reference (join [h, x] into one neuron initialised to One, biases 0) and the
    Hu stand-in (parity unpinned, fwgpu.h) is the same on both sides, so from-scratch runs are comparable."""
    mi = head_mi()
    ora = util.oracle_regressor(mi)
    re = fw.Regressor(mi)
    x_len = mi.num_combos + 8 * 9 // 2
    shapes = [(x_len, 16), (16, 12), (12 + x_len, 1)]
    assert re.nn_layer_count() == ora.nn_layer_count == 3
    for l, (n_in, n_out) in enumerate(shapes):
        n, nbytes = re.block_len(_lib.BLOCK_NN0 + l)
        assert n == (n_in + 1) * n_out and nbytes == 8 * n  # weights then accumulators (block_neural.rs:426-438)
        w, acc = re.get_nn(l)
        assert np.array_equal(w.view(np.uint32), ora.nn_weights(l).view(np.uint32))
        assert np.all(w[n_in * n_out:] == 0.0) and np.all(acc == 0.0)
    assert np.all(re.get_nn(2)[0][:-1] == 1.0)


@pytest.mark.parametrize("multi", [False, True])
@pytest.mark.parametrize("shape", [dict(), dict(ffm_k=8, F=5, widths=(20,), relu=(False,)), dict(ffm_k=3, F=4, combos=2, widths=(7, 9, 5), relu=(True, False, True))])
def test_head_predict_parity(shape, multi):
    rng = np.random.default_rng(5 + multi)
    mi = head_mi(**shape)
    ora = util.oracle_regressor(mi)
    randomise(ora, rng)
    re = fw.Regressor(mi)
    util.sync_tables_from_oracle(re, ora)
    d = util.random_csr(rng, 400, mi, multi_valued=multi, empty_prob=0.2 if multi else 0.0, value_one=not multi)
    want = ora.learn_batch(d, update=False)
    got = re.predict_batch(util.csr_from_dict(d))
    assert np.max(np.abs(got - want)) <= TOL, np.max(np.abs(got - want))
    assert np.array_equal(got, re.predict_batch(util.csr_from_dict(d)))  # read-only, deterministic
    for l in range(re.nn_layer_count()):
        assert np.array_equal(re.get_nn(l)[0], ora.nn_weights(l))


@pytest.mark.parametrize("opt", [Optimizer.AdagradLUT, Optimizer.AdagradFlex, Optimizer.SGD])
def test_head_learn_batch1_parity(opt):
    """One example per call = the reference's sequential semantics: every prediction within 1e-5 and every table
    (LR, FFM, all head layers, accumulators included) equal to the oracle's within float noise afterwards."""
    rng = np.random.default_rng(21)
    mi = head_mi(opt=opt)
    ora = util.oracle_regressor(mi)
    randomise(ora, rng)
    re = fw.Regressor(mi)
    util.sync_tables_from_oracle(re, ora)
    n = 200
    d = util.random_csr(rng, n, mi, multi_valued=True, empty_prob=0.1, value_one=False)
    d["importance"][::17] = 0.0  # regressor.rs:366-370: no update, predict-order forward
    d["importance"][5::13] = 0.5
    want = ora.learn_batch(d, update=True)
    batch = util.csr_from_dict(d)
    got = np.array([re.learn_batch(batch.slice(i, i + 1), True)[0] for i in range(n)], dtype=np.float32)
    assert np.max(np.abs(got - want)) <= TOL, np.max(np.abs(got - want))
    np.testing.assert_allclose(re.get_lr_table()[:, 0], ora.lr_table[:, 0], rtol=0, atol=2e-5)
    np.testing.assert_allclose(re.get_ffm()[0], ora.ffm_weights, rtol=0, atol=2e-5)
    for l in range(re.nn_layer_count()):
        w, acc = re.get_nn(l)
        np.testing.assert_allclose(w, ora.nn_weights(l), rtol=0, atol=2e-5)
        if acc is not None:
            np.testing.assert_allclose(acc, ora.nn_acc(l), rtol=2e-3, atol=1e-7)


def small_c5(n_ns=10, k=4):
    """A scaled-down config 5: n_ns single-valued namespaces = fields, FFM k, 2 x 32 ReLU head."""
    w = synth.Workload("c5s", synth._mi(n_ns, ffm_k=k, ffm_bits=14, bits=14, lr=0.05, ffm_lr=0.02, ffm_init_acc=0.1),
                       synth.NS_LETTERS[:n_ns], [50] * 4 + [2000] * (n_ns - 4), "scaled-down c5")
    w.mi.nn_layers = [{"width": "32", "activation": "relu"}, {"width": "32", "activation": "relu"}]
    w.mi.nn_learning_rate, w.mi.nn_power_t, w.mi.nn_init_acc_gradient = 0.02, 0.5, 0.1
    return w


@pytest.mark.parametrize("fast", [True, False])
def test_head_records_sequential_mode(fast, monkeypatch):
    """Raw records, one example in flight, a whole stream in one call, from the (identical) initial weights: the fused
    block-per-record kernel (fast) and the general kernel both follow the oracle's sequential run within 1e-5."""
    monkeypatch.setenv("FWGPU_FAST", "1" if fast else "0")
    w = small_c5()
    w.mi.hogwild_ramp_div = SEQUENTIAL
    n = 3000
    recs = w.records(n)
    recs[::11, 1] = 1  # make sure both labels occur
    ora = util.oracle_regressor(w.mi)
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    _, want = ora.hogwild(spec, recs.reshape(-1), rec_off, 1, want_preds=True)
    re = fw.Regressor(w.mi)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    assert np.max(np.abs(got - want)) <= TOL, np.max(np.abs(got - want))
    # weights: the head's gradient sums are split over blocks and added atomically, so their summation order varies from
    # run to run; once in a few thousand steps that flips an AdaGrad-LUT bucket of one weight (a step ~3e-5 apart)
    for l in range(re.nn_layer_count()):
        np.testing.assert_allclose(re.get_nn(l)[0], ora.nn_weights(l), rtol=0, atol=5e-5)
    np.testing.assert_allclose(re.get_ffm()[0], ora.ffm_weights, rtol=0, atol=5e-5)
    # predict-only pass over the same records on the trained model
    want_p = ora.learn_batch(util.oracle_translate_batch(spec, recs[:500], fixed_len=w.record_len), update=False)
    got_p = re.learn_records(recs[:500].reshape(-1), n_examples=500, update=False)
    assert np.max(np.abs(got_p - want_p)) <= 2 * TOL, np.max(np.abs(got_p - want_p))


def test_head_fast_path_with_leftovers():
    """Records the fused kernel cannot take (multi-valued / weighted namespaces) go through the general kernel inside the
    same sub-batch; the mixed stream equals the all-general run example by example in sequential mode."""
    w = small_c5(n_ns=6)
    w.mi.hogwild_ramp_div = SEQUENTIAL
    rng = np.random.default_rng(3)
    n = 600
    fixed = w.records(n)
    recs, off = [], [0]
    for i in range(n):
        r = fixed[i].copy()
        if i % 5 == 2:  # namespace 1 gets two weighted features: header slot -> dynamic pairs
            dyn = [int(rng.integers(0, 1 << 31)), int(np.float32(0.5).view(np.uint32)), int(rng.integers(0, 1 << 31)), int(np.float32(2.0).view(np.uint32))]
            start = len(r)
            r = np.concatenate([r, np.array(dyn, np.uint32)])
            r[3 + 1] = 0x80000000 | (start << 16) | (start + 4)
            r[0] = len(r)
        recs.append(r)
        off.append(off[-1] + len(r))
    flat, off = np.concatenate(recs).astype(np.uint32), np.array(off, np.uint32)
    outs = []
    for fast in ("1", "0"):
        os.environ["FWGPU_FAST"] = fast
        try:
            re = fw.Regressor(w.mi)
            outs.append((re.learn_records(flat, rec_off=off, update=True), re.get_nn(0)[0], re.get_ffm()[0]))
        finally:
            os.environ.pop("FWGPU_FAST", None)
    assert np.max(np.abs(outs[0][0] - outs[1][0])) <= TOL
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=0, atol=2e-5)
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=0, atol=2e-5)
    ora = util.oracle_regressor(w.mi)
    _, want = ora.hogwild(util.oracle_spec(w.mi), flat, off.astype(np.uint64), 1, want_preds=True)
    assert np.max(np.abs(outs[0][0] - want)) <= TOL, np.max(np.abs(outs[0][0] - want))


def test_head_hogwild_progressive_logloss():
    """Sub-batches of thousands of examples around the head's GEMMs (Hogwild on device): progressive logloss within 2 %
    (relative) of the sequential oracle on the same stream; sub-batch size stated: 2048."""
    os.environ["FWGPU_HEAD_BATCH"] = "2048"
    try:
        w = small_c5()
        n = 300_000
        recs = w.records(n)
        re = fw.Regressor(w.mi)
        got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    finally:
        os.environ.pop("FWGPU_HEAD_BATCH", None)
    ora = util.oracle_regressor(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    _, want = ora.hogwild(util.oracle_spec(w.mi), recs.reshape(-1), rec_off, 1, want_preds=True)
    labels = (recs[:, 1] == 1).astype(np.float32)
    ll_gpu, ll_ref = util.logloss(got, labels), util.logloss(want, labels)
    prior = util.logloss(np.full(n, labels.mean()), labels)
    assert np.all(np.isfinite(got))
    assert ll_gpu < prior, (ll_gpu, prior)                  # it learns
    assert abs(ll_gpu - ll_ref) / ll_ref < 0.02, (ll_gpu, ll_ref)


def test_head_full_size_config5_properties():
    """BASELINE config 5 at full width (39 fields, k=8, ffm_bit_precision 24, 2 x 256 ReLU): first examples against the
    oracle in sequential mode, then size-independent properties on a larger Hogwild batch (finite, in (0,1), predict is
    idempotent and read-only, export -> import round trip reproduces predictions bit for bit)."""
    w = synth.workload("c5")
    w.mi.hogwild_ramp_div = SEQUENTIAL
    n = 60
    recs = w.records(n)
    ora = util.oracle_regressor(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    _, want = ora.hogwild(util.oracle_spec(w.mi), recs.reshape(-1), rec_off, 1, want_preds=True)
    re = fw.Regressor(w.mi)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    assert np.max(np.abs(got - want)) <= TOL, np.max(np.abs(got - want))
    w2 = synth.workload("c5")
    m = 50_000
    recs2 = w2.records(m, first=1000)
    re2 = fw.Regressor(w2.mi)
    p = re2.learn_records(recs2.reshape(-1), n_examples=m, update=True)
    assert np.all(np.isfinite(p)) and p.min() > 0.0 and p.max() < 1.0
    a = re2.learn_records(recs2[:5000].reshape(-1), n_examples=5000, update=False)
    b = re2.learn_records(recs2[:5000].reshape(-1), n_examples=5000, update=False)
    assert np.array_equal(a, b)
    re3 = fw.Regressor(w2.mi)
    for blk in [_lib.BLOCK_LR, _lib.BLOCK_FFM] + [_lib.BLOCK_NN0 + l for l in range(re2.nn_layer_count())]:
        re3.import_block(blk, re2.export_block(blk), True)
    c = re3.learn_records(recs2[:5000].reshape(-1), n_examples=5000, update=False)
    assert np.array_equal(a, c)


def test_head_regressor_file_roundtrip(tmp_path):
    """Regressor file with a head (persistence.rs:55-97): LR, FFM, then every neuron layer as weights + accumulators
    (block_neural.rs:426-438); mutable reload resumes with the same state, immutable reload (weights only) and the
    converted inference file predict identically."""
    from fwumious_wabbit_b200 import host

    w = small_c5()
    vw = host.VwNamespaceMap.new("".join(f"{c},feature{c}\n" for c in w.ns_names))
    n = 20_000
    recs = w.records(n)
    re = fw.Regressor(w.mi)
    re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    path = str(tmp_path / "c5s.fw")
    host.save_regressor_to_filename(path, w.mi, vw, re)
    raw = open(path, "rb").read()
    l1 = int.from_bytes(raw[8:16], "little")
    l2 = int.from_bytes(raw[16 + l1:24 + l1], "little")
    body = 24 + l1 + l2
    x_len = w.mi.num_combos + 10 * 11 // 2
    nn_len = (x_len + 1) * 32 + (32 + 1) * 32 + (32 + x_len + 1)
    assert int.from_bytes(raw[body:body + 8], "little") == (1 << 14) + (1 << 14) + 40 + nn_len
    assert len(raw) == body + 8 + 8 * ((1 << 14) + (1 << 14) + 40 + nn_len)  # every block: weights + optimizer state
    want = re.learn_records(recs[:3000].reshape(-1), n_examples=3000, update=False)
    mi2, _, re2 = host.new_regressor_from_filename(path, False)
    assert [dict(l) for l in mi2.nn_layers] == [dict(l) for l in w.mi.nn_layers]
    assert np.array_equal(re2.learn_records(recs[:3000].reshape(-1), n_examples=3000, update=False), want)
    for l in range(re.nn_layer_count()):
        assert np.array_equal(re.get_nn(l)[0], re2.get_nn(l)[0]) and np.array_equal(re.get_nn(l)[1], re2.get_nn(l)[1])
    _, _, re3 = host.new_regressor_from_filename(path, True)
    assert re3.get_nn(0)[1] is None
    assert np.array_equal(re3.learn_records(recs[:3000].reshape(-1), n_examples=3000, update=False), want)
    inf = str(tmp_path / "c5s_inference.fw")
    host.save_regressor_to_filename(inf, w.mi, vw, re3)
    _, _, re4 = host.new_regressor_from_filename(inf, True)
    assert np.array_equal(re4.learn_records(recs[:3000].reshape(-1), n_examples=3000, update=False), want)
    with pytest.raises(fw._lib.FwgpuError):
        re3.learn_records(recs[:10].reshape(-1), n_examples=10, update=True)  # regressor.rs:362-365


def test_umma_gemm_tiles_against_double_precision():
    """The tcgen05 3xTF32 GEMM tiles the head uses for sub-batches >= 512 rows (csrc/fwgpu_umma.cuh): all operand layouts and
    epilogue kinds against a double-precision CPU product, ragged shapes included (tools/umma_gemm_test.cu prints one line per
    case and the measured error; tolerance 2e-5 of the largest output, i.e. fp32-sum level)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(fw.__file__)), "umma_gemm_test")
    assert os.path.exists(exe), "umma_gemm_test is built by fwumious_wabbit_b200/build.py"
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "all cases ok" in res.stdout, res.stdout + res.stderr
