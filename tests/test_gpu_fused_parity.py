"""The kernels bench.py times, pinned to the oracle in UPDATE mode (run on the B200 box: pytest -m gpu).

test_gpu_parity.py's bit-exact sequential tests run the general kernel (k_learn, exact_order); the throughput numbers come
from the fused kernels -- k_learn_fixed<16,4,1,LUT> on c2, k_learn_fixed_cta on c3 / c4, its two-phase form plus the tcgen05
head GEMMs on c5.  Here those kernels run with ONE record in flight (hogwild_max_inflight = 1: the reference's sequential
semantics, block_ffm.rs:265-288 / block_lr.rs:135-151) and every prediction and the final tables are compared with the
oracle; then at full concurrency against the oracle's logloss on full-shape streams (SURVEY 8d parity gates)."""
import numpy as np
import pytest

import fwumious_wabbit_b200 as fw
from fwumious_wabbit_b200 import _lib, synth
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle_run(w, recs, threads=1):
    n = recs.shape[0]
    ora = util.oracle_regressor(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    _, want = ora.hogwild(util.oracle_spec(w.mi), recs.reshape(-1), rec_off, threads, want_preds=True)
    return ora, want


def _table_report(name, got, want, atol):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bad = int(np.count_nonzero(d > atol))
    return f"{name}: max |d| {d.max():.3e}, entries over {atol:g}: {bad} of {d.size}", d.max(), bad


def _distinct_slots(w, recs):
    """Records in which no two FFM windows overlap and no two LR features share a cell.  The reference applies the updates of
    one example feature after feature (block_ffm.rs:269-287, block_lr.rs:140-150), so a slot that two features of ONE example
    share sees the first update's accumulator when the second steps; the fused kernels apply a record's updates concurrently
    (tests/test_gpu_parity.py::test_sequential_mode_bit_exact covers such records through the general kernel, which orders
    them).  c2: 0.2 % of the stream, c3: 2.7 %."""
    mi = w.mi
    F, k = len(mi.ffm_fields), mi.ffm_k
    kp = 1
    while kp < k:
        kp <<= 1
    ffm_mask = ((1 << mi.ffm_bit_precision) - 1) ^ (kp - 1)
    h = np.sort((recs[:, 3:3 + F] & np.uint32(ffm_mask)).astype(np.int64), axis=1)
    ok = np.all(np.diff(h, axis=1) >= F * k, axis=1)
    lr_mask = (1 << mi.bit_precision) - 1
    lr = np.concatenate([(recs[:, 3:3 + F] & np.uint32(lr_mask)).astype(np.int64),
                         np.full((recs.shape[0], 1), 11650396 & lr_mask, dtype=np.int64)], axis=1)
    lr.sort(axis=1)
    return ok & np.all(np.diff(lr, axis=1) > 0, axis=1)


@pytest.mark.parametrize("name,n,kernel", [("c2", 10_000, "fixed"), ("c3", 2_000, "fixed_cta")])
def test_fused_kernel_one_in_flight_bit_exact(name, n, kernel):
    """c2 through k_learn_fixed<16,4,1,LUT> and c3 through k_learn_rows<0> -- the kernels bench.py times -- with update = 1 and
    ONE record in flight: every prediction and the final weight / accumulator tables are BIT-EXACT with the sequential oracle
    over the whole stream.  In this mode the kernels sum the sigmoid's inputs in the reference's tape order; translate, gather,
    gradients, optimizer step and scatter are the code every other mode runs, so a wrong sign, pair or slot in any field's
    update shows up here within a few records.  Stream: records with pairwise distinct slots (see _distinct_slots)."""
    w = synth.workload(name)
    w.mi.hogwild_max_inflight = 1
    recs = w.records(int(n * 1.1) + 100)
    keep = _distinct_slots(w, recs)
    assert 0.9 < keep.mean() < 1.0
    recs = np.ascontiguousarray(recs[keep][:n])
    ora, want = _oracle_run(w, recs)
    re = fw.Regressor(w.mi)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    counts = re.path_counts()
    assert counts[kernel] > 0 and counts["general_examples"] == 0, counts   # the fused kernel did all of it
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.max(np.abs(got - want)))
    wts, acc = re.get_ffm()
    lr = re.get_lr_table()
    assert np.array_equal(wts.view(np.uint32), ora.ffm_weights.view(np.uint32))
    assert np.array_equal(acc.view(np.uint32), ora.ffm_acc.view(np.uint32))
    assert np.array_equal(lr.view(np.uint32), ora.lr_table.view(np.uint32))


@pytest.mark.parametrize("name,n,kernel", [("c2", 10_000, "fixed"), ("c3", 2_000, "fixed_cta")])
def test_fused_kernel_one_in_flight_whole_stream(name, n, kernel):
    """The same on the unfiltered stream: records whose rows share slots take their accumulators from atomics' return values
    (the wide kernel detects them while the rows are in flight; the narrow kernel always uses atomics; LR duplicates are
    ordered through an atomic too), i.e. in hardware order instead of the reference's feature order -- the same values up to
    the order of two additions.  Predictions stay within 1e-5 of the sequential oracle for EVERY record (measured: 0.0) and
    the tables within 2e-6 everywhere (measured: <= 1e-8); the touched cells are exactly the oracle's."""
    w = synth.workload(name)
    w.mi.hogwild_max_inflight = 1
    recs = w.records(n)
    ora, want = _oracle_run(w, recs)
    re = fw.Regressor(w.mi)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    assert re.path_counts()[kernel] > 0 and re.path_counts()["general_examples"] == 0
    d = np.abs(got - want)
    assert float(d.max()) <= TOL, float(d.max())
    wts, acc = re.get_ffm()
    msg, mx, bad = _table_report("ffm_w", wts, ora.ffm_weights, 2e-6)
    assert bad == 0, msg
    msg2, mx2, bad2 = _table_report("lr_w", re.get_lr_table()[:, 0], ora.lr_table[:, 0], 2e-6)
    assert bad2 == 0, msg2
    assert np.array_equal(acc != 0.0, ora.ffm_acc != 0.0)
    np.testing.assert_allclose(acc, ora.ffm_acc, rtol=1e-5, atol=1e-9)
    print(msg, ";", msg2, f"; max |dp| {float(d.max()):.2e}")


def test_fused_cta_phases_one_in_flight_match_oracle(monkeypatch):
    """The two-phase form of the block-per-record kernel (PHASE 1 forward -> head -> PHASE 2 update) on the full c5 shape,
    one example per sub-batch: predictions within 1e-5 of the sequential oracle over 300 examples, tables close."""
    w = synth.workload("c5")
    w.mi.hogwild_max_inflight = 1
    n = 300
    recs = w.records(n)
    ora, want = _oracle_run(w, recs)
    re = fw.Regressor(w.mi)
    util.sync_tables_from_oracle(re, util.oracle_regressor(w.mi))   # identical dense init on both sides
    re.set_examples_seen(0)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    assert re.path_counts()["fixed_cta"] > 0
    assert float(np.max(np.abs(got - want))) <= TOL
    msg, mx, bad = _table_report("ffm_w", re.get_ffm()[0], ora.ffm_weights, 2e-6)
    assert mx <= 2e-4 and bad <= 8, msg


# ------------------------------------------------------------------ full-shape Hogwild gates
@pytest.mark.parametrize("name,n,tol", [("c2", 10_000_000, 0.01), ("c3", 100_000, 0.01)])
def test_full_shape_hogwild_logloss_gate(name, n, tol):
    """Default mode (thousands of records in flight, concurrency ramp) on the full BASELINE shapes: progressive logloss
    within 1 % (relative) of the sequential oracle on the same stream -- c2 on 10^7 examples, c3 on 10^5.
    One fwgpu_learn_records call per stream, i.e. the chunking and ramp the bench uses."""
    w = synth.workload(name)
    recs = w.records(n)
    _, want = _oracle_run(w, recs)
    re = fw.Regressor(w.mi)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    counts = re.path_counts()
    assert counts["fixed" if name == "c2" else "fixed_cta"] > 0 and counts["general_examples"] == 0, counts
    labels = recs[:, 1].astype(np.float32)
    ll_o, ll_g = util.logloss(want, labels), util.logloss(got, labels)
    prior = util.logloss(np.full(n, labels.mean()), labels)
    assert ll_o < prior and ll_g < prior, (ll_o, ll_g, prior)
    assert abs(ll_g - ll_o) / ll_o < tol, (ll_g, ll_o)
    # the second half alone (a trained model under full concurrency)
    h = n // 2
    assert abs(util.logloss(got[h:], labels[h:]) - util.logloss(want[h:], labels[h:])) / util.logloss(want[h:], labels[h:]) < tol


def _small_c5(n_ns=10, k=4):
    w = synth.Workload("c5s", synth._mi(n_ns, ffm_k=k, ffm_bits=14, bits=14, lr=0.05, ffm_lr=0.02, ffm_init_acc=0.1),
                       synth.NS_LETTERS[:n_ns], [50] * 4 + [2000] * (n_ns - 4), "scaled-down c5")
    w.mi.nn_layers = [{"width": "32", "activation": "relu"}, {"width": "32", "activation": "relu"}]
    w.mi.nn_learning_rate, w.mi.nn_power_t, w.mi.nn_init_acc_gradient = 0.02, 0.5, 0.1
    return w


def _warm_pair(w, recs_warm):
    """Oracle and device regressor holding the same WARM state: the device trains on `recs_warm` in its default mode
    (concurrency ramp included) and the oracle takes over its tables.  A cold model with a thousand examples in flight
    is unstable on both sides -- that is what the ramp is for -- so the fixed-sub-batch comparisons start from here."""
    re = fw.Regressor(w.mi)
    re.learn_records(recs_warm.reshape(-1), n_examples=recs_warm.shape[0], update=True)
    re.set_examples_seen(1 << 40)       # trained: no ramp, full sub-batches from the next call on
    ora = util.oracle_regressor(w.mi)
    util.sync_oracle_from_gpu(ora, re)
    return ora, re


@pytest.mark.parametrize("shape", ["small", "c5"])
def test_head_umma_subbatch_matches_batched_oracle(shape, monkeypatch):
    """The tensor-core head path (sub-batches >= 512 rows: tcgen05 3xTF32 GEMMs, fused epilogues, summed-gradient optimizer
    step) against the oracle's batched-head semantics (fwo_learn_records_head_wave) at the SAME fixed sub-batch of 1024,
    both starting from the same warm model: every example of the first sub-batch sees the same snapshot on both sides ->
    per-example |dp| <= 2e-5 and the dense weights after the first step agree; over 8 sub-batches the sparse updates of a
    sub-batch land in a different order (and PHASE 2 re-reads rows other examples are updating), so those predictions
    are compared per example with a looser bound and through the logloss."""
    monkeypatch.setenv("FWGPU_HEAD_BATCH", "1024")
    w = _small_c5() if shape == "small" else synth.workload("c5")
    n, n0 = 1024, 200_000
    recs = w.records(n0 + 8 * n)
    warm, recs = recs[:n0], recs[n0:]
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    ora, re = _warm_pair(w, warm)
    want = ora.learn_head_wave(spec, recs[:n].reshape(-1), rec_off, n)
    got = re.learn_records(recs[:n].reshape(-1), n_examples=n, update=True)
    assert re.path_counts()["fixed_cta"] > 0
    err = float(np.max(np.abs(got - want)))
    assert err <= 2e-5, err
    for l in range(re.nn_layer_count()):
        gw, ga = re.get_nn(l)
        msg, mx, bad = _table_report(f"nn{l}_w", gw, ora.nn_weights(l), 5e-6)
        # LUT bucket edges: a weight whose accumulator lands on the other side of one takes a 6 % different step; measured 0.05 % of
        # a layer's weights (the warm state comes from a Hogwild run, so the count moves from run to run): bound 0.2 %
        assert mx <= 5e-4 and bad <= max(8, gw.size // 500), msg
        np.testing.assert_allclose(ga, ora.nn_acc(l), rtol=1e-3, atol=1e-7)
    # the whole stream: 8 sub-batches
    rec_off8 = np.arange(8 * n + 1, dtype=np.uint64) * w.record_len
    ora2, re2 = _warm_pair(w, warm)
    want8 = ora2.learn_head_wave(spec, recs.reshape(-1), rec_off8, n)
    got8 = re2.learn_records(recs.reshape(-1), n_examples=8 * n, update=True)
    labels = (recs[:, 1] == 1).astype(np.float32)
    ll_o, ll_g = util.logloss(want8, labels), util.logloss(got8, labels)
    d8 = float(np.max(np.abs(got8 - want8)))
    print(f"first sub-batch max |dp| {err:.2e}; 8 sub-batches max |dp| {d8:.2e}; logloss {ll_g:.5f} vs {ll_o:.5f}")
    assert ll_o < 0.7 and abs(ll_g - ll_o) / ll_o < 0.01, (ll_g, ll_o)
    assert d8 <= 1e-2, d8   # measured 6e-4 (small model) / 1.2e-3 (c5)


def test_c5_full_shape_hogwild_logloss_gate(monkeypatch):
    """Full c5 shape: 2*10^5 examples warm the model on the device (default mode), then 16000 examples run with the
    tensor-core head path forced (sub-batches of 1024): progressive logloss within 1 % of the batched oracle at the same
    sub-batch size and within 3 % of the sequential oracle, both continuing from the same warm state."""
    monkeypatch.setenv("FWGPU_HEAD_BATCH", "1024")
    w = synth.workload("c5")
    n0, n = 200_000, 16_000
    recs = w.records(n0 + n)
    warm, recs = recs[:n0], recs[n0:]
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    ora, re = _warm_pair(w, warm)
    want = ora.learn_head_wave(spec, recs.reshape(-1), rec_off, 1024)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    labels = (recs[:, 1] == 1).astype(np.float32)
    ll_o, ll_g = util.logloss(want, labels), util.logloss(got, labels)
    assert ll_o < 0.7 and abs(ll_g - ll_o) / ll_o < 0.01, (ll_g, ll_o)
    seq = util.oracle_regressor(w.mi)   # the sequential learner from the same warm state
    re0 = fw.Regressor(w.mi)
    re0.learn_records(warm.reshape(-1), n_examples=n0, update=True)
    util.sync_oracle_from_gpu(seq, re0)
    _, p_seq = seq.hogwild(spec, recs.reshape(-1), rec_off, 1, want_preds=True)
    ll_s = util.logloss(p_seq, labels)
    assert abs(ll_g - ll_s) / ll_s < 0.03, (ll_g, ll_s)


def test_translate_f32_namespace_bit_exact():
    """A namespace declared f32 (vwmap.rs:16-20, feature_buffer.rs:48-108): its dynamic pairs hold parsed floats and translate
    emits value 1.0 -- stream of records with f32 and categorical namespaces, translated on the device, u32-equal to the
    oracle's translate."""
    from fwumious_wabbit_b200 import ModelInstance, Optimizer
    from tests.test_gpu_parity import _random_records

    rng = np.random.default_rng(21)
    n_ns = 6
    mi = ModelInstance.new_empty()
    mi.bit_precision, mi.ffm_k, mi.ffm_bit_precision, mi.optimizer = 20, 4, 19, Optimizer.AdagradLUT
    mi.feature_combo_descs = [([0], 1.0), ([1], 2.0), ([2, 3], 1.0), ([4, 5, 0], 0.5), ([5], 1.0)]
    mi.ffm_fields, mi.num_namespaces = [[0], [1, 2], [3], [4, 5]], n_ns
    mi.ns_is_f32 = [0, 1, 0, 1, 0, 0]
    mi.max_ffm_per_example, mi.max_lr_per_example = 64, 128
    recs, offs = _random_records(rng, 500, n_ns, multi=True)
    re = fw.Regressor(mi)
    got = re.translate_records(recs, rec_off=offs)
    want = util.oracle_translate_batch(util.oracle_spec(mi), recs, rec_off=offs.astype(np.uint64))
    for key in ("lr_off", "lr_hash", "lr_combo", "ffm_off", "ffm_hash", "ffm_field"):
        assert np.array_equal(getattr(got, key), want[key]), key
    for key in ("labels", "importance", "lr_val", "ffm_val"):
        assert np.array_equal(getattr(got, key).view(np.uint32), want[key].view(np.uint32)), key
    # an f32 namespace really took the value-1.0 branch somewhere
    assert np.any(want["ffm_val"] == 1.0) and np.any(want["ffm_val"] != 1.0)
