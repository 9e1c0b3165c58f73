"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libfwgpu.so) and is compared with the CPU oracle on the same inputs.

Tolerance: north_star asks per-example predictions within 1e-5; integer work (hashes, indices,
feature lists) must be bit-exact."""
import numpy as np
import pytest

import fwumious_wabbit_b200 as fw
from fwumious_wabbit_b200 import FeatureBuffer, HashAndValue, HashAndValueAndSeq, ModelInstance, Optimizer, _lib, synth
from oracle import fw_oracle as fo
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-5


def close(got, want, tol=TOL):
    assert abs(float(got) - float(want)) <= tol, (float(got), float(want))


def lr_vec(v, importance=1.0):
    return FeatureBuffer(label=0.0, example_importance=importance, lr_buffer=[HashAndValue(*t) for t in v])


def ffm_vec(v, lr=()):
    return FeatureBuffer(label=0.0, lr_buffer=[HashAndValue(*t) for t in lr], ffm_buffer=[HashAndValueAndSeq(*t) for t in v])


def new_mi(**kw):
    mi = ModelInstance.new_empty()
    for k, v in kw.items():
        assert hasattr(mi, k), k
        setattr(mi, k, v)
    return mi


# ------------------------------------------------------------------ the reference's own LR tests on the GPU
def test_learning_turned_off():  # regressor.rs:556-594
    re = fw.Regressor(new_mi(optimizer=Optimizer.AdagradLUT))
    assert re.learn(lr_vec([]), False) == 0.5
    assert re.learn(lr_vec([(1, 1.0, 0)]), False) == 0.5
    assert re.learn(lr_vec([(1, 1.0, 0), (2, 1.0, 0)]), False) == 0.5


@pytest.mark.parametrize("opt", [Optimizer.AdagradFlex, Optimizer.AdagradLUT, Optimizer.SGD])
def test_power_t_zero(opt):  # regressor.rs:597-626
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.0, optimizer=opt))
    v = lr_vec([(1, 1.0, 0)])
    close(re.learn(v, True), 0.5)
    close(re.learn(v, True), 0.48750263)
    close(re.learn(v, True), 0.47533244)


def test_double_same_feature():  # regressor.rs:629-656 -- duplicates applied in buffer order
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.0, optimizer=Optimizer.AdagradLUT))
    v = lr_vec([(1, 1.0, 0), (1, 2.0, 0)])
    close(re.learn(v, True), 0.5)
    close(re.learn(v, True), 0.38936076)
    close(re.learn(v, True), 0.30993468)


def test_power_t_half():  # regressor.rs:659-704
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.5, init_acc_gradient=0.0, optimizer=Optimizer.AdagradFlex))
    v = lr_vec([(1, 1.0, 0)])
    close(re.learn(v, True), 0.5)
    close(re.learn(v, True), 0.4750208)
    close(re.learn(v, True), 0.45788094)


def test_power_t_half_fastmath():  # regressor.rs:707-748
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.5, init_acc_gradient=0.0, optimizer=Optimizer.AdagradLUT))
    v = lr_vec([(1, 1.0, 0)])
    close(re.learn(v, True), 0.5)
    close(re.learn(v, True), 0.475734)


def test_power_t_half_two_features():  # regressor.rs:751-812
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.5, init_acc_gradient=0.0, optimizer=Optimizer.AdagradFlex))
    v2 = lr_vec([(1, 1.0, 0), (2, 1.0, 0)])
    close(re.learn(v2, True), 0.5)
    close(re.learn(v2, True), 0.45016602)
    close(re.learn(lr_vec([(1, 1.0, 0)]), True), 0.45836908)


def test_non_one_weight():  # regressor.rs:815-861
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.0, optimizer=Optimizer.AdagradLUT))
    v = lr_vec([(1, 2.0, 0)])
    close(re.learn(v, True), 0.5)
    close(re.learn(v, True), 0.45016602)
    close(re.learn(v, True), 0.40611085)


def test_example_importance():  # regressor.rs:864-884
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.0, optimizer=Optimizer.AdagradLUT))
    v = lr_vec([(1, 1.0, 0)], importance=0.5)
    close(re.learn(v, True), 0.5)
    close(re.learn(v, True), 0.49375027)
    close(re.learn(v, True), 0.4875807)


def test_zero_importance_does_not_update():  # regressor.rs:366-370
    re = fw.Regressor(new_mi(learning_rate=0.1, power_t=0.0, optimizer=Optimizer.AdagradLUT))
    v0 = lr_vec([(1, 1.0, 0)], importance=0.0)
    close(re.learn(v0, True), 0.5)
    close(re.learn(v0, True), 0.5)
    assert np.all(re.get_lr_table()[:, 0] == 0.0)


def ffm_mi(k=1, F=2, opt=Optimizer.AdagradFlex):
    return new_mi(learning_rate=0.1, power_t=0.0, bit_precision=18, ffm_k=k, ffm_bit_precision=18, ffm_power_t=0.0,
                  ffm_learning_rate=0.1, ffm_fields=[[] for _ in range(F)], optimizer=opt)


def ffm_ones(re, mi):
    n, _ = re.block_len(_lib.BLOCK_FFM)
    acc0 = mi.ffm_init_acc_gradient if mi.optimizer == Optimizer.AdagradFlex else 0.0
    re.set_ffm(np.ones(n, np.float32), None if mi.optimizer == Optimizer.SGD else np.full(n, acc0, np.float32))


def test_save_load_and_test_mode_ffm_values():  # persistence.rs:342-421 (LR + FFM k=1 + triangle)
    mi = ffm_mi()
    re = fw.Regressor(mi)
    ffm_ones(re, mi)
    v = ffm_vec([(1, 1.0, 0), (3000, 1.0, 0), (100, 2.0, 1)])
    close(re.learn(v, True), 0.9933072)
    close(re.learn(v, False), 0.9395168)
    close(re.predict(v), 0.9395168)
    # immutable (forward-only) regressor from the same weights: regressor.rs:471-534
    w, _ = re.get_ffm()
    fixed = fw.Regressor(mi, immutable=True)
    fixed.set_ffm(w)
    fixed.set_lr_table(re.get_lr_table()[:, 0].copy())
    close(fixed.predict(v), 0.9395168)
    assert fixed.get_name() == 'Regressor with optimizer "SGD"'
    with pytest.raises(_lib.FwgpuError):
        fixed.learn(v, True)


def test_hogwild_load_values():  # persistence.rs:437-555 (arithmetic part)
    mi = ffm_mi()
    r1, r2 = fw.Regressor(mi), fw.Regressor(mi)
    ffm_ones(r1, mi); ffm_ones(r2, mi)
    fb1 = ffm_vec([(1, 0.5, 0), (3000, 1.0, 0), (101, 2.0, 1)], lr=[(52, 0.5, 0), (2, 1.0, 0)])
    fb2 = ffm_vec([(1, 1.0, 0), (3000, 1.0, 0), (100, 2.0, 1)], lr=[(1, 1.0, 0), (2, 1.0, 0)])
    close(r1.learn(fb1, True), 0.97068775)
    close(r1.learn(fb1, False), 0.8922257)
    close(r1.predict(fb1), 0.8922257)
    close(r2.learn(fb2, True), 0.9933072)
    close(r2.learn(fb2, False), 0.92719215)
    close(r2.learn(fb1, False), 0.93763095)
    close(r1.learn(fb2, False), 0.98559695)


@pytest.mark.parametrize("k,F,opt", [(1, 2, Optimizer.AdagradLUT), (4, 2, Optimizer.AdagradFlex), (4, 3, Optimizer.AdagradLUT),
                                     (3, 3, Optimizer.AdagradLUT), (2, 5, Optimizer.SGD), (8, 4, Optimizer.AdagradLUT)])
def test_ffm_small_cases_vs_oracle(k, F, opt):
    """Hand-built multi-valued / missing-field examples like block_ffm.rs:1658-1944, against the oracle
    (full regressor graph, weights forced to 1.0 on both sides)."""
    mi = ffm_mi(k=k, F=F, opt=opt)
    re = fw.Regressor(mi)
    ora = util.oracle_regressor(mi)
    ora.ffm_weights[:] = 1.0
    ffm_ones(re, mi)
    kp = 1
    while kp < k:
        kp <<= 1
    h = lambda x: (x * kp) & ((1 << 18) - 1)
    cases = [
        [(h(1), 1.0, 0)],
        [(h(1), 1.0, 0), (h(100), 1.0, k)],
        [(h(1), 2.0, 0), (h(100), 2.0, k)],
        [(h(1), 1.0, 0), (h(3000), 1.0, 0), (h(100), 2.0, k)],
        [(h(5), 1.0, k)],
        [(h(7), 0.5, 0), (h(7), 0.5, k)],            # same row seen from two fields
        [(h(9), 1.5, (F - 1) * k)],
    ]
    for rep in range(3):
        for case in cases:
            fb_o = fo.feature_buffer(ffm=case, label=float(rep & 1))
            fb_g = FeatureBuffer(label=float(rep & 1), ffm_buffer=[HashAndValueAndSeq(*t) for t in case])
            close(re.predict(fb_g), ora.predict(fb_o))
            close(re.learn(fb_g, True), ora.learn(fb_o, True))
    w, acc = re.get_ffm()
    np.testing.assert_allclose(w, ora.ffm_weights, rtol=0, atol=2e-6)
    if acc is not None:
        np.testing.assert_allclose(acc, ora.ffm_acc, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------ init parity
def test_table_init_matches_reference_restatement():
    """merand48 init (block_ffm.rs:796-806) and LUT (optimizer.rs:121-144): bit-exact with the oracle."""
    mi = synth.workload("c2").mi
    re = fw.Regressor(mi)
    ora = util.oracle_regressor(mi)
    w, acc = re.get_ffm()
    assert np.array_equal(w.view(np.uint32), ora.ffm_weights.view(np.uint32))
    assert np.array_equal(acc, ora.ffm_acc)
    assert np.array_equal(re.get_lr_table(), ora.lr_table)
    for which in (0, 1):
        assert np.array_equal(re.lut(which).view(np.uint32), ora.lut(which).view(np.uint32))


# ------------------------------------------------------------------ translate: bit-exact
def _random_records(rng, n, n_ns, multi=True):
    recs, offs = [], [0]
    for _ in range(n):
        slots, dyn = [], []
        for _j in range(n_ns):
            r = rng.random()
            if r < 0.15:
                slots.append(0x80000000)
            elif r < 0.75 or not multi:
                slots.append(int(rng.integers(0, 1 << 31)))
            else:
                m = int(rng.integers(1, 4))
                start = 3 + n_ns + len(dyn)
                for _t in range(m):
                    dyn += [int(rng.integers(0, 1 << 31)), int(np.float32(rng.uniform(0.25, 3.0)).view(np.uint32))]
                slots.append(0x80000000 | (start << 16) | (3 + n_ns + len(dyn)))
        label = int(rng.integers(0, 2))
        imp = int(np.float32(rng.choice([1.0, 1.0, 0.5, 2.0])).view(np.uint32))
        rec = [3 + n_ns + len(dyn), label, imp] + slots + dyn
        recs += rec
        offs.append(len(recs))
    return np.array(recs, dtype=np.uint32), np.array(offs, dtype=np.uint32)


@pytest.mark.parametrize("ffm_k", [0, 3, 4, 8])
def test_translate_bit_exact(ffm_k):
    rng = np.random.default_rng(7 + ffm_k)
    n_ns = 6
    mi = new_mi(bit_precision=20, ffm_k=ffm_k, ffm_bit_precision=19, optimizer=Optimizer.AdagradLUT,
                feature_combo_descs=[([0], 1.0), ([1], 2.0), ([2, 3], 1.0), ([4, 5, 0], 0.5), ([5], 1.0)],
                ffm_fields=[[0], [1, 2], [3], [4, 5]] if ffm_k else [], num_namespaces=n_ns,
                max_ffm_per_example=64, max_lr_per_example=128)
    recs, offs = _random_records(rng, 500, n_ns)
    re = fw.Regressor(mi)
    got = re.translate_records(recs, rec_off=offs)
    want = util.oracle_translate_batch(util.oracle_spec(mi), recs, rec_off=offs.astype(np.uint64))
    for key in ("lr_off", "lr_hash", "lr_combo", "ffm_off", "ffm_hash", "ffm_field"):
        assert np.array_equal(getattr(got, key), want[key]), key
    for key in ("labels", "importance", "lr_val", "ffm_val"):
        assert np.array_equal(getattr(got, key).view(np.uint32), want[key].view(np.uint32)), key


def test_translate_fixed_width_matches_synth():
    w = synth.workload("c2")
    recs = w.records(2000)
    re = fw.Regressor(w.mi)
    got = re.translate_records(recs.reshape(-1), n_examples=2000)
    want = util.oracle_translate_batch(util.oracle_spec(w.mi), recs, fixed_len=w.record_len)
    for key in ("lr_off", "lr_hash", "lr_combo", "ffm_off", "ffm_hash", "ffm_field"):
        assert np.array_equal(getattr(got, key), want[key]), key


# ------------------------------------------------------------------ predict / learn parity on random data
CONFIGS = {
    "lr_only": dict(ffm_k=0, F=0, combos=6),
    "k4_f8": dict(ffm_k=4, F=8, combos=8),
    "k8_f39": dict(ffm_k=8, F=39, combos=39),
    "k2_f5": dict(ffm_k=2, F=5, combos=3),
    "k1_f3": dict(ffm_k=1, F=3, combos=2),
    "k3_f7": dict(ffm_k=3, F=7, combos=4),
    "k10_f6": dict(ffm_k=10, F=6, combos=4),
    "k16_f4": dict(ffm_k=16, F=4, combos=2),
}


def cfg_mi(name, opt=Optimizer.AdagradLUT, bits=14, ffm_bits=14):
    c = CONFIGS[name]
    return new_mi(learning_rate=0.1, power_t=0.5, ffm_learning_rate=0.05, ffm_power_t=0.5, bit_precision=bits,
                  ffm_k=c["ffm_k"], ffm_bit_precision=ffm_bits, optimizer=opt,
                  feature_combo_descs=[([j], 1.0) for j in range(c["combos"])], ffm_fields=[[j] for j in range(c["F"])],
                  num_namespaces=max(c["combos"], c["F"]))


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("multi", [False, True])
def test_predict_parity_random_tables(name, multi):
    rng = np.random.default_rng(hash(name) % 1000 + multi)
    mi = cfg_mi(name)
    ora = util.oracle_regressor(mi)
    ora.lr_table[:, 0] = rng.normal(0, 0.2, ora.lr_table.shape[0]).astype(np.float32)
    if mi.ffm_k:
        ora.ffm_weights[:] = rng.normal(0, 0.3, ora.ffm_weights.shape[0]).astype(np.float32)
    re = fw.Regressor(mi)
    util.sync_tables_from_oracle(re, ora)
    d = util.random_csr(rng, 300, mi, multi_valued=multi, empty_prob=0.2 if multi else 0.0, value_one=not multi)
    want = ora.learn_batch(d, update=False)
    got = re.predict_batch(util.csr_from_dict(d))
    assert np.max(np.abs(got - want)) <= TOL, np.max(np.abs(got - want))
    # predict is read-only and idempotent
    assert np.array_equal(got, re.predict_batch(util.csr_from_dict(d)))
    assert np.array_equal(re.get_lr_table(), ora.lr_table)


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("opt", [Optimizer.AdagradLUT, Optimizer.AdagradFlex, Optimizer.SGD])
def test_learn_batch1_parity(name, opt):
    """batch = 1 is the sequential reference semantics: per-example prediction within 1e-5 and the
    tables equal afterwards."""
    if opt != Optimizer.AdagradLUT and name not in ("k4_f8", "lr_only", "k3_f7"):
        pytest.skip("optimizer variants covered on three shapes")
    rng = np.random.default_rng(11)
    mi = cfg_mi(name, opt=opt)
    ora = util.oracle_regressor(mi)
    re = fw.Regressor(mi)
    util.sync_tables_from_oracle(re, ora)
    n = 150 if name == "k8_f39" else 300
    d = util.random_csr(rng, n, mi, multi_valued=(name in ("k2_f5", "k3_f7")), empty_prob=0.1, value_one=False)
    want = ora.learn_batch(d, update=True)
    batch = util.csr_from_dict(d)
    got = np.array([re.learn_batch(batch.slice(i, i + 1), True)[0] for i in range(n)], dtype=np.float32)
    exact = opt != Optimizer.AdagradFlex  # Flex calls powf: CUDA's and glibc's differ in the last bits, everything else is bit-exact
    if exact:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.max(np.abs(got - want)))
    else:
        assert np.max(np.abs(got - want)) <= TOL, np.max(np.abs(got - want))
    t = re.get_lr_table()
    if mi.ffm_k:
        w, acc = re.get_ffm()
    if exact:
        assert np.array_equal(t[:, 0].view(np.uint32), ora.lr_table[:, 0].view(np.uint32))
        if mi.ffm_k:
            assert np.array_equal(w.view(np.uint32), ora.ffm_weights.view(np.uint32))
            if acc is not None:
                assert np.array_equal(acc.view(np.uint32), ora.ffm_acc.view(np.uint32))
    else:
        np.testing.assert_allclose(t[:, 0], ora.lr_table[:, 0], rtol=0, atol=5e-6)
        if mi.ffm_k:
            np.testing.assert_allclose(w, ora.ffm_weights, rtol=0, atol=5e-6)
            np.testing.assert_allclose(acc, ora.ffm_acc, rtol=1e-4, atol=1e-7)


SEQUENTIAL = 0x7FFFFFFF  # hogwild_ramp_div so large that one example is in flight for the whole run


@pytest.mark.parametrize("name,n", [("c2", 20_000), ("c3", 1_500), ("c1", 20_000)])
def test_sequential_mode_bit_exact(name, n):
    """One example in flight (the reference's default single-threaded loop, main.rs:213-258), a whole
    stream in a single call: per-example predictions and the final tables are BIT-EXACT with the oracle on
    the BASELINE shapes (single-valued namespaces, value 1.0), AdagradLUT included."""
    w = synth.workload(name)
    mi = w.mi
    mi.hogwild_ramp_div = SEQUENTIAL
    recs = w.records(n)
    ora = util.oracle_regressor(mi)
    spec = util.oracle_spec(mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    # the training-order forward is what the reference returns from learn(update=true) (regressor.rs:356-379)
    _, want = ora.hogwild(spec, recs.reshape(-1), rec_off, 1, want_preds=True)
    re = fw.Regressor(mi)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.max(np.abs(got - want)))
    assert np.array_equal(re.get_lr_table().view(np.uint32), ora.lr_table.view(np.uint32))
    if mi.ffm_k:
        wts, acc = re.get_ffm()
        assert np.array_equal(wts.view(np.uint32), ora.ffm_weights.view(np.uint32))
        assert np.array_equal(acc.view(np.uint32), ora.ffm_acc.view(np.uint32))


def test_learn_large_batch_progressive_logloss():
    """Hogwild-on-device with a large batch: progressive logloss within 1 % (relative) of the
    sequential oracle on the same stream (BASELINE.md parity gate, batch size stated: 8192)."""
    w = synth.workload("c2")
    n = 200_000
    recs = w.records(n)
    ora = util.oracle_regressor(w.mi)
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    _, want = ora.hogwild(spec, recs.reshape(-1), rec_off, 1, want_preds=True)
    re = fw.Regressor(w.mi)
    got = np.empty(n, np.float32)
    B = 8192
    for a in range(0, n, B):
        b = min(n, a + B)
        re.learn_records(recs[a:b].reshape(-1), n_examples=b - a, update=True, out=got[a:b])
    labels = recs[:, 1].astype(np.float32)
    ll_o, ll_g = util.logloss(want, labels), util.logloss(got, labels)
    assert ll_g < 0.69 and ll_o < 0.69
    assert abs(ll_g - ll_o) / ll_o < 0.01, (ll_g, ll_o)
    # second half (model has learned something): both well below chance
    h = n // 2
    assert util.logloss(got[h:], labels[h:]) < util.logloss(np.full(n - h, labels.mean()), labels[h:])


def test_concurrency_ramp_bookkeeping():
    """The cold-start ramp: examples_seen advances with learned examples only, importing optimizer state
    marks the model as trained, and a cold model trained in one big call still learns (no overshoot)."""
    w = synth.workload("c2")
    n = 300_000
    recs = w.records(n)
    re = fw.Regressor(w.mi)
    assert re.examples_seen() == 0
    re.learn_records(recs[:1000].reshape(-1), n_examples=1000, update=False)
    assert re.examples_seen() == 0
    p = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    assert re.examples_seen() == n
    labels = recs[:, 1].astype(np.float32)
    prior = util.logloss(np.full(n, labels.mean()), labels)
    assert util.logloss(p, labels) < prior  # progressive logloss of a cold model beats the prior even in one call
    re2 = fw.Regressor(w.mi)
    re2.import_block(_lib.BLOCK_FFM, re.export_block(_lib.BLOCK_FFM), True)
    assert re2.examples_seen() >= 1 << 40


def test_records_path_equals_csr_path():
    """learn_records (device translate) and learn_batch (host CSR) are the same computation."""
    w = synth.workload("c2")
    n = 4096
    recs = w.records(n)
    d = util.oracle_translate_batch(util.oracle_spec(w.mi), recs, fixed_len=w.record_len)
    r1, r2 = fw.Regressor(w.mi), fw.Regressor(w.mi)
    p1 = r1.learn_records(recs.reshape(-1), n_examples=n, update=False)  # fused fast kernel (warp per record)
    p2 = r2.predict_batch(util.csr_from_dict(d))                           # general kernel on translated features
    assert np.max(np.abs(p1 - p2)) <= 1e-6
    want = util.oracle_regressor(w.mi).learn_batch(d, update=False)
    assert np.max(np.abs(p1 - want)) <= TOL and np.max(np.abs(p2 - want)) <= TOL


@pytest.mark.parametrize("k", [4, 8])
def test_fast_path_with_leftovers(k):
    """Records whose namespaces hold several features / weights do not fit the fused kernel: they are listed
    and go through translate + the general kernel.  Every prediction still matches the oracle, and training on
    the mixed stream moves the same weights."""
    rng = np.random.default_rng(5 + k)
    n_ns = 6
    mi = new_mi(learning_rate=0.1, power_t=0.5, ffm_learning_rate=0.05, ffm_power_t=0.5, bit_precision=18, ffm_k=k,
                ffm_bit_precision=18, optimizer=Optimizer.AdagradLUT,
                feature_combo_descs=[([j], 1.0) for j in range(n_ns)] + [([0, 1], 1.0), ([2, 3, 4], 0.5)],
                ffm_fields=[[j] for j in range(n_ns)], num_namespaces=n_ns, max_ffm_per_example=32, max_lr_per_example=64)
    recs, offs = _random_records(rng, 3000, n_ns, multi=True)
    ora = util.oracle_regressor(mi)
    ora.lr_table[:, 0] = rng.normal(0, 0.2, ora.lr_table.shape[0]).astype(np.float32)
    ora.ffm_weights[:] = rng.normal(0, 0.3, ora.ffm_weights.shape[0]).astype(np.float32)
    re = fw.Regressor(mi)
    util.sync_tables_from_oracle(re, ora)
    re.set_examples_seen(0)
    d = util.oracle_translate_batch(util.oracle_spec(mi), recs, rec_off=offs.astype(np.uint64))
    want = ora.learn_batch(d, update=False)
    got = re.learn_records(recs, rec_off=offs, update=False)
    assert np.max(np.abs(got - want)) <= TOL
    # now learn on it: every example must have been processed exactly once (LR accumulators count the hits)
    re.learn_records(recs, rec_off=offs, update=True)
    acc = re.get_lr_table()[:, 1]
    const_cell = 11650396 & ((1 << 18) - 1)
    ora.learn_batch(d, update=True)
    assert acc[const_cell] > 0
    touched_gpu = np.flatnonzero(re.get_lr_table()[:, 1] != ora.lr_table[:, 1] * 0)
    touched_ora = np.flatnonzero(ora.lr_table[:, 1] != 0)
    assert np.array_equal(touched_gpu, touched_ora)


FUSED_SHAPES = {
    # name: (fields, k, combos) -> the k_learn_fixed instantiation the shape selects (fwgpu.cu launch_fixed)
    "g16_nch4": (8, 4, [[j] for j in range(8)]),                       # 64 chunks: two records per warp, 4 chunks per lane (c2)
    "g16_nch2": (4, 8, [[j] for j in range(4)]),                       # 32 chunks
    "g16_nch3": (6, 4, [[j] for j in range(6)] + [[0, 1], [2, 3, 4]]),  # 36 chunks, interactions
    "g32_nch4": (8, 8, [[j] for j in range(8)]),                       # 128 chunks: one record per warp
    "g32_nch4_f10": (10, 4, [[j] for j in range(10)]),                 # 100 chunks, F <= 16 but too many chunks for two records
    "g32_nlr2": (5, 4, [[j] for j in range(5)] + [[a, b] for a in range(5) for b in range(a + 1, 5)] +
                 [[a, b, c] for a in range(5) for b in range(a + 1, 5) for c in range(b + 1, 5)] +
                 [[0, 1, 2, 3], [1, 2, 3, 4], [0, 2, 3, 4], [0, 1, 3, 4], [0, 1, 2, 4], [0, 1, 2, 3, 4], [4, 3], [3, 1]]),  # 34 LR entries > 32
}


@pytest.mark.parametrize("name", list(FUSED_SHAPES))
def test_fused_kernel_shapes(name):
    """Every instantiation family of the warp-per-record fused kernel (16- and 32-lane groups, 1..4 chunks per lane, one or
    two LR entries per lane) on raw single-valued records with empty slots: predictions on random tables match the oracle
    (<= 1e-5), and training from a cold model with 4 records in flight tracks the sequential oracle (progressive logloss within 3 %)."""
    F, k, combos = FUSED_SHAPES[name]
    rng = np.random.default_rng(len(name) * 7 + F + k)
    mi = new_mi(learning_rate=0.1, power_t=0.5, ffm_learning_rate=0.05, ffm_power_t=0.5, bit_precision=16, ffm_k=k,
                ffm_bit_precision=15, optimizer=Optimizer.AdagradLUT, feature_combo_descs=[(c, 1.0) for c in combos],
                ffm_fields=[[j] for j in range(F)], num_namespaces=F)
    n = 6000
    # ids from a small vocabulary so that rows repeat and there is something to learn; the label depends on two of them
    recs = np.empty((n, 3 + F), dtype=np.uint32)
    ids = rng.integers(0, 50, size=(n, F))
    recs[:, 3:] = (ids * 2654435761 + np.arange(F) * 40503) % (1 << 31)
    recs[:, 3:][rng.random((n, F)) < 0.1] = 0x80000000  # empty slots
    recs[:, 0] = 3 + F
    recs[:, 1] = ((ids[:, 0] + ids[:, F - 1]) % 3 == 0).astype(np.uint32)
    recs[:, 2] = np.float32(1.0).view(np.uint32)
    ora = util.oracle_regressor(mi)
    ora.lr_table[:, 0] = rng.normal(0, 0.2, ora.lr_table.shape[0]).astype(np.float32)
    ora.ffm_weights[:] = rng.normal(0, 0.3, ora.ffm_weights.shape[0]).astype(np.float32)
    re = fw.Regressor(mi)
    util.sync_tables_from_oracle(re, ora)
    spec = util.oracle_spec(mi)
    d = util.oracle_translate_batch(spec, recs, fixed_len=3 + F)
    want = ora.learn_batch(d, update=False)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=False)
    assert np.max(np.abs(got - want)) <= TOL, np.max(np.abs(got - want))
    # training from a cold model, at most 4 records in flight: the stream's vocabulary is tiny (every row is hot), so
    # full Hogwild concurrency would measure staleness, not the kernel's update arithmetic
    ora2 = util.oracle_regressor(mi)
    mi.hogwild_max_inflight = 4
    re2 = fw.Regressor(mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * (3 + F)
    _, want_l = ora2.hogwild(spec, recs.reshape(-1), rec_off, 1, want_preds=True)
    got_l = re2.learn_records(recs.reshape(-1), n_examples=n, update=True)
    labels = recs[:, 1].astype(np.float32)
    ll_o, ll_g = util.logloss(want_l, labels), util.logloss(got_l, labels)
    # measured: below 1 % for five of the shapes, 1.7 % for 4 fields x k=8; the oracle's own emulation of 4 records in flight
    # (learn_wave) moves the logloss of these tiny-vocabulary streams by 1-4 %, so this is update order, not arithmetic
    assert abs(ll_g - ll_o) / ll_o < 0.03, (ll_g, ll_o)
    assert util.logloss(got_l[n // 2:], labels[n // 2:]) < util.logloss(np.full(n - n // 2, labels.mean()), labels[n // 2:])


def test_dataset_resident_path_and_determinism_of_predict():
    w = synth.workload("c2")
    n = 50_000
    recs = w.records(n)
    re = fw.Regressor(w.mi)
    ds = re.upload_dataset(recs.reshape(-1), n_examples=n)
    p = np.empty(n, np.float32)
    re.learn_dataset(ds, 0, n, update=True, out=p)
    q1, q2 = np.empty(n, np.float32), np.empty(n, np.float32)
    re.learn_dataset(ds, 0, n, update=False, out=q1)
    re.learn_dataset(ds, 0, n, update=False, out=q2)
    assert np.array_equal(q1, q2)
    labels = recs[:, 1].astype(np.float32)
    assert util.logloss(q1, labels) < util.logloss(p, labels)  # a trained model scores its training set better
    ds.free()


def test_export_import_roundtrip_and_byte_layout():
    mi = cfg_mi("k4_f8")
    re = fw.Regressor(mi)
    rng = np.random.default_rng(3)
    d = util.random_csr(rng, 2000, mi)
    re.learn_batch(util.csr_from_dict(d), True)
    n_lr, b_lr = re.block_len(_lib.BLOCK_LR)
    n_f, b_f = re.block_len(_lib.BLOCK_FFM)
    assert n_lr == 1 << mi.bit_precision and b_lr == n_lr * 8          # {f32 w, f32 acc} (block_helpers.rs:23-28)
    assert n_f == (1 << mi.ffm_bit_precision) + 8 * 4 and b_f == n_f * 8  # w x L then acc x L (block_ffm.rs:835-848)
    lr, ffm = re.export_block(_lib.BLOCK_LR), re.export_block(_lib.BLOCK_FFM)
    re2 = fw.Regressor(mi)
    re2.import_block(_lib.BLOCK_LR, lr, True)
    re2.import_block(_lib.BLOCK_FFM, ffm, True)
    b = util.csr_from_dict(d)
    assert np.array_equal(re.predict_batch(b), re2.predict_batch(b))
    assert np.array_equal(re2.export_block(_lib.BLOCK_FFM), ffm)


def test_edge_cases():
    mi = cfg_mi("k4_f8")
    re = fw.Regressor(mi)
    ora = util.oracle_regressor(mi)
    util.sync_tables_from_oracle(re, ora)
    # empty batch, empty example, NaN value -> p = 0.5 / no update (block_loss_functions.rs:125-133)
    empty = util.random_csr(np.random.default_rng(0), 0, mi)
    assert re.learn_batch(util.csr_from_dict(empty), True).shape == (0,)
    fb = FeatureBuffer(label=1.0)
    close(re.learn(fb, True), 0.5)
    before = re.get_lr_table().copy()
    nanfb = FeatureBuffer(label=1.0, lr_buffer=[HashAndValue(5, float("nan"), 0)])
    ora.lr_table[5, 0] = 0.25
    util.sync_tables_from_oracle(re, ora)
    close(re.learn(nanfb, True), 0.5)
    after = re.get_lr_table()
    assert after[5, 0] == np.float32(0.25) and np.array_equal(after[:5], before[:5])
    # saturation: |wsum| > 50 -> clamped probability, zero gradient
    ora.lr_table[7, 0] = 100.0
    util.sync_tables_from_oracle(re, ora)
    big = FeatureBuffer(label=0.0, lr_buffer=[HashAndValue(7, 1.0, 0)])
    close(re.learn(big, True), fo.Regressor(learning_rate=0.1).predict(fo.feature_buffer()) * 0 + 1.0, tol=1e-6)
    assert re.get_lr_table()[7, 0] == np.float32(100.0)
    # last row of the table: the window spills into the F*k tail (block_ffm.rs:92-94)
    last = ((1 << mi.ffm_bit_precision) - 1) & ~3
    fbs_o = fo.feature_buffer(ffm=[(last, 1.0, 0), (last, 1.0, 4)], label=1.0)
    fbs_g = FeatureBuffer(label=1.0, ffm_buffer=[HashAndValueAndSeq(last, 1.0, 0), HashAndValueAndSeq(last, 1.0, 4)])
    close(re.learn(fbs_g, True), ora.learn(fbs_o, True))
    w, _ = re.get_ffm()
    np.testing.assert_allclose(w[last:], ora.ffm_weights[last:], rtol=0, atol=2e-6)


def test_too_many_features_is_reported():
    """An example with more FFM features than max_ffm_per_example cannot be staged: the call reports
    FWGPU_ERR_TOO_LARGE at sync (its prediction is NaN, nothing is learned from it) instead of corrupting memory."""
    rng = np.random.default_rng(9)
    n_ns = 6
    mi = new_mi(bit_precision=18, ffm_k=4, ffm_bit_precision=18, optimizer=Optimizer.AdagradLUT,
                feature_combo_descs=[([j], 1.0) for j in range(n_ns)], ffm_fields=[[j] for j in range(n_ns)],
                num_namespaces=n_ns, max_ffm_per_example=6, max_lr_per_example=64)
    recs, offs = _random_records(rng, 200, n_ns, multi=True)  # multi-valued namespaces: up to 18 features per example
    re = fw.Regressor(mi)
    with pytest.raises(_lib.FwgpuError) as ei:
        re.learn_records(recs, rec_off=offs, update=True)
    assert ei.value.status == _lib.ERR_TOO_LARGE
    # the ctx stays usable afterwards
    simple = np.array([3 + n_ns, 1, 0x3F800000] + [7 * (j + 1) for j in range(n_ns)], dtype=np.uint32)
    assert re.learn_records(simple, n_examples=1, update=False).shape == (1,)


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_full_size_config_properties(name):
    """BASELINE.json shapes at full table size: parity on a prefix + size-independent properties."""
    w = synth.workload(name)
    n = 20_000 if name == "c3" else 100_000
    recs = w.records(n)
    re = fw.Regressor(w.mi)
    ora = util.oracle_regressor(w.mi)
    spec = util.oracle_spec(w.mi)
    m = 2000
    rec_off = np.arange(m + 1, dtype=np.uint64) * w.record_len
    # predict-only on identical (freshly initialised, merand48) tables
    _, _ = None, None
    want = np.array([ora.predict(_fb(spec, recs[i])) for i in range(m)], dtype=np.float32)
    got = re.learn_records(recs[:m].reshape(-1), n_examples=m, update=False)
    assert np.max(np.abs(got - want)) <= TOL
    # train on everything, then: predictions finite in (0,1), logloss beats the prior, predict deterministic
    p = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    assert np.all(np.isfinite(p)) and p.min() > 0 and p.max() < 1
    q1 = re.learn_records(recs.reshape(-1), n_examples=n, update=False)
    q2 = re.learn_records(recs.reshape(-1), n_examples=n, update=False)
    assert np.array_equal(q1, q2)
    labels = recs[:, 1].astype(np.float32)
    assert util.logloss(q1, labels) < util.logloss(np.full(n, labels.mean()), labels)
    # accumulators only grow, weights stay finite
    wts, acc = re.get_ffm()
    assert np.all(np.isfinite(wts)) and np.all(acc >= 0)


def _fb(spec, rec):
    lab, imp, lr, ffm = spec.translate(rec)
    return fo.feature_buffer(lr=lr, ffm=ffm, label=lab, importance=imp)
