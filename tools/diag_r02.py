"""Round-2 diagnostics (GPU box): head sequential-mode drift of the rows kernel on the small model, c2 Hogwild gate spread."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fwumious_wabbit_b200 as fw  # noqa: E402
from fwumious_wabbit_b200 import synth  # noqa: E402
from tests import util  # noqa: E402


def small_c5(bits=14, n_ns=10, k=4, head=True):
    w = synth.Workload("c5s", synth._mi(n_ns, ffm_k=k, ffm_bits=bits, bits=bits, lr=0.05, ffm_lr=0.02, ffm_init_acc=0.1),
                       synth.NS_LETTERS[:n_ns], [50] * 4 + [2000] * (n_ns - 4), "scaled-down c5")
    if head:
        w.mi.nn_layers = [{"width": "32", "activation": "relu"}, {"width": "32", "activation": "relu"}]
        w.mi.nn_learning_rate, w.mi.nn_power_t, w.mi.nn_init_acc_gradient = 0.02, 0.5, 0.1
    return w


def seq_run(w, n, rows, mode):
    os.environ["FWGPU_ROWS"] = rows
    if mode == "ramp":
        w.mi.hogwild_ramp_div = 0x7FFFFFFF
    else:
        w.mi.hogwild_max_inflight = 1
    recs = w.records(n)
    ora = util.oracle_regressor(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    _, want = ora.hogwild(util.oracle_spec(w.mi), recs.reshape(-1), rec_off, 1, want_preds=True)
    re = fw.Regressor(w.mi)
    util.sync_tables_from_oracle(re, util.oracle_regressor(w.mi))
    re.set_examples_seen(0)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    d = np.abs(got - want)
    first_bad = int(np.argmax(d > 1e-5)) if np.any(d > 1e-5) else -1
    return [float(d[:m].max()) for m in (100, 300, 1000, n)], first_bad, re.path_counts()


for head in (True, False):
    for bits in (14,):
        for k in (4, 8):
            for rows in ("1",):
                for mode in ("ramp", "inflight1"):
                    if not head and mode == "ramp":
                        continue
                    try:
                        r = seq_run(small_c5(bits=bits, k=k, head=head), 3000, rows, mode)
                    except Exception as e:  # noqa: BLE001
                        r = repr(e)
                    print(f"head={head} bits={bits} k={k} rows_kernel={rows} mode={mode}: max|dp| at 100/300/1000/3000 = {r}", flush=True)
os.environ.pop("FWGPU_ROWS", None)

# c2 Hogwild gate: spread over repeated runs
w = synth.workload("c2")
n = 10_000_000
recs = w.records(n)
ora = util.oracle_regressor(w.mi)
rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
_, want = ora.hogwild(util.oracle_spec(w.mi), recs.reshape(-1), rec_off, 1, want_preds=True)
labels = recs[:, 1].astype(np.float32)
ll_o = util.logloss(want, labels)
dec = n // 10
for div in (32, 64, 128, 256, 1024):
    os.environ["FWGPU_RAMP_DIV"] = str(div)
    re = fw.Regressor(w.mi)
    got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
    ll_g = util.logloss(got, labels)
    gaps = [util.logloss(got[i * dec:(i + 1) * dec], labels[i * dec:(i + 1) * dec]) - util.logloss(want[i * dec:(i + 1) * dec], labels[i * dec:(i + 1) * dec]) for i in range(10)]
    first = [util.logloss(got[a:b], labels[a:b]) - util.logloss(want[a:b], labels[a:b]) for a, b in ((0, 10_000), (10_000, 100_000), (100_000, 300_000), (300_000, 1_000_000))]
    print(f"c2 1e7 hogwild ramp_div {div}: gpu {ll_g:.5f} oracle {ll_o:.5f} rel {abs(ll_g - ll_o) / ll_o:.4f}; decile gaps {[round(g, 4) for g in gaps]}; first 1e4/1e5/3e5/1e6 gaps {[round(g, 4) for g in first]}", flush=True)
    re.close()
os.environ.pop("FWGPU_RAMP_DIV", None)
