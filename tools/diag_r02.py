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
for div, bias in ((32, -1),):
    os.environ["FWGPU_RAMP_DIV"] = str(div)
    if bias >= 0:
        os.environ["FWGPU_BIAS_PERIOD"] = str(bias)
    else:
        os.environ.pop("FWGPU_BIAS_PERIOD", None)
    # the env var is read once per process (static): run every variant in its own interpreter
    import subprocess
    code = (
        "import os,sys,numpy as np; sys.path.insert(0, %r)\n"
        "import fwumious_wabbit_b200 as fw\nfrom fwumious_wabbit_b200 import synth\nfrom tests import util\n"
        "w = synth.workload('c2'); n = %d; recs = w.records(n); want = np.load('/tmp/c2_want.npy'); labels = recs[:, 1].astype(np.float32)\n"
        "re = fw.Regressor(w.mi); got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)\n"
        "dec = n // 10\n"
        "gaps = [round(util.logloss(got[i*dec:(i+1)*dec], labels[i*dec:(i+1)*dec]) - util.logloss(want[i*dec:(i+1)*dec], labels[i*dec:(i+1)*dec]), 4) for i in range(10)]\n"
        "print('rel %%.4f' %% (abs(util.logloss(got, labels) - util.logloss(want, labels)) / util.logloss(want, labels)), 'decile gaps', gaps)\n"
    ) % (ROOT, n)
    np.save("/tmp/c2_want.npy", want)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print(f"c2 1e7 hogwild ramp_div {div} bias_period {bias}: {out.stdout.strip()} {out.stderr.strip()[-300:]}", flush=True)

# the transition to combined bias updates at 2^24 examples: 3e7 examples in one call, decile gaps around the switch
n3 = 30_000_000
recs3 = w.records(n3)
ora3 = util.oracle_regressor(w.mi)
_, want3 = ora3.hogwild(util.oracle_spec(w.mi), recs3.reshape(-1), np.arange(n3 + 1, dtype=np.uint64) * w.record_len, 1, want_preds=True)
os.environ.pop("FWGPU_RAMP_DIV", None); os.environ.pop("FWGPU_BIAS_PERIOD", None)
re3 = fw.Regressor(w.mi)
got3 = re3.learn_records(recs3.reshape(-1), n_examples=n3, update=True)
lab3 = recs3[:, 1].astype(np.float32)
d3 = n3 // 15
print("c2 3e7 default settings: rel %.4f; gaps per 2M examples" % (abs(util.logloss(got3, lab3) - util.logloss(want3, lab3)) / util.logloss(want3, lab3)),
      [round(util.logloss(got3[i * d3:(i + 1) * d3], lab3[i * d3:(i + 1) * d3]) - util.logloss(want3[i * d3:(i + 1) * d3], lab3[i * d3:(i + 1) * d3]), 4) for i in range(15)], flush=True)
