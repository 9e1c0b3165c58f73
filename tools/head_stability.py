#!/usr/bin/env python
"""Experiment: progressive logloss of head models under different sub-batch policies vs the sequential oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwumious_wabbit_b200 as fw
from fwumious_wabbit_b200 import synth
from tests import util
from tests.test_gpu_head import small_c5


def ll_windows(p, y, k=4):
    n = len(p)
    return [round(util.logloss(p[i * n // k:(i + 1) * n // k], y[i * n // k:(i + 1) * n // k]), 4) for i in range(k)]


def run(wname, n, n_ora, configs):
    mk = (lambda: small_c5()) if wname == "c5s" else (lambda: synth.workload(wname))
    w = mk()
    recs = w.records(n)
    y = (recs[:, 1] == 1).astype(np.float32)
    t = time.time()
    ora = util.oracle_regressor(w.mi)
    off = np.arange(n_ora + 1, dtype=np.uint64) * w.record_len
    _, want = ora.hogwild(util.oracle_spec(w.mi), recs[:n_ora].reshape(-1), off, 1, want_preds=True)
    print(f"{wname} oracle sequential first {n_ora}: {ll_windows(want, y[:n_ora])}  ({time.time() - t:.0f} s)  prior {util.logloss(np.full(n, y.mean()), y):.4f}", flush=True)
    for hb, mul in configs:
        os.environ["FWGPU_HEAD_BATCH"] = str(hb); os.environ["FWGPU_HEAD_RAMP_MUL"] = str(mul)
        re = fw.Regressor(mk().mi)
        t = time.time()
        got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
        dt = time.time() - t
        print(f"  head_batch {hb:5d} ramp_mul {mul:4}: first-{n_ora} {ll_windows(got[:n_ora], y[:n_ora])}  all {ll_windows(got, y)}  {n / dt / 1e6:.2f} M ex/s", flush=True)
        re.close()


if __name__ == "__main__":
    cfgs = [(256, 0), (1024, 0), (4096, 0), (1024, 2), (4096, 2), (4096, 8), (16384, 2)]
    run("c5s", 1_000_000, 200_000, cfgs)
    run("c5", 1_000_000, 40_000, cfgs)
