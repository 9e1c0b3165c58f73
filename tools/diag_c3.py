import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fwumious_wabbit_b200 as fw
from fwumious_wabbit_b200 import synth
from tests import util
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
w = synth.workload("c3"); mi = w.mi; mi.hogwild_ramp_div = 0x7FFFFFFF
recs = w.records(n)
ora = util.oracle_regressor(mi); spec = util.oracle_spec(mi)
off = np.arange(n + 1, dtype=np.uint64) * w.record_len
_, want = ora.hogwild(spec, recs.reshape(-1), off, 1, want_preds=True)
re = fw.Regressor(mi)
got = re.learn_records(recs.reshape(-1), n_examples=n, update=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("FWGPU")})
print("pred mismatches", int(np.sum(got.view(np.uint32) != want.view(np.uint32))), "first", np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))[:5])
wts, acc = re.get_ffm()
dw = np.flatnonzero(wts.view(np.uint32) != ora.ffm_weights.view(np.uint32))
da = np.flatnonzero(acc.view(np.uint32) != ora.ffm_acc.view(np.uint32))
print("weight mismatches", dw.size, "acc mismatches", da.size)
for i in dw[:12]:
    print("  w idx", i, "gpu", repr(float(wts[i])), "ora", repr(float(ora.ffm_weights[i])), "diff", float(wts[i]) - float(ora.ffm_weights[i]), "acc gpu/ora", float(acc[i]), float(ora.ffm_acc[i]))
# which examples touch the first mismatching slot
if dw.size:
    s0 = int(dw[0]); F, k = 39, 8
    mask = ((1 << mi.ffm_bit_precision) - 1) ^ 7
    h = recs[:, 3:] & mask
    hit = np.argwhere((h <= s0) & (s0 < h + F * k))
    print("  examples touching slot", s0, ":", hit[:10].tolist(), "... count", len(hit))
    ex = int(hit[-1][0]); hs = np.sort(h[ex]); print("  last toucher", ex, "min gap between its rows", int(np.min(np.diff(hs))))
