"""Hash-range-sharded table over two GPUs (BASELINE config 4's mechanism; fwgpu_create_sharded): one process per GPU,
tables mapped into one virtual range on both, remote rows gathered / updated over NVLink inside the learn kernels.
Needs two GPUs with peer access; skipped otherwise (the driver's single-GPU box skips it; run with gpurun --gpus 2)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2] + ([8] if _n_gpus() >= 8 else []))
def test_sharded_table(world):
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "rv")
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "shard_worker.py"), str(r), str(world), prefix, d],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
        outs = [p.communicate(timeout=900)[0] for p in procs]
        for r, p in enumerate(procs):
            assert p.returncode == 0, f"rank {r} failed:\n{outs[r][-3000:]}"
        res = [np.load(os.path.join(d, f"rank{r}.npz")) for r in range(world)]
    # A. layout: rank r holds half of the FFM table; predictions through the sharded table are bit-exact
    for r in range(world):
        rank, w, first, count = res[r]["info"]
        assert (rank, w) == (r, world) and count >= (1 << 22) // world and first == r * ((1 << 22) // world)
        assert res[r]["init_equal"][0]
        assert np.array_equal(res[r]["pred_sharded"], res[r]["pred_single"])
    # B. sequential training through remote memory is bit-exact with the unsharded run
    last = res[world - 1]
    print("B: narrow sequential max |dp|", float(np.max(np.abs(last["seq_sharded"] - last["seq_single"]))), "tables", last["seq_tables_equal"],
          "| D1: wide sequential max |dp|", float(np.max(np.abs(last["wide_seq_sharded"] - last["wide_seq_single"]))), "tables", last["wide_seq_tables_equal"],
          "| wide one-in-flight (bulk-copy kernel) max |dp|", float(np.max(np.abs(last["wide_one_sharded"] - last["wide_one_single"]))), "table max diff", last["wide_one_tables_maxdiff"])
    assert np.array_equal(last["seq_sharded"].view(np.uint32), last["seq_single"].view(np.uint32))
    assert last["seq_tables_equal"].all(), last["seq_tables_equal"]
    # C. two GPUs training one model concurrently: progressive logloss within 1 % of one GPU training the whole stream
    p = np.concatenate([res[r]["hog_preds"] for r in range(world)])
    y = np.concatenate([res[r]["hog_labels"] for r in range(world)])
    ll_sh = util.logloss(p, y)
    ll_1 = util.logloss(res[0]["hog_single_preds"], res[0]["hog_single_labels"])
    assert abs(ll_sh - ll_1) / ll_1 < 0.01, (ll_sh, ll_1)
    rate = sum(len(res[r]["hog_preds"]) for r in range(world)) / max(float(res[r]["hog_secs2"][0]) for r in range(world))
    print(f"sharded x{world}: logloss {ll_sh:.4f} vs single {ll_1:.4f}; warm pass {rate / 1e6:.1f} M examples/s (wall clock, small batch)")
    # D1. wide model through remote memory: sequential mode (general kernel) is bit-exact with the unsharded run; the bulk-copy
    #     kernel with one record in flight follows it within 1e-5 (its rows come through the copy engine, whose reads of a
    #     peer's memory are not ordered bit-for-bit against the previous record's remote reductions)
    assert np.array_equal(last["wide_seq_sharded"].view(np.uint32), last["wide_seq_single"].view(np.uint32))
    assert last["wide_seq_tables_equal"].all(), last["wide_seq_tables_equal"]
    assert last["wide_seq_paths"][1] == 400
    assert float(np.max(np.abs(last["wide_one_sharded"] - last["wide_one_single"]))) <= 1e-5
    assert last["wide_one_paths"][0] > 0 and last["wide_one_paths"][1] == 0 and last["wide_one_tables_maxdiff"][0] <= 2e-3
    # D2. wide model, every rank trains its slice on ONE model through the owner-side update path
    from fwumious_wabbit_b200 import synth

    n_per = len(res[0]["wide_hog_preds"])
    p = np.concatenate([res[r]["wide_hog_preds"] for r in range(world)])
    y = np.concatenate([res[r]["wide_hog_labels"] for r in range(world)])
    for r in range(world):
        assert res[r]["wide_hog_paths"][0] > 0 and res[r]["wide_hog_paths"][1] == 0
        assert np.array_equal(res[r]["wide_tables_digest"], res[0]["wide_tables_digest"])   # one model: every rank reads the same tables
        assert res[r]["wide_acc_touched"][1] == 1 and res[r]["wide_acc_touched"][2] == 1
    # every gradient row reached its owner and its slot: exactly the slots a single GPU touches on the same stream
    assert res[0]["wide_single_touched_equal"][0] == 1, (res[0]["wide_single_touched_equal"], res[0]["wide_acc_touched"])
    ll_sh = util.logloss(p, y)
    ll_1 = util.logloss(res[0]["wide_single_preds"], y)
    w = synth.workload("c3")
    recs = w.records(world * n_per)
    ora = util.oracle_regressor(w.mi)
    _, want = ora.hogwild(util.oracle_spec(w.mi), recs.reshape(-1), np.arange(world * n_per + 1, dtype=np.uint64) * w.record_len, 1, want_preds=True)
    ll_o = util.logloss(want, y)
    rate = world * n_per / max(float(res[r]["wide_hog_secs2"][0]) for r in range(world))
    print(f"wide sharded x{world}: logloss {ll_sh:.4f} vs single GPU {ll_1:.4f} vs sequential oracle {ll_o:.4f}; warm pass {rate / 1e6:.2f} M examples/s (wall clock)")
    assert abs(ll_sh - ll_o) / ll_o < 0.02 and abs(ll_sh - ll_1) / ll_1 < 0.02, (ll_sh, ll_1, ll_o)
