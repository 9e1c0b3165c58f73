"""One rank of the 2-GPU sharded-table test (spawned by tests/test_gpu_shard.py, one process per GPU).
usage: shard_worker.py RANK WORLD PREFIX OUTDIR"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fwumious_wabbit_b200 as fw  # noqa: E402
from fwumious_wabbit_b200 import synth  # noqa: E402

SEQUENTIAL = 0x7FFFFFFF


def wl():
    """config 2's shape with a 16 MiB FFM table: 8 MiB per rank, whole allocation granules on any driver"""
    w = synth.workload("c2")
    w.mi.ffm_bit_precision = 22
    return w


def main():
    rank, world, prefix, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    out = {}
    w = wl()
    n = 6000
    recs = w.records(n)

    # A. tables initialised shard by shard by their owners == the single-GPU initialisation; remote rows read over NVLink
    sh = fw.Regressor(w.mi, device=rank, shard=(rank, world, prefix + ".a"))
    single = fw.Regressor(w.mi, device=rank)
    out["info"] = np.array(sh.shard_info(), dtype=np.int64)
    out["pred_sharded"] = sh.learn_records(recs.reshape(-1), n_examples=n, update=False)
    out["pred_single"] = single.learn_records(recs.reshape(-1), n_examples=n, update=False)
    out["init_equal"] = np.array([np.array_equal(sh.get_ffm()[0], single.get_ffm()[0])])
    sh.shard_barrier()
    sh.close(); single.close()

    # B. the LAST rank alone trains, one example in flight, through mostly remote memory (the LR table lives on rank 0,
    #    half of the FFM rows too): predictions and tables bit-exact with the same run on an unsharded table
    ws = wl()
    ws.mi.hogwild_ramp_div = SEQUENTIAL
    sh = fw.Regressor(ws.mi, device=rank, shard=(rank, world, prefix + ".b"))
    if rank == world - 1:
        single = fw.Regressor(ws.mi, device=rank)
        out["seq_sharded"] = sh.learn_records(recs.reshape(-1), n_examples=n, update=True)
        out["seq_single"] = single.learn_records(recs.reshape(-1), n_examples=n, update=True)
        sh.sync()
        sw, sa = sh.get_ffm(); uw, ua = single.get_ffm()
        out["seq_tables_equal"] = np.array([np.array_equal(sw, uw), np.array_equal(sa, ua), np.array_equal(sh.get_lr_table(), single.get_lr_table())])
        single.close()
    sh.shard_barrier()
    sh.close()

    # C. every rank trains its own half of a stream on the ONE shared model, concurrently (Hogwild across GPUs)
    m = 400_000
    wc = wl()
    recs_c = wc.records(m)
    half = m // world
    mine = recs_c[rank * half:(rank + 1) * half]
    sh = fw.Regressor(wc.mi, device=rank, shard=(rank, world, prefix + ".c"))
    sh.shard_barrier()
    t = time.time()
    out["hog_preds"] = sh.learn_records(mine.reshape(-1), n_examples=half, update=True)
    sh.shard_barrier()
    out["hog_secs"] = np.array([time.time() - t])
    # second pass: throughput with a warm model, no ramp
    t = time.time()
    sh.learn_records(mine.reshape(-1), n_examples=half, update=True)
    sh.shard_barrier()
    out["hog_secs2"] = np.array([time.time() - t])
    out["hog_labels"] = (mine[:, 1] == 1).astype(np.float32)
    if rank == 0:
        single = fw.Regressor(wl().mi, device=0)
        out["hog_single_preds"] = single.learn_records(recs_c.reshape(-1), n_examples=m, update=True)
        out["hog_single_labels"] = (recs_c[:, 1] == 1).astype(np.float32)
        single.close()
    sh.shard_barrier()
    sh.close()

    # D. WIDE model (config 3's shape: 39 fields x k = 8): the bulk-copy kernel.  Rows are pulled from their owners over NVLink;
    #    D1 one record in flight (direct mode: gradient rows go back as bulk reductions into the owner's L2), last rank alone;
    #    D2 every rank trains its slice of one stream on the ONE model: gradient rows are pushed to the owners' inboxes, one
    #       NCCL all-gather per chunk, owners apply AdaGrad (k_learn_rows<PUSH> + k_apply_inbox).
    os.environ["FWGPU_SHARD_CHUNK"] = "2048"
    n1 = 400
    for tag, setter in (("wide_seq", lambda mi: setattr(mi, "hogwild_ramp_div", SEQUENTIAL)),      # sequential mode: the general kernel
                        ("wide_one", lambda mi: setattr(mi, "hogwild_max_inflight", 1))):            # one record in flight: the bulk-copy kernel
        wd = synth.workload("c3")
        setter(wd.mi)
        recs_d = wd.records(n1)
        sh = fw.Regressor(wd.mi, device=rank, shard=(rank, world, prefix + "." + tag))
        if rank == world - 1:
            single = fw.Regressor(wd.mi, device=rank)
            out[tag + "_sharded"] = sh.learn_records(recs_d.reshape(-1), n_examples=n1, update=True)
            out[tag + "_single"] = single.learn_records(recs_d.reshape(-1), n_examples=n1, update=True)
            sh.sync()
            sw, sa = sh.get_ffm(); uw, ua = single.get_ffm()
            out[tag + "_tables_equal"] = np.array([np.array_equal(sw, uw), np.array_equal(sa, ua), np.array_equal(sh.get_lr_table(), single.get_lr_table())])
            out[tag + "_tables_maxdiff"] = np.array([float(np.max(np.abs(sw - uw))), float(np.max(np.abs(sa - ua)))])
            out[tag + "_paths"] = np.array([sh.path_counts()["fixed_cta"], sh.path_counts()["general_examples"]])
            single.close()
        sh.shard_barrier()
        sh.close()

    n_per = int(os.environ.get("FWGPU_TEST_SHARD_N", "40000"))
    we = synth.workload("c3")
    mine = we.records(n_per, first=rank * n_per)
    sh = fw.Regressor(we.mi, device=rank, shard=(rank, world, prefix + ".d2"))
    sh.shard_barrier()
    t = time.time()
    out["wide_hog_preds"] = sh.learn_records(mine.reshape(-1), n_examples=n_per, update=True)
    sh.shard_barrier()
    out["wide_hog_secs"] = np.array([time.time() - t])
    out["wide_hog_labels"] = (mine[:, 1] == 1).astype(np.float32)
    out["wide_hog_paths"] = np.array([sh.path_counts()["fixed_cta"], sh.path_counts()["general_examples"]])
    # a second, warm pass for a throughput figure, then the tables as every rank sees them
    more = we.records(n_per, first=(world + rank) * n_per)
    sh.shard_barrier()
    t = time.time()
    sh.learn_records(more.reshape(-1), n_examples=n_per, update=True, want_preds=False)
    sh.shard_barrier()
    out["wide_hog_secs2"] = np.array([time.time() - t])
    w_all, a_all = sh.get_ffm()
    import hashlib
    out["wide_tables_digest"] = np.frombuffer(hashlib.sha256(w_all.tobytes() + a_all.tobytes()).digest(), dtype=np.uint8)
    out["wide_acc_touched"] = np.array([int(np.count_nonzero(a_all)), int(np.all(np.isfinite(w_all))), int(np.all(a_all >= 0))])
    if rank == 0:
        single = fw.Regressor(synth.workload("c3").mi, device=0)
        allrecs = we.records(2 * world * n_per)
        out["wide_single_preds"] = single.learn_records(allrecs[:world * n_per].reshape(-1), n_examples=world * n_per, update=True)
        single.learn_records(allrecs[world * n_per:].reshape(-1), n_examples=world * n_per, update=True, want_preds=False)
        _, a1 = single.get_ffm()
        out["wide_single_touched_equal"] = np.array([int(np.array_equal(a1 != 0, a_all != 0)), int(np.count_nonzero(a1))])
        single.close()
    sh.shard_barrier()
    sh.close()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **out)
    print(f"rank {rank} done", flush=True)


if __name__ == "__main__":
    main()
