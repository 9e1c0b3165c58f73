"""world_size-2 gloo test (CPU) of the multi-GPU host logic: disjoint contiguous shards of the
synthetic stream, max-over-ranks timing, whole-job aggregation.  The data path has no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fwumious_wabbit_b200 import dist_util, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    w_, r_, lr_, d = dist_util.init("gloo")
    assert (w_, r_) == (world, rank) and d is not None
    w = synth.workload("c2")
    first, count = dist_util.shard(rank, world, n)
    recs = w.records(count, first=first, seed=1)
    # every rank reports its shard's checksum and its (fake) step time
    digest = torch.tensor([int(recs.astype(np.uint64).sum() % (1 << 62)), first, count], dtype=torch.int64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    d.all_gather(gathered, digest)
    fake_ms = 10.0 + 5.0 * rank
    mx = dist_util.max_over_ranks(fake_ms, d)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([g.tolist() for g in gathered] + [[int(mx * 1000), 0, 0]]))
    d.barrier()
    d.destroy_process_group()


def test_two_rank_shards_and_timing(tmp_path):
    world, n = 2, 5000
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    w = synth.workload("c2")
    whole = w.records(world * n, first=0, seed=1)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        for q in range(world):
            want = int(whole[q * n:(q + 1) * n].astype(np.uint64).sum() % (1 << 62))
            assert got[q].tolist() == [want, q * n, n]          # shards are the disjoint contiguous slices of one stream
        assert got[world][0] == int((10.0 + 5.0 * (world - 1)) * 1000)  # max over ranks
    assert dist_util.whole_job_rate(n, 3, world, 15.0) == world * n * 3 / 0.015


def test_shard_plan_and_owner_mapping():
    """Layout arithmetic of the one-model path (fwgpu_create_sharded), no GPU: equal hash ranges over the ranks, the spill-over
    tail (block_ffm.rs:93-94) on the last rank, and the shift the push kernel uses to find a row's owner agrees with the byte
    ranges -- for config 4 (ffm_bit_precision 28, 39 fields x k = 8) on 2, 4 and 8 ranks; a table too small to split into
    whole allocation granules stays on rank 0."""
    import ctypes as C

    from fwumious_wabbit_b200 import _lib

    L = _lib.lib()
    gran = 2 << 20
    for world in (2, 4, 8):
        sizes = (C.c_uint64 * world)()
        shift = C.c_uint32(0)
        n_floats, tail = 1 << 28, 39 * 8 + 64
        assert L.fwgpu_debug_shard_plan(n_floats * 4, tail * 4, world, gran, sizes, C.byref(shift)) == 0
        sizes = list(sizes)
        assert all(s % gran == 0 for s in sizes) and sum(sizes) >= (n_floats + tail) * 4
        assert sizes[:-1] == [n_floats * 4 // world] * (world - 1) and sizes[-1] >= n_floats * 4 // world + tail * 4
        offs = np.concatenate([[0], np.cumsum(sizes)])
        rng = np.random.default_rng(world)
        for i in list(rng.integers(0, n_floats, 2000)) + [0, n_floats - 1, n_floats, n_floats + tail - 1] + [r * n_floats // world for r in range(world)]:
            owner = min(int(i) >> shift.value, world - 1)
            assert offs[owner] <= int(i) * 4 < offs[owner + 1], (world, i, owner)
    sizes = (C.c_uint64 * 2)()
    shift = C.c_uint32(0)
    assert L.fwgpu_debug_shard_plan((1 << 18) * 8, 0, 2, gran, sizes, C.byref(shift)) == 0   # a 2 MiB LR table: one granule
    assert list(sizes) == [gran, 0] and shift.value == 32
