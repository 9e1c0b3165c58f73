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


def _one_model_worker(rank, world, port, n_per, out_dir):
    """The exchange protocol of the one-model path with gloo and numpy standing in for NVLink and CUDA:
    every rank turns its records into (row base, field, gradient row) entries, buckets them by the row's OWNER with the same
    shift the push kernel uses (fwgpu_debug_shard_plan), all-gathers the per-owner counts (the step NCCL does per chunk) and
    hands each owner its entries; owners apply them to the range they hold."""
    import ctypes as C

    from fwumious_wabbit_b200 import _lib

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    _, _, _, d = dist_util.init("gloo")
    F, k, bits = 6, 4, 14
    Fk, n_floats = F * k, 1 << bits
    w = synth.Workload("t", synth._mi(F, ffm_k=k, ffm_bits=bits, bits=12), synth.NS_LETTERS[:F], [40, 40, 500, 500, 5000, 5000], "tiny")
    sizes = (C.c_uint64 * world)()
    shift = C.c_uint32(0)
    gran = 4096
    assert _lib.lib().fwgpu_debug_shard_plan(n_floats * 4, (Fk + 64) * 4, world, gran, sizes, C.byref(shift)) == 0
    offs = np.concatenate([[0], np.cumsum(list(sizes))]) // 4          # float ranges per owner
    first, _ = dist_util.shard(rank, world, n_per)
    recs = w.records(n_per, first=first)
    mask = ((1 << bits) - 1) ^ (k - 1)
    # entries this rank produces: integer "gradients" so that the result does not depend on the order of application
    base = (recs[:, 3:3 + F] & np.uint32(mask)).astype(np.int64)                       # [n, F]
    grad = (base[:, :, None] * 7 + np.arange(Fk)[None, None, :] * 3 + np.arange(F)[None, :, None]) % 11 - 5   # [n, F, Fk]
    owner = np.minimum(base >> shift.value, world - 1)
    counts = np.array([int(np.sum(owner == r)) for r in range(world)], dtype=np.int64)
    all_counts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    d.all_gather(all_counts, torch.from_numpy(counts))                                # counts_all[s][r]: what rank s sends to owner r
    outbox = [(base[owner == r], grad[owner == r]) for r in range(world)]
    inbox = [None] * world
    gathered = [None] * world
    d.all_gather_object(gathered, outbox)                                             # stand-in for the bulk stores into the owners' inboxes
    for s in range(world):
        inbox[s] = gathered[s][rank]
        assert len(inbox[s][0]) == int(all_counts[s][rank])                           # the owner reads exactly what the all-gather announced
    # owner-side apply on the range this rank holds (+ the tail on the last rank)
    table = np.zeros(int(offs[-1]), dtype=np.int64)
    for b, g in inbox:
        for row, gr in zip(b, g):
            assert offs[rank] <= row < offs[rank + 1], (rank, row)                      # every entry reached the rank that holds its row
            table[row:row + Fk] += gr                                                  # a window may spill over the range's end: last rank's tail, or the neighbour's memory through the shared mapping
    np.save(os.path.join(out_dir, f"table{rank}.npy"), table)
    np.save(os.path.join(out_dir, f"counts{rank}.npy"), np.stack([c.numpy() for c in all_counts]))
    d.barrier()
    d.destroy_process_group()


def test_one_model_exchange_protocol_two_ranks(tmp_path):
    world, n_per = 2, 3000
    mp.spawn(_one_model_worker, args=(world, _free_port(), n_per, str(tmp_path)), nprocs=world, join=True)
    F, k, bits = 6, 4, 14
    Fk = F * k
    w = synth.Workload("t", synth._mi(F, ffm_k=k, ffm_bits=bits, bits=12), synth.NS_LETTERS[:F], [40, 40, 500, 500, 5000, 5000], "tiny")
    recs = w.records(world * n_per)
    mask = ((1 << bits) - 1) ^ (k - 1)
    base = (recs[:, 3:3 + F] & np.uint32(mask)).astype(np.int64)
    grad = (base[:, :, None] * 7 + np.arange(Fk)[None, None, :] * 3 + np.arange(F)[None, :, None]) % 11 - 5
    tables = [np.load(tmp_path / f"table{r}.npy") for r in range(world)]
    want = np.zeros_like(tables[0])
    for b, g in zip(base.reshape(-1), grad.reshape(-1, Fk)):
        want[b:b + Fk] += g
    assert np.array_equal(sum(tables), want)                     # one model: the owners' updates add up to the single-process result
    c0, c1 = np.load(tmp_path / "counts0.npy"), np.load(tmp_path / "counts1.npy")
    assert np.array_equal(c0, c1) and c0.sum() == world * n_per * F and np.all(c0 > 0)
