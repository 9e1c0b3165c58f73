#!/usr/bin/env python
"""bench.py -- examples/sec of FFM training on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus 1 --steps K --warmup W                       # our arm (CUDA, libfwgpu.so)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N replicas, one per GPU
  python bench.py --impl reference ...                                # the reference's CPU algorithm (oracle port), host cores

A step = one training pass of the hot path over one batch of synthetic records of the named
workload (default: BASELINE.json configs[1], "c2": FFM k=4, 8 fields, ffm_bit_precision=20, 10M
examples per step).  `value` is measured with the records already resident in HBM; `e2e` goes
through the C-ABI call that takes HOST buffers (H2D of the records and D2H of the predictions inside
the timed region).  Multi-GPU = independent replicas on disjoint example shards (the path has no
exchange step; DESIGN.md "multi-GPU"), reported as weak scaling.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--examples", type=int, default=0, help="examples per step (0 = workload default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="examples in the cpu_baseline sample (0 = default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sharded", action="store_true",
                    help="N > 1: ONE model whose tables are hash-range-sharded over the N GPUs (NVLink peer memory) instead of N replicas")
    ap.add_argument("--predict-only", action="store_true", help="diagnostic: time the forward pass only (update = 0)")
    ap.add_argument("--uniform-ids", action="store_true", help="diagnostic: uniform feature ids instead of Zipf (no hot rows)")
    return ap.parse_args()


DEFAULT_EXAMPLES = {"c1": 10_000_000, "c2": 10_000_000, "c3": 2_000_000, "c4": 2_000_000, "c5": 1_000_000}
CPU_SAMPLE = {"c1": 4_000_000, "c2": 2_000_000, "c3": 100_000, "c4": 100_000, "c5": 15_000}
MAX_SLICES = 12
REF_STEP_EXAMPLES = {"c1": 8_000_000, "c2": 4_000_000, "c3": 200_000, "c4": 200_000, "c5": 30_000}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (pynvml, 10 ms period).
    NVML is initialised in the constructor, i.e. before the timed region starts, so even a region of a
    few tens of milliseconds gets samples; one more sample is taken when the region closes."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz, self.err = index, False, [], set(), None, None
        self.nv = self.h = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def sample(self):
        if self.h is None:
            return
        try:
            self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.01)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "note": self.err or "no samples"}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def dist_setup(n_gpus):
    from fwumious_wabbit_b200 import dist_util

    return dist_util.init("nccl")


def cpu_baseline(w, n_sample, threads, seed=1):
    """The reference's algorithm on the host cores: oracle port, Hogwild with `threads` workers
    (hogwild.rs:24-103) or the sequential loop (main.rs:213-258) when threads == 1."""
    from oracle import fw_oracle as fo
    from tests import util

    recs = w.records(n_sample, first=0, seed=seed)
    ora = util.oracle_regressor(w.mi)
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n_sample + 1, dtype=np.uint64) * w.record_len
    secs, _ = ora.hogwild(spec, recs.reshape(-1), rec_off, threads, want_preds=False)
    return n_sample / secs, secs


def run_reference(args):
    from fwumious_wabbit_b200 import synth

    world, rank, local_rank, dist = 1, int(os.environ.get("RANK", "0")), 0, None
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    w = synth.workload(args.workload)
    threads = os.cpu_count() or 1
    n = args.examples or REF_STEP_EXAMPLES[args.workload]
    from oracle import fw_oracle as fo
    from tests import util

    ora = util.oracle_regressor(w.mi)
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    total_steps = args.warmup + args.steps
    # a fresh slice of the stream per step, generated outside the timed region
    t_sum = 0.0
    for s in range(total_steps):
        recs = w.records(n, first=s * n, seed=1)
        secs, _ = ora.hogwild(spec, recs.reshape(-1), rec_off, threads, want_preds=False)
        if s >= args.warmup:
            t_sum += secs
    value = n * args.steps / t_sum
    line = {
        "impl": "reference", "metric": "examples/sec FFM training", "value": value, "unit": "examples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_sum / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{w.name}: {w.description}", "examples_per_step": n,
                   "note": "reference Rust cannot be built here (no cargo/rustc); this is the C port of its algorithm (oracle/), Hogwild threads on the host cores"},
        "cpu_baseline": {"value": value, "unit": "examples/s", "cores": threads, "kind": "port",
                         "sample": f"{n} examples per step x {args.steps} steps of the {w.name} stream, Hogwild {threads} threads"},
        "e2e": {"value": value, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    import torch

    import fwumious_wabbit_b200 as fw
    from fwumious_wabbit_b200 import synth

    world, rank, local_rank, dist = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    w = synth.workload(args.workload)
    n = args.examples or DEFAULT_EXAMPLES[args.workload]
    sharded = args.sharded and world > 1
    shard = (rank, world, f"/tmp/fwgpu_shard_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}") if sharded else None
    re = fw.Regressor(w.mi, device=local_rank, shard=shard)
    stream = torch.cuda.ExternalStream(re.stream_ptr(), device=torch.device("cuda", local_rank))

    # every rank trains its own replica on a disjoint shard of the stream
    L = fw._lib.lib()
    import ctypes as C

    nbytes = n * w.record_len * 4
    # an online learner must not see the same batch twice: every step (warm-up included) trains on a FRESH slice of the
    # stream (repeating one batch drives the gradients to zero and lets the kernel skip work).  At most MAX_SLICES slices
    # are kept; longer runs cycle through them.
    n_slices = max(1, min(args.warmup + args.steps, MAX_SLICES, max(1, (12 << 30) // nbytes)))
    hp = C.c_void_p()
    assert L.fwgpu_host_alloc(C.byref(hp), nbytes * n_slices) == 0
    recs_all = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(n_slices * n, w.record_len))
    from fwumious_wabbit_b200 import dist_util

    first, _ = dist_util.shard(rank, world, n * n_slices)
    w.records(n * n_slices, first=first * 1, seed=1, out=recs_all, uniform=args.uniform_ids)
    pp = C.c_void_p()
    assert L.fwgpu_host_alloc(C.byref(pp), n * 4) == 0
    preds = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), shape=(n,))
    ds = re.upload_dataset(recs_all.reshape(-1), n_examples=n * n_slices)

    def slice_of(step):
        return (step % n_slices) * n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        re.sync()

    def max_over_ranks(x):
        return dist_util.max_over_ranks(x, dist, device="cuda")

    # ---------------- value: records resident in HBM ----------------
    upd = not args.predict_only
    for i in range(args.warmup):
        re.learn_dataset(ds, slice_of(i), n, update=True, sync=False)
    barrier()
    re.set_profiling(True)
    re.kernel_time(0); re.kernel_time(1); re.kernel_time(2)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = re.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        re.learn_dataset(ds, slice_of(args.warmup + i), n, update=upd, sync=False)
    ev1.record(stream)
    sampler.sample()  # the queue is still draining here: a sample under load even for a very short region
    barrier()
    sampler.stop_flag = True
    launches = re.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    k_ms, k_n = re.kernel_time(0)
    t_ms, t_n = re.kernel_time(1)
    h_ms, h_n = re.kernel_time(2)
    re.set_profiling(False)
    sampler.join(timeout=2)
    value = dist_util.whole_job_rate(n, args.steps, world, ms_total)

    # ---------------- e2e: host buffers through the C ABI ----------------
    e2e = None
    if not args.no_e2e:
        def host_slice(step):
            return recs_all[slice_of(step):slice_of(step) + n]

        e_warm = max(1, min(args.warmup, 2))
        for i in range(e_warm):
            re.learn_records(host_slice(i).reshape(-1), n_examples=n, update=True, out=preds, sync=False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(args.steps):
            re.learn_records(host_slice(e_warm + i).reshape(-1), n_examples=n, update=True, out=preds, sync=False)
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
        e_ms = max(e_ms, max_over_ranks(wall_ms) * 0.0)  # device-timed; wall clock kept for the record below
        e2e = {"value": world * n * args.steps / (e_ms * 1e-3), "unit": "examples/s",
               "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(n * 4),
               "ms_per_step": e_ms / args.steps, "wall_ms_per_step": max_over_ranks(wall_ms) / args.steps}
        recs = host_slice(e_warm + args.steps - 1)
        ll = float(-np.mean(np.where(recs[:, 1] == 1, np.log(np.clip(preds, 1e-7, 1)), np.log(np.clip(1 - preds, 1e-7, 1)))))
        e2e["last_step_logloss"] = ll
        e2e["note"] = "second epoch over the slices the value run trained on (same model, host buffers)"

    # ---------------- roofline of the dominant kernel (k_learn) ----------------
    peak, peak_src = measured_peaks()
    alg_bytes = w.algorithmic_bytes_per_example(train=True)
    roof = None
    if k_n:
        ex_per_launch = n * args.steps / k_n
        achieved = alg_bytes * ex_per_launch / (k_ms / k_n * 1e-3) * 1e-9
        traffic = None
        tp = os.path.join(ROOT, "profiles", f"traffic_{w.name}.json")
        if os.path.exists(tp):
            try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch, scaled to this run's launch size
                traffic = json.load(open(tp)).get("dram_bytes_per_example") * ex_per_launch
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": {"c2": "k_learn_fixed<16,4,1,OPT_LUT> (16 lanes per record, two records per warp)",
                           "c3": "k_learn_fixed_cta<2,0> (block per record)", "c4": "k_learn_fixed_cta<2,0> (block per record)",
                           "c5": "k_learn_fixed_cta<2,1> + <2,2> (block per record, forward / update phases around the head)"}.get(w.name, "k_learn"),
                "peak_source": peak_src, "algorithmic_bytes_per_example": alg_bytes,
                "examples_per_launch": ex_per_launch, "avg_launch_ms": k_ms / k_n, "launches_timed": int(k_n),
                "kernel_share_of_step": k_ms / (ev0.elapsed_time(ev1)), "translate_ms_per_launch": (t_ms / t_n) if t_n else None}
        sp = os.path.join(ROOT, "profiles", f"skeleton_{w.name}.json")
        if os.path.exists(sp) and not args.predict_only:
            try:  # the kernel's memory operations alone (tools/skeleton_microbench.cu): what the memory system sustains for this access pattern
                sk = json.load(open(sp))
                ref_rate = sk["records_per_s_uniform_ids"] if args.uniform_ids else sk["records_per_s_bench_id_law"]
                roof["memory_skeleton"] = {"records_per_s": ref_rate, "frac": (ex_per_launch / (k_ms / k_n * 1e-3)) / ref_rate, "source": sk["source"]}
            except Exception:
                pass
        if h_n:
            # dense head (config 5): GEMMs on the tensor cores (tcgen05, 3xTF32 split operands, fp32 accumulation in TMEM) for
            # sub-batches >= 512 rows, fp32 FFMA tiles below that; reported against the head's own fp32-equivalent flop count
            # (forward + the two backward GEMMs + the squared-gradient sums of every layer), not against the HBM roofline of
            # the gather/scatter kernel
            mi = w.mi
            x_len = mi.num_combos + len(mi.ffm_fields) * (len(mi.ffm_fields) + 1) // 2
            dims, n_in = [], x_len
            for layer in mi.nn_layers:
                dims.append((n_in, int(layer.get("width", 20)))); n_in = dims[-1][1]
            dims.append((n_in + x_len, 1))
            fma = sum(a * b * (1 + 1 + 2) for a, b in dims)  # forward, input gradient, sum g and sum g^2
            roof["head"] = {"ms_per_pass": h_ms / h_n, "passes": int(h_n), "share_of_step": h_ms / ev0.elapsed_time(ev1),
                            "fp32_tflops": 2 * fma * n * args.steps / (h_ms * 1e-3) * 1e-12, "fma_per_example": fma,
                            "engine": "fp32 FFMA tiles" if os.environ.get("FWGPU_HEAD_UMMA_ROWS") == "0" else
                                      "tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators (sub-batches >= 512 rows)"}

    # ---------------- cpu baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = args.cpu_sample or CPU_SAMPLE[args.workload]
        v, secs = cpu_baseline(w, ns, 1)
        cpu = {"value": v, "unit": "examples/s", "cores": 1, "kind": "port",
               "sample": f"first {ns} examples of the same stream, sequential learn (reference default mode), {secs:.1f} s"}

    if rank == 0:
        line = {
            "metric": "examples/sec FFM training", "value": value, "unit": "examples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{w.name}: {w.description}" + (" [DIAGNOSTIC: uniform ids]" if args.uniform_ids else "") + (" [DIAGNOSTIC: predict only]" if args.predict_only else ""), "examples_per_step_per_gpu": n, "fresh_slices": n_slices,
                       "parallelism": (f"one model, tables hash-range-sharded x{world} over NVLink peer memory, disjoint example shards" if sharded else
                                       f"replicas x{world} (independent models, disjoint example shards)") if world > 1 else "single GPU",
                       "l2_policy": f"inputs larger than L2: {nbytes >> 20} MiB of records per step; table {(w.mi.ffm_k and ((1 << w.mi.ffm_bit_precision) * 8 >> 20))} MiB w+acc vs 126 MB L2 (c2's 8 MiB table is L2-resident by nature; the records are not)",
                       "optimizer": "AdagradLUT", "semantics": "Hogwild on device, chunked launches"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        }
        emit(line)
    ds.free()
    re.close()
    L.fwgpu_host_free(hp)
    L.fwgpu_host_free(pp)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's real stdout; everything else any library prints on fd 1 (NCCL's version
    banner, for instance) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    # NCCL prints its version banner on stdout when NCCL_DEBUG is VERSION/INFO; stdout carries exactly one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("FWGPU_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
