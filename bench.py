#!/usr/bin/env python
"""bench.py -- examples/sec of FFM training on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus 1 --steps K --warmup W                       # our arm (CUDA, libfwgpu.so)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU
  python bench.py --impl reference ...                                # the reference's CPU algorithm (oracle port), host cores

A step = one training pass of the hot path over one batch of synthetic records.  The headline workload is BASELINE.json's
config 3 ("c3": FFM k=8, 39 fields, ffm_bit_precision 24 -- the largest single-GPU FFM configuration), 2 M examples per step
and GPU; `value` is measured with the records already resident in HBM, `e2e` goes through the C-ABI call that takes HOST
buffers (H2D of the records and D2H of the predictions inside the timed region).  Every timed step -- warm-up, value and e2e
alike -- trains on a slice of the stream the model has never seen.

The same run also measures, under `extra` (shorter runs, same rules): c2, c5 (dense head), c3 predict-only and "c4x1" --
config 4's 2^28-row table (1 GiB + 1 GiB, far beyond the 126 MB L2) on ONE GPU with Zipf and with uniform ids, the
HBM-resident measurement of the gather/scatter kernel.  `--workload X` makes X the headline instead.

Multi-GPU (N > 1): the headline stays c3 as N independent replicas on disjoint example shards (the path has no exchange
step; `value` must be one workload at every N for the driver's scaling curve), and `extra.c4_one_model` is ONE model whose
2^28-row tables are hash-range-sharded over the N GPUs (rows pulled over NVLink, gradients pushed to the owner in batches;
DESIGN.md section 6).  `--workload c4` under torchrun makes that one-model run the headline.
"""
import argparse
import hashlib
import json
import os
import subprocess
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
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--examples", type=int, default=0, help="examples per step and GPU (0 = workload default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="examples in the cpu_baseline sample (0 = default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only")
    ap.add_argument("--replicas", action="store_true", help="N > 1 with --workload c4: independent replicas instead of one sharded model")
    ap.add_argument("--predict-only", action="store_true", help="diagnostic: time the forward pass only (update = 0)")
    ap.add_argument("--uniform-ids", action="store_true", help="diagnostic: uniform feature ids instead of Zipf (no hot rows)")
    return ap.parse_args()


# examples per step and GPU -- the same in both arms (the driver compares the two lines' configs)
STEP_EXAMPLES = {"c1": 10_000_000, "c2": 10_000_000, "c3": 2_000_000, "c4": 2_000_000, "c5": 1_000_000}
CPU_SAMPLE = {"c1": 4_000_000, "c2": 2_000_000, "c3": 100_000, "c4": 100_000, "c5": 15_000}          # sequential, 1 thread
CPU_SAMPLE_HOGWILD = {"c1": 16_000_000, "c2": 8_000_000, "c3": 1_000_000, "c4": 1_000_000, "c5": 100_000}  # multi-thread legs
EXTRA_DEADLINE_S = 420
KERNEL_NAME = {"c2": "k_learn_fixed<16,4,1,OPT_LUT> (16 lanes per record, two records per warp)",
               "c3": "k_learn_rows<0,LUT> (block per record: one bulk copy per row and array in, pair-owned in-place update, one bulk reduction per row and array out)",
               "c4": "k_learn_rows<0,LUT> (block per record: one bulk copy per row and array in, pair-owned in-place update, one bulk reduction per row and array out)",
               "c5": "k_learn_rows<1> + <2> (forward / update phases around the head's GEMMs)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_sha():
    """Identity of the device code the numbers come from: profiles/traffic_*.json carries the hash of the build it was
    captured on and is ignored (traffic = null) when the kernels have changed since."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "fwumious_wabbit_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


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


def nvlink_bytes(index):
    """(tx_bytes, rx_bytes) summed over the GPU's NVLinks from the driver's data counters (`nvidia-smi nvlink -gt d`), or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        tx = rx = 0
        for line in out.splitlines():
            if "Data Tx" in line:
                tx += int(line.split(":")[-1].strip().split()[0])
            elif "Data Rx" in line:
                rx += int(line.split(":")[-1].strip().split()[0])
        return (tx * 1024, rx * 1024) if (tx or rx) else None
    except Exception:  # noqa: BLE001
        return None


def make_config(w, n, world, one_model, args, tags=()):
    """The `config` object -- built by the same function in both arms."""
    table_mib = (w.mi.ffm_k and ((1 << w.mi.ffm_bit_precision) * 8 >> 20)) or 0
    return {
        "workload": f"{w.name}: {w.description}" + "".join(f" [{t}]" for t in tags),
        "examples_per_step_per_gpu": n,
        "parallelism": "single GPU" if world == 1 else
                       (f"one model, tables hash-range-sharded x{world} (rows pulled over NVLink, gradients pushed to the owner in batches), disjoint example shards"
                        if one_model else f"replicas x{world} (independent models, disjoint example shards)"),
        "l2_policy": f"inputs larger than L2: {n * w.record_len * 4 >> 20} MiB of fresh records per step; table {table_mib} MiB w+acc vs 126 MB L2",
        "optimizer": "AdagradLUT", "semantics": "Hogwild (lock-free concurrent updates of one shared table)",
        "fresh_data": "every step (warm-up, timed, e2e) trains on a slice of the stream the model has not seen",
    }


# ----------------------------------------------------------------------------------------------- CPU arm
def _oracle(native=True):
    from oracle import fw_oracle as fo

    flags = fo.use_native_build() if native else "-O3 -march=x86-64-v3"
    return fo, flags


def cpu_run(w, n_sample, threads, first=0, seed=1):
    """The reference's algorithm on the host cores: oracle port, Hogwild with `threads` workers (hogwild.rs:24-103) or
    the sequential loop (main.rs:213-258) when threads == 1.  Returns (examples/s, seconds)."""
    from tests import util

    recs = w.records(n_sample, first=first, seed=seed)
    ora = util.oracle_regressor(w.mi)
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n_sample + 1, dtype=np.uint64) * w.record_len
    secs, _ = ora.hogwild(spec, recs.reshape(-1), rec_off, threads, want_preds=False)
    return n_sample / secs, secs


def probe_reference_binary(w, n_sample, threads):
    """if the driver left a prebuilt reference at baseline/_ref/fw, time the real binary on the same stream
    (text rendered from the same generator, cache built first so that parsing is outside the timed run)."""
    exe = os.path.join(ROOT, "baseline", "_ref", "fw")
    if not (os.path.isfile(exe) and os.access(exe, os.X_OK)):
        return {"available": False, "why": "baseline/_ref/fw not present (the reference is Rust; no cargo/rustc in the image, crates not vendored)"}
    import tempfile

    try:
        d = tempfile.mkdtemp(prefix="fwref_")
        open(os.path.join(d, "vw_namespace_map.csv"), "w").write("".join(f"{c},f{c}\n" for c in w.ns_names))
        with open(os.path.join(d, "train.vw"), "w") as f:
            for i in range(n_sample):
                f.write(w.line(i) + "\n")
        mi = w.mi
        flags = []
        for c in w.ns_names:
            flags += ["--keep", c]
        if mi.ffm_k:
            for c in w.ns_names:
                flags += ["--ffm_field", c]
            flags += ["--ffm_k", str(mi.ffm_k), "--ffm_bit_precision", str(mi.ffm_bit_precision), "--ffm_learning_rate", str(mi.ffm_learning_rate),
                      "--ffm_power_t", str(mi.ffm_power_t), "--ffm_init_acc_gradient", str(mi.ffm_init_acc_gradient)]
        flags += ["-b", str(mi.bit_precision), "-l", str(mi.learning_rate), "--power_t", str(mi.power_t), "--adaptive", "--sgd", "-c",
                  "--data", os.path.join(d, "train.vw")]
        subprocess.run([exe] + flags + ["--build_cache_without_training"], capture_output=True, timeout=900, check=True)
        t0 = time.perf_counter()
        subprocess.run([exe] + flags + ["--hogwild_training", "--hogwild_threads", str(threads)], capture_output=True, timeout=900, check=True)
        secs = time.perf_counter() - t0
        return {"available": True, "value": n_sample / secs, "unit": "examples/s", "cores": threads, "kind": "reference",
                "sample": f"{n_sample} examples from the cache, fw --hogwild_training --hogwild_threads {threads}, wall clock incl. start-up"}
    except Exception as e:  # noqa: BLE001
        return {"available": False, "why": f"baseline/_ref/fw failed: {e}"}


def cpu_baseline_block(w, args):
    """Bounded samples of the same stream on the host cores: sequential (the reference's default mode), Hogwild with the
    reference's default 16 threads (main.rs:193) and with every core."""
    fo, flags = _oracle()
    cores = os.cpu_count() or 1
    ns = args.cpu_sample or CPU_SAMPLE[w.name]
    nh = args.cpu_sample or CPU_SAMPLE_HOGWILD[w.name]
    v1, s1 = cpu_run(w, ns, 1)
    v16, s16 = cpu_run(w, nh, 16)
    vall, sall = (v16, s16) if cores == 16 else cpu_run(w, nh, cores)
    return {"value": vall, "unit": "examples/s", "cores": cores, "kind": "port",
            "sample": f"first {nh} examples of the same stream, Hogwild on all {cores} host cores, {sall:.1f} s",
            "sequential_1_thread": {"value": v1, "cores": 1, "sample": f"first {ns} examples, sequential learn (reference default mode), {s1:.1f} s"},
            "hogwild_16_threads": {"value": v16, "cores": 16, "sample": f"first {nh} examples, 16 threads (reference default --hogwild_threads) on {cores} cores, {s16:.1f} s"},
            "build": f"oracle/fw_oracle.c, gcc {flags} -ffp-contract=off, with the reference's prefetches (block_ffm.rs:163,184,194,205)",
            "reference_binary": probe_reference_binary(w, min(nh, 200_000), cores)}


def run_reference(args):
    from fwumious_wabbit_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    fo, flags = _oracle()
    from tests import util

    w = synth.workload(args.workload)
    threads = os.cpu_count() or 1
    n = args.examples or STEP_EXAMPLES[args.workload]
    ora = util.oracle_regressor(w.mi)
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    total_steps = args.warmup + args.steps
    t_sum = 0.0
    for s in range(total_steps):  # one model over fresh slices of the stream, generated outside the timed region
        recs = w.records(n, first=s * n, seed=1)
        secs, _ = ora.hogwild(spec, recs.reshape(-1), rec_off, threads, want_preds=False)
        if s >= args.warmup:
            t_sum += secs
    value = n * args.steps / t_sum
    line = {
        "impl": "reference", "metric": "examples/sec FFM training", "value": value, "unit": "examples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_sum / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(w, n, args.gpus, False, args),
        "cpu_baseline": {"value": value, "unit": "examples/s", "cores": threads, "kind": "port",
                         "sample": f"{n} examples per step x {args.steps} steps of the {w.name} stream, Hogwild {threads} threads",
                         "build": f"oracle/fw_oracle.c, gcc {flags} -ffp-contract=off, with the reference's prefetches",
                         "note": "the reference is Rust and cannot be built here (no cargo/rustc, crates not vendored): this is the C restatement of its algorithm (oracle/)"},
        "e2e": {"value": value, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- GPU arm
class Env:
    def __init__(self, world, rank, local_rank, dist):
        self.world, self.rank, self.local_rank, self.dist = world, rank, local_rank, dist


def logloss(preds, labels):
    p = np.clip(preds.astype(np.float64), 1e-7, 1 - 1e-7)
    return float(-np.mean(np.where(labels == 1, np.log(p), np.log(1 - p))))


def measure(env, wname, n, steps, warmup, *, do_e2e=True, predict_only=False, uniform=False, one_model=False, max_inflight=0, tags=()):
    """One workload on this rank's GPU: value (records resident in HBM), optionally e2e (host buffers through the C ABI),
    roofline of the dominant kernel.  Returns the dict of a bench line (rank 0) -- collective over ranks."""
    import ctypes as C

    import torch

    import fwumious_wabbit_b200 as fw
    from fwumious_wabbit_b200 import dist_util, synth

    world, rank, local_rank, dist = env.world, env.rank, env.local_rank, env.dist
    w = synth.workload(wname)
    if max_inflight:
        w.mi.hogwild_max_inflight = max_inflight   # 1 = parity mode: the reference's sequential semantics through the fused kernel
    shard = (rank, world, f"/tmp/fwgpu_shard_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}_{wname}") if (one_model and world > 1) else None
    re = fw.Regressor(w.mi, device=local_rank, shard=shard)
    stream = torch.cuda.ExternalStream(re.stream_ptr(), device=torch.device("cuda", local_rank))
    L = fw._lib.lib()

    e_warm = max(1, min(warmup, 2)) if do_e2e else 0
    n_value, n_e2e = warmup + steps, (e_warm + steps) if do_e2e else 0
    n_slices = n_value + n_e2e
    nbytes = n * w.record_len * 4
    hp = C.c_void_p()
    assert L.fwgpu_host_alloc(C.byref(hp), nbytes * n_slices) == 0
    recs_all = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(n_slices * n, w.record_len))
    # rank r's slices are examples [r * n_slices * n, (r + 1) * n_slices * n) of the stream: disjoint between ranks and steps
    first, _ = dist_util.shard(rank, world, n * n_slices)
    w.records(n * n_slices, first=first, seed=1, out=recs_all, uniform=uniform)
    pp = C.c_void_p()
    assert L.fwgpu_host_alloc(C.byref(pp), n * 4) == 0
    preds = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), shape=(n,))
    ds = re.upload_dataset(recs_all[:n_value * n].reshape(-1), n_examples=n_value * n)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        re.sync()
        if shard is not None:
            re.shard_barrier()

    def max_over_ranks(x):
        return dist_util.max_over_ranks(x, dist, device="cuda")

    def all_ranks(x):
        if dist is None:
            return [float(x)]
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    # ---------------- value: records resident in HBM ----------------
    upd = not predict_only
    for i in range(warmup):
        re.learn_dataset(ds, i * n, n, update=True, sync=False)
    barrier()
    re.set_profiling(True)
    re.kernel_time(0); re.kernel_time(1); re.kernel_time(2)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = re.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    nvl0 = nvlink_bytes(local_rank) if (one_model and world > 1 and rank == 0) else None
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for i in range(steps):
        re.learn_dataset(ds, (warmup + i) * n, n, update=upd, sync=False)
    ev1.record(stream)
    sampler.sample()  # the queue is still draining here: a sample under load even for a very short region
    barrier()
    sampler.stop_flag = True
    nvlink = None
    if nvl0 is not None:
        nvl1, wall = nvlink_bytes(local_rank), time.perf_counter() - t_wall0
        if nvl1 is not None:  # driver counters of rank 0's GPU around the timed region (barriers included: a lower bound of the in-kernel rate)
            nvlink = {"tx_bytes_per_example": (nvl1[0] - nvl0[0]) / (n * steps), "rx_bytes_per_example": (nvl1[1] - nvl0[1]) / (n * steps),
                      "tx_gbs_over_wall": (nvl1[0] - nvl0[0]) / wall * 1e-9, "rx_gbs_over_wall": (nvl1[1] - nvl0[1]) / wall * 1e-9,
                      "source": "nvidia-smi nvlink -gt d, GPU of rank 0, before / after the timed steps"}
    launches = re.launch_count() - launches0
    ms_local = ev0.elapsed_time(ev1)
    ms_total = max_over_ranks(ms_local)
    per_rank_ms = all_ranks(ms_local / steps)
    k_ms, k_n = re.kernel_time(0)
    t_ms, t_n = re.kernel_time(1)
    h_ms, h_n = re.kernel_time(2)
    re.set_profiling(False)
    sampler.join(timeout=2)
    value = dist_util.whole_job_rate(n, steps, world, ms_total)
    counts = re.path_counts()

    # ---------------- e2e: host buffers through the C ABI, fresh slices ----------------
    e2e = None
    if do_e2e:
        def host_slice(j):
            return recs_all[(n_value + j) * n:(n_value + j + 1) * n]

        for i in range(e_warm):
            re.learn_records(host_slice(i).reshape(-1), n_examples=n, update=True, out=preds, sync=False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            re.learn_records(host_slice(e_warm + i).reshape(-1), n_examples=n, update=upd, out=preds, sync=False)
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
        e2e = {"value": world * n * steps / (e_ms * 1e-3), "unit": "examples/s",
               "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(n * 4),
               "ms_per_step": e_ms / steps, "wall_ms_per_step": max_over_ranks(wall_ms) / steps,
               "last_step_logloss": logloss(preds, host_slice(e_warm + steps - 1)[:, 1]),
               "note": "fwgpu_learn_records on pinned host records the model has not seen; predictions copied back every step"}

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peaks()
    alg_bytes = w.algorithmic_bytes_per_example(train=upd)
    roof = None
    if k_n:
        ex_per_launch = n * steps / k_n
        achieved = alg_bytes * ex_per_launch / (k_ms / k_n * 1e-3) * 1e-9
        traffic, traffic_note = None, "no ncu capture for this workload in profiles/"
        tname = "c4x1" + ("_uniform" if uniform else "") if (wname == "c4" and world == 1) else wname
        tp = os.path.join(ROOT, "profiles", f"traffic_{tname}.json")
        if os.path.exists(tp):
            try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch, scaled to this run's launch size
                tj = json.load(open(tp))
                if tj.get("kernel_source_sha") == kernel_source_sha():
                    traffic = tj["dram_bytes_per_example"] * ex_per_launch
                    traffic_note = f"ncu capture {tj.get('source', tp)} (same kernel source, sha {tj['kernel_source_sha']}), per launch"
                else:
                    traffic_note = f"profiles/traffic_{tname}.json was captured on other kernel source (sha {tj.get('kernel_source_sha')} vs {kernel_source_sha()}): ignored"
            except Exception as e:  # noqa: BLE001
                traffic_note = f"unreadable traffic file: {e}"
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note, "kernel": KERNEL_NAME.get(w.name, "k_learn"),
                "peak_source": peak_src, "algorithmic_bytes_per_example": alg_bytes,
                "examples_per_launch": ex_per_launch, "avg_launch_ms": k_ms / k_n, "launches_timed": int(k_n),
                "kernel_share_of_step": k_ms / ms_local, "translate_ms_per_launch": (t_ms / t_n) if t_n else None}
        if h_n:
            # dense head (config 5): GEMMs on the tensor cores (tcgen05, 3xTF32 split operands, fp32 accumulation in TMEM) for
            # sub-batches >= 512 rows; reported against the head's own fp32-equivalent flop count (forward + the two backward
            # GEMMs + the squared-gradient sums of every layer), not against the HBM roofline of the gather/scatter kernel
            mi = w.mi
            x_len = mi.num_combos + len(mi.ffm_fields) * (len(mi.ffm_fields) + 1) // 2
            dims, n_in = [], x_len
            for layer in mi.nn_layers:
                dims.append((n_in, int(layer.get("width", 20)))); n_in = dims[-1][1]
            dims.append((n_in + x_len, 1))
            fma = sum(a * b * (1 + 1 + 2) for a, b in dims)  # forward, input gradient, sum g and sum g^2
            roof["head"] = {"ms_per_pass": h_ms / h_n, "passes": int(h_n), "share_of_step": h_ms / ms_local,
                            "fp32_tflops": 2 * fma * n * steps / (h_ms * 1e-3) * 1e-12, "fma_per_example": fma,
                            "engine": "tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators (sub-batches >= 512 rows)"}

    out = None
    if rank == 0:
        tg = list(tags) + (["DIAGNOSTIC: uniform ids"] if uniform else []) + (["predict only"] if predict_only else [])
        out = {
            "metric": "examples/sec FFM training", "value": value, "unit": "examples/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(w, n, world, one_model, None, tg),
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
            "per_rank_ms_per_step": per_rank_ms, "kernel_paths": counts,
        }
        if nvlink:
            out["nvlink"] = nvlink
    ds.free()
    re.close()
    L.fwgpu_host_free(hp)
    L.fwgpu_host_free(pp)
    return out


def run_ours(args):
    import torch

    from fwumious_wabbit_b200 import dist_util, synth

    world, rank, local_rank, dist = dist_util.init("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    env = Env(world, rank, local_rank, dist)
    head = args.workload
    n = args.examples or STEP_EXAMPLES[head]
    one_model = head == "c4" and world > 1 and not args.replicas
    line = measure(env, head, n, args.steps, args.warmup, do_e2e=not args.no_e2e, predict_only=args.predict_only,
                   uniform=args.uniform_ids, one_model=one_model,
                   tags=(["c4x1: the 2^28-row table on one GPU"] if head == "c4" and world == 1 else []))
    extra = {}
    # The headline is measured; whatever happens in the extras (a rank that dies inside a collective would leave the others
    # waiting), the ONE JSON line still goes out: after EXTRA_DEADLINE_S every rank leaves, rank 0 printing what it has.
    def _deadline():
        if rank == 0:
            line["extra"] = dict(extra, _note=f"extras cut off after {EXTRA_DEADLINE_S} s")
            line["cpu_baseline"] = None
            emit(line)
        os._exit(0)

    watchdog = threading.Timer(EXTRA_DEADLINE_S, _deadline)
    watchdog.daemon = True
    watchdog.start()
    if not args.no_extra and not args.predict_only and not args.uniform_ids:
        xs, xw = 3, 3  # extras: shorter runs, same timing rules (>= 3 warm-up steps, fresh data, device timing)

        def sub(key, wname, **kw):
            try:
                r = measure(env, wname, kw.pop("n", STEP_EXAMPLES[wname]), xs, xw, **kw)
                if rank == 0:
                    extra[key] = r
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    extra[key] = {"error": str(e)}

        if world == 1:
            for wn in ("c2", "c3", "c5"):
                if wn != head:
                    sub(wn, wn, do_e2e=False)
            if head != "c4":
                sub("c4x1", "c4", do_e2e=False, tags=["c4x1: the 2^28-row table on one GPU"])
            sub("c4x1_uniform_ids", "c4", do_e2e=False, uniform=True, tags=["c4x1: the 2^28-row table on one GPU"])
            sub(f"{head}_predict_only", head, do_e2e=False, predict_only=True)
            # the parity tier's price: the same fused kernel with ONE record in flight (bit-exact with the reference's
            # sequential learner, tests/test_gpu_fused_parity.py) next to the Hogwild number above
            sub(f"{head}_one_record_in_flight", head, n=20_000, do_e2e=False, max_inflight=1, tags=["parity mode: one record in flight, bit-exact with the sequential reference"])
        elif head != "c4":
            sub("c4_one_model", "c4", do_e2e=False, one_model=True)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(synth.workload(head), args)
    watchdog.cancel()
    if rank == 0:
        line["extra"] = extra
        line["cpu_baseline"] = cpu
        emit(line)
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
    os.dup2(2, 1)  # stdout carries exactly one JSON line; library chatter (NCCL_DEBUG output included) goes to stderr
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
