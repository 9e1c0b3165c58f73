#!/bin/bash
# One GPU-box session: parity tests, bench lines (c2 headline, c3, c5), reference arm, ncu launch lists.
# usage: tools/gpu_round.sh [tag]    (outputs under gpurun_out/)
TAG=${1:-r01}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1)
grep -E "AssertionError:|Mismatched|Max abs|^FAILED|passed|failed|^E  " gpurun_out/pytest_gpu_$TAG.txt | head -30
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err)
(timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err)
(timeout 400 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err)
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c2_$TAG.json 2> gpurun_out/bench_ref_c2_$TAG.err)
# launch lists of the bench command itself (all kernels, gpu time).  c2 / c3 keep the concurrency ramp (a dozen short launches at
# the start), c5 skips its ramp (hundreds of tiny sub-batches) so that the list shows steady-state sub-batches
for W in c2 c3 c5; do
  EX=2000000; [ $W = c3 ] && EX=400000; [ $W = c5 ] && EX=200000
  RD=32; [ $W = c5 ] && RD=4294967295
  (FWGPU_RAMP_DIV=$RD timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${W}_$TAG.csv \
     python bench.py --workload $W --examples $EX --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_${W}_$TAG.log 2>&1)
done
python - <<PY
import json
for w in ("c2","c3","c5"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{w}_$TAG.json").read()); r=d["roofline"]
        print(w, "value %.2fM ex/s"%(d["value"]/1e6), "e2e %.2fM"%(d["e2e"]["value"]/1e6), "frac %.3f"%r["frac"], "launch ms %.3f"%r["avg_launch_ms"], "share %.3f"%r["kernel_share_of_step"], "head", (r.get("head") or {}).get("fp32_tflops"), "logloss", d["e2e"]["last_step_logloss"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, d["clocks"])
    except Exception as e: print(w, "bench parse failed", e)
try: print("ref", open("gpurun_out/bench_ref_c2_$TAG.json").read()[:300])
except Exception as e: print(e)
PY
