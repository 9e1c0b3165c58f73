#!/bin/bash
# tests + a bench line with the rows kernel and one with the old block-per-record kernel (FWGPU_ROWS=0) for comparison
TAG=${1:-r02c}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_$TAG.txt 2>&1)
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_$TAG.txt | head -40
grep -E "^E  " gpurun_out/pytest_gpu_$TAG.txt | head -50
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err)
tail -3 gpurun_out/bench_$TAG.err
(FWGPU_ROWS=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_oldcta.json 2> gpurun_out/bench_${TAG}_oldcta.err)
python - <<PY
import json
for f in ("gpurun_out/bench_$TAG.json", "gpurun_out/bench_${TAG}_oldcta.json"):
  try:
    d=json.loads(open(f).read())
    def show(k,x):
        if not x or "error" in x: print(k, x); return
        r=x.get("roofline") or {}
        print(k, "value %.2fM"%(x["value"]/1e6), "e2e", (x.get("e2e") or {}).get("value"), "frac %.3f"%r.get("frac",0), "launch ms %.3f"%r.get("avg_launch_ms",0), "share %.3f"%r.get("kernel_share_of_step",0), x.get("kernel_paths"), "ll", (x.get("e2e") or {}).get("last_step_logloss"))
    print(f); show("headline", d)
    for k,x in d["extra"].items(): show(k,x)
  except Exception as e: print("bench parse failed", f, e)
PY
