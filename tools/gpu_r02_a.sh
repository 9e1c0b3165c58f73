#!/bin/bash
# Round-2 session A: the whole -m gpu suite (new fused-kernel parity tests included) and the new bench line (c3 headline + extras).
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi -L
nproc
(timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu_$TAG.txt 2>&1)
grep -E "AssertionError|Mismatched|Max abs|^FAILED|passed|failed|^E  |max \|d\||logloss" gpurun_out/pytest_gpu_$TAG.txt | head -40
tail -25 gpurun_out/pytest_gpu_$TAG.txt
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err)
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$TAG.json").read())
    def show(k,x):
        if not x or "error" in x: print(k, x); return
        r=x.get("roofline") or {}
        print(k, "value %.2fM"%(x["value"]/1e6), "e2e", (x.get("e2e") or {}).get("value"), "frac %.3f"%r.get("frac",0), "launch ms %.3f"%r.get("avg_launch_ms",0), "share %.3f"%r.get("kernel_share_of_step",0), x.get("kernel_paths"), x["clocks"].get("sm_mhz"))
    show("headline", d)
    for k,x in d["extra"].items(): show(k,x)
    print("cpu", json.dumps(d["cpu_baseline"])[:900])
except Exception as e: print("bench parse failed", e)
PY
