#!/bin/bash
# Round-2 closing session on ONE GPU: the whole -m gpu suite, the driver's two bench commands, launch list and full ncu captures of
# the dominant kernels (steady-state launches: FWGPU_RAMP_DIV=4294967295 skips the concurrency ramp under the profiler).
TAG=${1:-r02f}
mkdir -p gpurun_out
[ -z "$SKIP_TESTS" ] && (timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu_$TAG.txt 2>&1)
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_$TAG.txt | head -20
grep -E "^E  " gpurun_out/pytest_gpu_$TAG.txt | head -20 | cut -c1-300
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err); tail -2 gpurun_out/bench_$TAG.err
(timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err); tail -2 gpurun_out/bench_ref_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$TAG.json").read())
    def show(k,x):
        if not x or "error" in x: print(k, x); return
        r=x.get("roofline") or {}
        print(k, "value %.2fM"%(x["value"]/1e6), "e2e", (x.get("e2e") or {}).get("value"), "frac %.3f"%r.get("frac",0), "launch ms %.3f"%r.get("avg_launch_ms",0), "share %.3f"%r.get("kernel_share_of_step",0), x.get("kernel_paths"), "ll", (x.get("e2e") or {}).get("last_step_logloss"), r.get("traffic_note"))
    show("headline", d)
    for k,x in d["extra"].items(): show(k,x)
    print("cpu", json.dumps(d["cpu_baseline"])[:700])
    r=json.loads(open("gpurun_out/bench_ref_$TAG.json").read()); print("reference arm", r["value"], r["config"]==d["config"], r["cpu_baseline"]["cores"])
except Exception as e: print("bench parse failed", e)
PY
[ -z "$SKIP_TESTS" ] && (timeout 600 python tools/diag_r02.py 2>&1 | tail -4 | cut -c1-500 | tee gpurun_out/diag_$TAG.txt)
# launch list of the bench command (all kernels, gpu time)
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3_$TAG.csv python bench.py --workload c3 --examples 600000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_list_c3_$TAG.log 2>&1)
python tools/launch_summary.py gpurun_out/launches_c3_$TAG.csv 2>/dev/null | head -12
# full captures, one steady-state launch each
cap() { # name kernel-regex skip workload examples extra-args traffic-name
  (FWGPU_RAMP_DIV=4294967295 FWGPU_CHUNK_MB=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/ncu_$1_$TAG python bench.py --workload $4 --examples $5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra $6 > gpurun_out/ncu_$1_$TAG.log 2>&1)
  # the report stays on the box (gpurun_out is capped at 64 MiB): its summary and the traffic file travel
  python tools/ncu_summary.py /tmp/ncu_$1_$TAG.ncu-rep $2 --traffic $7 --examples $5 --top 14 > gpurun_out/ncu_$1_full_$TAG.txt 2>&1
  cp profiles/traffic_$7.json gpurun_out/ 2>/dev/null
  head -22 gpurun_out/ncu_$1_full_$TAG.txt | cut -c1-200
}
cap c3_rows k_learn_rows 4 c3 600000 "" c3
cap c4x1_rows k_learn_rows 4 c4 600000 "" c4x1
cap c4x1_uniform_rows k_learn_rows 4 c4 600000 "--uniform-ids" c4x1_uniform
cap c2_fixed k_learn_fixed 4 c2 3000000 "" c2
