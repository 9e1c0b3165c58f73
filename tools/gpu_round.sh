#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list and one full capture of k_learn.
# usage: tools/gpu_round.sh [tag]    (outputs under gpurun_out/)
TAG=${1:-r01}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1)
grep -E "AssertionError:|Mismatched|Max abs|^FAILED|passed|failed" gpurun_out/pytest_gpu_$TAG.txt | head -30
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err)
(timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err)
kill $SMI
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c2_$TAG.json 2> gpurun_out/bench_ref_c2_$TAG.err)
for W in c2 c3; do
  EX=2000000; [ $W = c3 ] && EX=400000
  (timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${W}_$TAG.csv \
     python bench.py --workload $W --examples $EX --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_${W}_$TAG.log 2>&1)
  (timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_learn -s 24 -c 2 -f -o gpurun_out/prof_${W}_$TAG \
     python bench.py --workload $W --examples $EX --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_${W}_$TAG.log 2>&1)
done
python - <<PY
import json
for w in ("c2","c3"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{w}_$TAG.json").read())
        r=d["roofline"]; print(w, "value %.1fM ex/s"%(d["value"]/1e6), "e2e %.1fM"%(d["e2e"]["value"]/1e6), "frac %.3f"%r["frac"], "launch ms %.3f"%r["avg_launch_ms"], "share %.3f"%r["kernel_share_of_step"], "logloss", d["e2e"]["last_step_logloss"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, d["clocks"])
    except Exception as e: print(w, "bench parse failed", e)
try: print("ref", open("gpurun_out/bench_ref_c2_$TAG.json").read()[:400])
except Exception as e: print(e)
PY
ls -la gpurun_out | tail -20
