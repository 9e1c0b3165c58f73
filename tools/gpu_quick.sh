#!/bin/bash
# quick GPU check: parity tests, then bench lines for the workloads named on the command line (default c2)
# usage: tools/gpu_quick.sh tag [workloads...]      env: SKIP_TESTS=1, BENCH_ARGS="..."
TAG=${1:-q}; shift
WL=${@:-c2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.txt 2>&1); grep -E "AssertionError|Mismatch|Max abs|^FAILED|passed|failed|^E  " gpurun_out/pytest_gpu_$TAG.txt | head -20
fi
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), ('e2e %.1fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'], 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None, 'head', (r.get('head') or {}).get('fp32_tflops'))
" $1 "$2" 2>&1 | tail -1; }
for W in $WL; do
timeout 400 python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/q_${W}_$TAG.json 2> gpurun_out/q_${W}_$TAG.err; summ gpurun_out/q_${W}_$TAG.json "$W"; tail -2 gpurun_out/q_${W}_$TAG.err
done
