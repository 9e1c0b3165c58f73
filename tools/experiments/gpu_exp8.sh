#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_exp8.txt 2>&1); grep -E "^FAILED|passed|failed|^E  " gpurun_out/pytest_gpu_exp8.txt | head -12
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), ('e2e %.1fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'], 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None)
" $1 "$2" 2>&1 | tail -1; }
for UB in 1 2 4; do
FWGPU_UB=$UB timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/exp8_c3_ub$UB.json 2> gpurun_out/exp8.err; summ gpurun_out/exp8_c3_ub$UB.json "c3 cta ub=$UB"; tail -2 gpurun_out/exp8.err
done
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --predict-only > gpurun_out/exp8_c3p.json 2> gpurun_out/exp8.err; summ gpurun_out/exp8_c3p.json "c3 cta predict-only"
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/exp8_c3_full.json 2> gpurun_out/exp8.err; summ gpurun_out/exp8_c3_full.json "c3 default with e2e"
