#!/bin/bash
# k_learn_fixed: two records per warp (G=16) vs one (G=32) on c2
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_exp10.txt 2>&1); grep -E "AssertionError|Mismatch|Max abs|^FAILED|passed|failed|^E  " gpurun_out/pytest_gpu_exp10.txt | head -20
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), ('e2e %.1fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'], 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None)
" $1 "$2" 2>&1 | tail -1; }
for G in 1 0; do
FWGPU_G16=$G timeout 300 python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/exp10_c2_g$G.json 2> gpurun_out/exp10.err; summ gpurun_out/exp10_c2_g$G.json "c2 g16=$G"; tail -2 gpurun_out/exp10.err
done
FWGPU_G16=1 timeout 300 python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --uniform-ids > gpurun_out/exp10_c2_g1u.json 2> gpurun_out/exp10.err; summ gpurun_out/exp10_c2_g1u.json "c2 g16=1 uniform"
FWGPU_G16=1 timeout 300 python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --predict-only > gpurun_out/exp10_c2_g1p.json 2> gpurun_out/exp10.err; summ gpurun_out/exp10_c2_g1p.json "c2 g16=1 predict-only"
