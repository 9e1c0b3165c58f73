#!/bin/bash
# experiments: GPU tests, then bench variants selected by environment knobs
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_exp.txt 2>&1)
grep -E "AssertionError:|Mismatched|Max abs|^FAILED|passed|failed|Error" gpurun_out/pytest_gpu_exp.txt | head -20
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), ('e2e %.1fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'], 'share %.3f'%r['kernel_share_of_step'], 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None)
" $1 "$2" 2>&1 | tail -1; }
for F in 1 0; do
  FWGPU_FAST=$F timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/exp_c2_fast$F.json 2> gpurun_out/exp_c2_fast$F.err; summ gpurun_out/exp_c2_fast$F.json "c2 fast=$F"; tail -2 gpurun_out/exp_c2_fast$F.err
done
for M in 2 3 4; do
  FWGPU_MINB=$M timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/exp_c3_minb$M.json 2> gpurun_out/exp_c3_minb$M.err; summ gpurun_out/exp_c3_minb$M.json "c3 minb=$M"; tail -2 gpurun_out/exp_c3_minb$M.err
done
