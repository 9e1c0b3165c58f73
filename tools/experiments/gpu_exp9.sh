#!/bin/bash
# k_learn_fixed variants on c2: accumulator snapshot on/off, 3 or 2 blocks per SM
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_exp9.txt 2>&1); grep -E "AssertionError|Mismatch|Max abs|^FAILED|passed|failed|^E  " gpurun_out/pytest_gpu_exp9.txt | head -20
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), ('e2e %.1fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'], 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None)
" $1 "$2" 2>&1 | tail -1; }
for CFG in "1 3" "1 2" "0 3"; do
set -- $CFG
FWGPU_SNAP=$1 FWGPU_FIXED_MINB=$2 timeout 300 python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/exp9_c2_s$1_m$2.json 2> gpurun_out/exp9.err; summ gpurun_out/exp9_c2_s$1_m$2.json "c2 snap=$1 minb=$2"; tail -2 gpurun_out/exp9.err
done
