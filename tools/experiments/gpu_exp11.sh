#!/bin/bash
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), ('e2e %.1fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'], 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None)
" $1 "$2" 2>&1 | tail -1; }
for S in 0 1 0 1; do
FWGPU_SNAP=$S timeout 300 python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/exp11_c2_s$S.json 2> gpurun_out/exp11.err; summ gpurun_out/exp11_c2_s$S.json "c2 snap=$S"; tail -2 gpurun_out/exp11.err
done
