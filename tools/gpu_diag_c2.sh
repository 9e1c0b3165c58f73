#!/bin/bash
# what bounds k_learn_fixed on c2: Zipf vs uniform ids, learn vs predict-only
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'])
" $1 "$2" 2>&1 | tail -1; }
for V in "" "--uniform-ids" "--predict-only" "--uniform-ids --predict-only"; do
T=$(echo "$V" | tr -d ' -'); T=${T:-base}
timeout 300 python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $V > gpurun_out/diag_c2_$T.json 2> gpurun_out/diag.err; summ gpurun_out/diag_c2_$T.json "c2 $V"; tail -2 gpurun_out/diag.err
done
