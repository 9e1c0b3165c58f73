#!/bin/bash
# e2e of c2: chunk size of the host-record pipeline (FWGPU_CHUNK_MB = slab budget of the general path; unset = 16 MB of records)
mkdir -p gpurun_out
for MB in unset 384 768 192; do
if [ $MB = unset ]; then unset FWGPU_CHUNK_MB; else export FWGPU_CHUNK_MB=$MB; fi
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp13_$MB.json 2> gpurun_out/exp13.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/exp13_$MB.json').read()); print('chunk_mb=$MB value %.1fM e2e %.1fM e2e ms/step %.2f'%(d['value']/1e6,d['e2e']['value']/1e6,d['e2e']['ms_per_step']))"
done
