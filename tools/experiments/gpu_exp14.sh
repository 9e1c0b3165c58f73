#!/bin/bash
# c5: cap of the head's sub-batch (rows per pass) now that its GEMMs run on the tensor cores
mkdir -p gpurun_out
for HB in 4096 8192 16384; do
FWGPU_HEAD_BATCH=$HB timeout 400 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/exp14_c5_hb$HB.json 2> gpurun_out/exp14.err
python -c "
import json
d=json.loads(open('gpurun_out/exp14_c5_hb$HB.json').read()); h=d['roofline']['head']
print('head_batch=$HB value %.2fM e2e %.2fM head ms/pass %.3f passes %d share %.2f tflops %.1f logloss %.4f'%(d['value']/1e6,d['e2e']['value']/1e6,h['ms_per_pass'],h['passes'],h['share_of_step'],h['fp32_tflops'],d['e2e']['last_step_logloss']))"
done
