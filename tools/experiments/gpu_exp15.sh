#!/bin/bash
# c5: update GEMMs of the head on a side stream (fork/join) vs everything on one stream
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_head.py -m gpu -q -x > gpurun_out/pytest_head_fork.txt 2>&1); grep -E "AssertionError|Mismatch|Max abs|^FAILED|passed|failed|^E  " gpurun_out/pytest_head_fork.txt | head -20
for NF in 0 1; do
if [ $NF = 1 ]; then export FWGPU_HEAD_NO_FORK=1; else unset FWGPU_HEAD_NO_FORK; fi
timeout 400 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/exp15_c5_nf$NF.json 2> gpurun_out/exp15.err
python -c "
import json
d=json.loads(open('gpurun_out/exp15_c5_nf$NF.json').read()); h=d['roofline']['head']
print('no_fork=$NF value %.2fM e2e %.2fM head ms/pass %.3f share %.2f tflops %.1f logloss %.4f'%(d['value']/1e6,d['e2e']['value']/1e6,h['ms_per_pass'],h['share_of_step'],h['fp32_tflops'],d['e2e']['last_step_logloss']))"
done
