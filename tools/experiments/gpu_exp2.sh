#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/hotrow_microbench > gpurun_out/hotrow.txt 2>&1; cat gpurun_out/hotrow.txt
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sequential or fast_path or records_path" > gpurun_out/pytest_gpu_exp2.txt 2>&1); tail -3 gpurun_out/pytest_gpu_exp2.txt
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), 'frac %.3f'%r['frac'], 'launch ms %.3f'%r['avg_launch_ms'])
" $1 "$2" 2>&1 | tail -1; }
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --uniform-ids > gpurun_out/exp_c2_uniform.json 2> gpurun_out/exp_c2_uniform.err; summ gpurun_out/exp_c2_uniform.json "c2 uniform fast"
FWGPU_FAST=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --uniform-ids > gpurun_out/exp_c2_uniform_g.json 2> gpurun_out/exp_c2_uniform_g.err; summ gpurun_out/exp_c2_uniform_g.json "c2 uniform general"
for M in 3 4; do for S in 0 1; do
FWGPU_MINB=$M FWGPU_SIMPLE_UPDATE=$S timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/exp_c3_m${M}s$S.json 2> gpurun_out/exp_c3.err; summ gpurun_out/exp_c3_m${M}s$S.json "c3 minb=$M simple=$S"
done; done
FWGPU_MINB=4 FWGPU_SIMPLE_UPDATE=1 timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --uniform-ids > gpurun_out/exp_c3_uni.json 2> gpurun_out/exp_c3.err; summ gpurun_out/exp_c3_uni.json "c3 uniform minb=4 simple=1"
