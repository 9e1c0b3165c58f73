#!/bin/bash
# ncu full capture: c2 fused kernel (k_learn_fixed) and c3 general kernel (k_learn), after the concurrency ramp.
TAG=${1:-r01}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1); tail -3 gpurun_out/pytest_gpu_$TAG.txt
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_learn_fixed -s 17 -c 1 -f -o gpurun_out/prof_c2_$TAG \
   python bench.py --workload c2 --examples 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c2_$TAG.log 2>&1)
tail -2 gpurun_out/ncu_full_c2_$TAG.log | cut -c1-200
(timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_learn$' -s 14 -c 1 -f -o gpurun_out/prof_c3_$TAG \
   python bench.py --workload c3 --examples 400000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c3_$TAG.log 2>&1)
tail -2 gpurun_out/ncu_full_c3_$TAG.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
