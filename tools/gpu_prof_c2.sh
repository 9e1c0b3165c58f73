#!/bin/bash
# full ncu capture of one steady-state k_learn_fixed launch on c2; env passes through (FWGPU_SNAP, FWGPU_FIXED_MINB)
TAG=${1:-cur}
mkdir -p gpurun_out
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_learn_fixed -s 24 -c 1 -f -o gpurun_out/prof_c2_$TAG \
   python bench.py --workload c2 --examples 2000000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c2_$TAG.log 2>&1)
tail -2 gpurun_out/ncu_full_c2_$TAG.log | cut -c1-200
ls -la gpurun_out/prof_c2_$TAG.ncu-rep
