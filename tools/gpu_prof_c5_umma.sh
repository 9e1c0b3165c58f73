#!/bin/bash
# ncu --set full capture of the head's tcgen05 GEMMs in a steady-state c5 pass (forward, error, update kinds)
mkdir -p gpurun_out
(FWGPU_RAMP_DIV=4294967295 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm -s 36 -c 6 -f -o gpurun_out/prof_c5_umma \
   python bench.py --workload c5 --examples 200000 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c5_umma.log 2>&1)
tail -2 gpurun_out/ncu_full_c5_umma.log | cut -c1-200
ls -la gpurun_out/prof_c5_umma.ncu-rep
