#!/bin/bash
# launch lists (all kernels, gpu time) for c5 at full sub-batch size, full captures of the c3 block-per-record kernel and the c5 head GEMMs
mkdir -p gpurun_out
(FWGPU_RAMP_DIV=4294967295 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_c5.csv \
   python bench.py --workload c5 --examples 200000 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_c5.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_learn_fixed_cta -s 14 -c 1 -f -o gpurun_out/prof_c3_cta \
   python bench.py --workload c3 --examples 400000 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c3.log 2>&1)
(FWGPU_RAMP_DIV=4294967295 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_head_gemm -s 70 -c 7 -f -o gpurun_out/prof_c5_head \
   python bench.py --workload c5 --examples 200000 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c5.log 2>&1)
tail -3 gpurun_out/ncu_full_c3.log | cut -c1-200; tail -3 gpurun_out/ncu_full_c5.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
