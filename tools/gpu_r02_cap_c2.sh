TAG=r02f
(FWGPU_RAMP_DIV=4294967295 FWGPU_CHUNK_MB=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_learn_fixed -s 4 -c 1 -o /tmp/ncu_c2_fixed_$TAG python bench.py --workload c2 --examples 6000000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_c2_fixed_$TAG.log 2>&1)
python tools/ncu_summary.py /tmp/ncu_c2_fixed_$TAG.ncu-rep k_learn_fixed --traffic c2 --examples 6000000 --top 14 > gpurun_out/ncu_c2_fixed_full_$TAG.txt 2>&1
cp profiles/traffic_c2.json gpurun_out/
head -24 gpurun_out/ncu_c2_fixed_full_$TAG.txt | cut -c1-220
