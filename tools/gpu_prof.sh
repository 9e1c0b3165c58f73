#!/bin/bash
# ncu full capture of k_learn for c2 and c3 (after the concurrency ramp), plus launch lists.
TAG=${1:-r01}
mkdir -p gpurun_out
for W in c2 c3; do
  EX=2000000; [ $W = c3 ] && EX=400000
  (timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_learn -s 19 -c 2 -f -o gpurun_out/prof_${W}_$TAG \
     python bench.py --workload $W --examples $EX --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_${W}_$TAG.log 2>&1)
  tail -3 gpurun_out/ncu_full_${W}_$TAG.log | cut -c1-300
done
ls -la gpurun_out/*.ncu-rep
