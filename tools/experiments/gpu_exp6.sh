#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_persistence.py -m gpu -q > gpurun_out/pytest_gpu_exp6.txt 2>&1); grep -E "^FAILED|passed|failed|^E  " gpurun_out/pytest_gpu_exp6.txt | head -10
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.1fM ex/s'%(d['value']/1e6), 'launch ms %.3f'%r['avg_launch_ms'])
" $1 "$2" 2>&1 | tail -1; }
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/exp6_c3.json 2> gpurun_out/exp6.err; summ gpurun_out/exp6_c3.json "c3 train"
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --predict-only > gpurun_out/exp6_c3p.json 2> gpurun_out/exp6.err; summ gpurun_out/exp6_c3p.json "c3 predict-only"
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --predict-only > gpurun_out/exp6_c2p.json 2> gpurun_out/exp6.err; summ gpurun_out/exp6_c2p.json "c2 predict-only"
