#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_head.py -m gpu -q > gpurun_out/pytest_head.txt 2>&1); grep -E "^FAILED|passed|failed|^E  |Error" gpurun_out/pytest_head.txt | head -30
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']; h=r.get('head') or {}
print(sys.argv[2], 'value %.2fM ex/s'%(d['value']/1e6), ('e2e %.2fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'head share %.3f tflops %.1f ms/pass %.3f'%(h.get('share_of_step',0),h.get('fp32_tflops',0),h.get('ms_per_pass',0)), 'k_learn share %.3f'%r['kernel_share_of_step'], 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None, 'launches', d['gpu_launches'])
" $1 "$2" 2>&1 | tail -1; }
for HB in 4096 8192; do for TILE in 0 64; do
FWGPU_HEAD_BATCH=$HB FWGPU_HEAD_TILE=$TILE timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c5_hb${HB}_t$TILE.json 2> gpurun_out/c5.err; summ gpurun_out/c5_hb${HB}_t$TILE.json "c5 hb=$HB tile=$TILE"; tail -2 gpurun_out/c5.err
done; done
