#!/bin/bash
# dense head on the tensor cores (tcgen05 3xTF32) vs fp32 FFMA tiles: head tests, then c5 bench both ways
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_head.py -m gpu -q -x > gpurun_out/pytest_head_umma.txt 2>&1); grep -E "AssertionError|Mismatch|Max abs|^FAILED|passed|failed|^E  " gpurun_out/pytest_head_umma.txt | head -20
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']; h=r.get('head') or {}
print(sys.argv[2], 'value %.2fM ex/s'%(d['value']/1e6), ('e2e %.2fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'head ms/pass %.3f share %.2f tflops %.1f'%(h.get('ms_per_pass',0),h.get('share_of_step',0),h.get('fp32_tflops',0)), 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None)
" $1 "$2" 2>&1 | tail -1; }
for R in 512 0; do
FWGPU_HEAD_UMMA_ROWS=$R timeout 400 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/exp12_c5_u$R.json 2> gpurun_out/exp12.err; summ gpurun_out/exp12_c5_u$R.json "c5 umma_rows=$R"; tail -2 gpurun_out/exp12.err
done
