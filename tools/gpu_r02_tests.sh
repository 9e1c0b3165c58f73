#!/bin/bash
# the whole -m gpu suite, all failures listed
TAG=${1:-r02}
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu_$TAG.txt 2>&1)
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_$TAG.txt | head -40
grep -E "^E  " gpurun_out/pytest_gpu_$TAG.txt | head -60
grep -E "max \|d\||logloss|sub-batch" gpurun_out/pytest_gpu_$TAG.txt | head
tail -18 gpurun_out/pytest_gpu_$TAG.txt
