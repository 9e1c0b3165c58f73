#!/bin/bash
# dense-head session: head parity tests, stability experiment
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_head.py -m gpu -q > gpurun_out/pytest_head.txt 2>&1); grep -E "^FAILED|passed|failed|^E  |Error" gpurun_out/pytest_head.txt | head -30
(timeout 900 python tools/head_stability.py > gpurun_out/head_stability.txt 2>&1); cat gpurun_out/head_stability.txt | tail -30
