#!/bin/bash
# N-GPU session: the sharded-table tests (narrow model: direct remote access; wide model: owner-side update path) and the one-model
# c4 bench line under torchrun.   usage: tools/gpu_r02_shard.sh TAG N
TAG=${1:-r02s}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
(timeout 1200 python -m pytest tests/test_gpu_shard.py -m gpu -q -s -x > gpurun_out/pytest_shard_${TAG}.txt 2>&1)
grep -E "^FAILED|^ERROR|passed|failed|sharded x|wide sharded" gpurun_out/pytest_shard_${TAG}.txt | head -20
grep -E "^E  |rank [0-9] failed|Error|error" gpurun_out/pytest_shard_${TAG}.txt | head -30 | cut -c1-300
(NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_c4_one_model_${N}gpu_${TAG}.json 2> gpurun_out/bench_c4_one_model_${N}gpu_${TAG}.err)
tail -5 gpurun_out/bench_c4_one_model_${N}gpu_${TAG}.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c4_one_model_${N}gpu_${TAG}.json").read())
    r=d.get("roofline") or {}
    print("c4 one model x$N: value %.2fM"%(d["value"]/1e6), "e2e", (d.get("e2e") or {}).get("value"), "frac %.3f"%r.get("frac",0), "launch ms %.3f"%r.get("avg_launch_ms",0), "share %.3f"%r.get("kernel_share_of_step",0), d.get("kernel_paths"), "per-rank ms", d.get("per_rank_ms_per_step"), "ll", (d.get("e2e") or {}).get("last_step_logloss"))
except Exception as e: print("bench parse failed", e)
PY
