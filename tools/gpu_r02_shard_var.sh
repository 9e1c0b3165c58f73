#!/bin/bash
# one-model c4 bench variants under torchrun (diagnostics of the owner-side update path)   usage: TAG N
TAG=${1:-r02v}; N=${2:-2}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -q -s > gpurun_out/pytest_shard_${TAG}.txt 2>&1)
grep -E "^FAILED|^ERROR|passed|failed|sharded x|wide sharded|^B:" gpurun_out/pytest_shard_${TAG}.txt | head -20
grep -E "^E  |rank [0-9] failed" gpurun_out/pytest_shard_${TAG}.txt | head -12 | cut -c1-300
run() { # name, env...
  name=$1; shift
  (env "$@" NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/bench_c4_${name}_${N}gpu_${TAG}.json 2> gpurun_out/bench_c4_${name}_${N}gpu_${TAG}.err)
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c4_${name}_${N}gpu_${TAG}.json").read())
    r=d.get("roofline") or {}
    print("$name x$N: value %.2fM"%(d["value"]/1e6), "launch ms %.3f"%r.get("avg_launch_ms",0), "share %.3f"%r.get("kernel_share_of_step",0), "ex/launch %d" % r.get("examples_per_launch",0), "ms/step", d["ms_per_step"])
except Exception as e: print("$name: bench parse failed", e); print(open("gpurun_out/bench_c4_${name}_${N}gpu_${TAG}.err").read()[-1500:])
PY
}
run default X=1
run overlap FWGPU_SHARD_OVERLAP=1
run overlap_apply1 FWGPU_SHARD_OVERLAP=1 FWGPU_SHARD_APPLY_BLOCKS=1
