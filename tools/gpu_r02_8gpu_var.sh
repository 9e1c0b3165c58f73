#!/bin/bash
# 8-GPU one-model variants (chunk size, overlap, push side alone)
TAG=${1:-r02g8v}; N=${2:-8}
mkdir -p gpurun_out
run() { # name env...
  name=$1; shift
  UNI=""; case $name in uniform*) UNI="--uniform-ids";; esac
  (env "$@" NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-e2e $UNI > gpurun_out/bench_c4_${name}_${N}gpu_$TAG.json 2> gpurun_out/bench_c4_${name}_${N}gpu_$TAG.err)
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c4_${name}_${N}gpu_$TAG.json").read())
    r=d.get("roofline") or {}
    print("$name x$N: value %.2fM"%(d["value"]/1e6), "launch ms %.3f"%r.get("avg_launch_ms",0), "ex/launch %d" % r.get("examples_per_launch",0), "ms/step %.1f" % d["ms_per_step"], "per-rank", [round(v,1) for v in d.get("per_rank_ms_per_step",[])])
except Exception as e: print("$name: bench parse failed", e); print(open("gpurun_out/bench_c4_${name}_${N}gpu_$TAG.err").read()[-1200:])
PY
}
run chunk16k FWGPU_SHARD_CHUNK=16384
run chunk32k FWGPU_SHARD_CHUNK=32768
run chunk16k_overlap FWGPU_SHARD_CHUNK=16384 FWGPU_SHARD_OVERLAP=1
run chunk16k_noapply FWGPU_SHARD_CHUNK=16384 FWGPU_SHARD_NO_APPLY=1
run uniform_chunk16k FWGPU_SHARD_CHUNK=16384
