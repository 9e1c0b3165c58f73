#!/bin/bash
# 8-GPU session: sharded tests at world 2 and 8, the one-model c4 line at N = 8 and 4, the default bench at N = 8 (replicas + extra.c4_one_model)
TAG=${1:-r02g8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu_$TAG.txt 2>&1
nvidia-smi nvlink -gt d -i 0 | head -8
(timeout 1200 python -m pytest tests/test_gpu_shard.py -m gpu -q -s > gpurun_out/pytest_shard_8gpu_$TAG.txt 2>&1)
grep -E "^FAILED|^ERROR|passed|failed|sharded x|wide sharded|^B:" gpurun_out/pytest_shard_8gpu_$TAG.txt | head -20 | cut -c1-400
grep -E "^E  |rank [0-9] failed" gpurun_out/pytest_shard_8gpu_$TAG.txt | head -12 | cut -c1-300
run() { # N name extra-args...
  N=$1; name=$2; shift; shift
  (NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_${name}_${N}gpu_$TAG.json 2> gpurun_out/bench_${name}_${N}gpu_$TAG.err)
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${name}_${N}gpu_$TAG.json").read())
    def show(k,x):
        if not x or "error" in x: print(k, x); return
        r=x.get("roofline") or {}
        print(k, "x$N: value %.2fM"%(x["value"]/1e6), "e2e", (x.get("e2e") or {}).get("value"), "launch ms %.3f"%r.get("avg_launch_ms",0), "ms/step %.1f"%x["ms_per_step"], "per-rank", [round(v,1) for v in x.get("per_rank_ms_per_step",[])], x.get("nvlink"))
    show("$name", d)
    for k,x in (d.get("extra") or {}).items(): show("  extra."+k, x)
except Exception as e: print("$name x$N: bench parse failed", e); print(open("gpurun_out/bench_${name}_${N}gpu_$TAG.err").read()[-1500:])
PY
}
run 8 c4_one_model --workload c4 --no-extra
run 4 c4_one_model --workload c4 --no-extra --no-e2e
run 8 default
