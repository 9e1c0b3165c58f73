#!/bin/bash
# 2-GPU session: sharded-table test, then sharded vs replica bench lines
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -q -s > gpurun_out/pytest_shard.txt 2>&1); grep -E "^FAILED|passed|failed|skipped|^E  |Error|sharded x" gpurun_out/pytest_shard.txt | head -30; tail -25 gpurun_out/pytest_shard.txt | cut -c1-300
summ() { python -c "
import sys,json
d=json.loads(open(sys.argv[1]).read()); r=d['roofline']
print(sys.argv[2], 'value %.2fM ex/s'%(d['value']/1e6), ('e2e %.2fM'%(d['e2e']['value']/1e6)) if d.get('e2e') else '', 'll', d['e2e']['last_step_logloss'] if d.get('e2e') else None, d['config']['parallelism'][:40])
" $1 "$2" 2>&1 | tail -1; }
for W in c2 c3; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --sharded --workload $W --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/shard_$W.json 2> gpurun_out/shard_$W.err; summ gpurun_out/shard_$W.json "sharded $W"; tail -3 gpurun_out/shard_$W.err | cut -c1-300
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/repl_c3.json 2> gpurun_out/repl_c3.err; summ gpurun_out/repl_c3.json "replicas c3"
