#!/bin/bash
# BASELINE config 4 on one 8-GPU box: one model, FFM tables (ffm_bit_precision 28: 1 GiB + 1 GiB) hash-range-sharded over the
# GPUs through NVLink peer memory; every rank feeds its own shard of the example stream
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --sharded --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/shard_c4_$N.json 2> gpurun_out/shard_c4_$N.err
echo "rc=$?"; tail -3 gpurun_out/shard_c4_$N.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/shard_c4_$N.json').read()); print('n_gpus', d['n_gpus'], 'value %.2fM ex/s'%(d['value']/1e6), 'e2e %.2fM'%(d['e2e']['value']/1e6), 'logloss', d['e2e']['last_step_logloss'], d['config']['parallelism'])"
