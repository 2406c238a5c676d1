#!/bin/bash
# usage: tools/gpu_benchN.sh <tag> <n> : the sharded bench line on n GPUs (torchrun, one rank per GPU)
set -u
TAG=${1:-x}
N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err || tail -8 gpurun_out/bench_n${N}_${TAG}.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_${TAG}.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['stage_ms'], d.get('transport'), d.get('batch_mode',{}).get('ms_per_step'))"
