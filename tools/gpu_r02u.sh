#!/bin/bash
# usage: tools/gpu_r02u.sh <tag> <n> : multi-rank tests on n GPUs + the sharded bench line (torchrun, one rank per GPU)
set -u
TAG=${1:-r02u}
N=${2:-2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -x -q ) 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err || tail -8 gpurun_out/bench_n${N}_${TAG}.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_${TAG}.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['stage_ms'], d.get('transport'), d.get('batch_mode',{}).get('ms_per_step'))"
