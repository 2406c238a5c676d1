#!/bin/bash
# usage: tools/gpu_r02t.sh <tag> : A/B of PM_SIDE_SCALARS (root_and_scalars beside node_deltas), parity tests, bench line
set -u
TAG=${1:-r02t}
mkdir -p gpurun_out
P="python tools/tune_probe.py 1000000 resident"

$P 2>&1 | tail -1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-file-span > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['wall_ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['stage_ms'])"
