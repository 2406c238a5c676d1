#!/bin/bash
# usage: tools/gpu_final.sh <tag> : parity tests + a short bench line
set -u
TAG=${1:-final}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-file-span > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['wall_ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['stage_ms'])"
