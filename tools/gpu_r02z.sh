#!/bin/bash
# usage: tools/gpu_r02z.sh <tag> : parity tests, default bench line (stage timers off in the timed loop), launch list
set -u
TAG=${1:-r02z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['wall_ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['stage_ms'], d['kernel_ms'], d['roofline']['frac'], d['roofline_scoring']['frac'], d['gpu_launches'], d.get('cpu_baseline',{}).get('gpu_same_sample'))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err || tail -5 gpurun_out/bench_ref_${TAG}.err
tail -c 600 gpurun_out/bench_ref_${TAG}.json
