#!/bin/bash
# usage: tools/gpu_r02s.sh <tag> : evidence pass -- parity tests, default bench line, c1 bench line (with index_build), launch list, ncu --set full of the hot kernels
set -u
TAG=${1:-r02s}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/bench_c1_${TAG}.json 2> gpurun_out/bench_c1_${TAG}.err || tail -5 gpurun_out/bench_c1_${TAG}.err
python -c "
import json
for f in ('gpurun_out/bench_${TAG}.json','gpurun_out/bench_c1_${TAG}.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['wall_ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['kernel_ms'], d.get('index_build'), d.get('cpu_baseline',{}).get('gpu_same_sample'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv
ncu --set full --clock-control none --import-source on -k "regex:syncmers_rank|count_seeds_lane|node_deltas|prefix_scores|table_scan|entries_finalize|root_and_scalars" --launch-skip 30 -c 7 -f -o gpurun_out/prof_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --set full --clock-control none -k "regex:genome_materialize|seeds_sort|node_diff" -c 6 -f -o gpurun_out/prof_${TAG}_build python tools/build_probe.py 0 > gpurun_out/ncu_build_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep
