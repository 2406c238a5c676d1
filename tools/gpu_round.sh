#!/bin/bash
# usage: tools/gpu_round.sh <tag> : the whole single-GPU evidence pass of a round -- parity tests, bench lines (c3 default, c1), ingest phase
# times, launch list, one --set full capture of the hot kernels.  Outputs under gpurun_out/.
set -u
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -20 gpurun_out/bench_${TAG}.err
python - <<EOF
import json
d=json.loads([l for l in open('gpurun_out/bench_${TAG}.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['stage_ms'], d['clocks'], d.get('spans'), d.get('index_open'), d['roofline']['frac'], d['roofline_scoring']['frac'])
print(d.get('cpu_baseline'))
EOF
timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/bench_c1_${TAG}.json 2> gpurun_out/bench_c1_${TAG}.err || tail -5 gpurun_out/bench_c1_${TAG}.err
PM_INGEST_TIMING=1 timeout 300 python tools/ingest_probe.py c3 2> gpurun_out/ingest_${TAG}.log; tail -12 gpurun_out/ingest_${TAG}.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv
ncu --set full --clock-control none --import-source on -k "regex:syncmers_rank|count_seeds_lane|node_deltas|prefix_scores|table_scan|entries_finalize|bfs_gather" --launch-skip 21 -c 7 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
