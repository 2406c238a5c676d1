#!/bin/bash
# usage: tools/gpu_check.sh <tag> [kernel-regex for the ncu --set full capture]
# GPU-box sequence: parity tests, plain bench, launch list (ncu time-only pass), one full capture of the named kernels.
set -u
TAG=${1:-x}
KRE=${2:-node_deltas|prefix_scores}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || { tail -20 gpurun_out/bench_${TAG}.err; exit 1; }
cat gpurun_out/bench_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv
ncu --set full --clock-control none --import-source on -k "regex:${KRE}" --launch-skip 6 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
