#!/bin/bash
# usage: tools/gpu_r02v.sh : per-kernel times of the sharded data plane at 2 and 8 ranks (all ranks on one GPU, ncu launch lists)
set -u
mkdir -p gpurun_out
for n in 2 8; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_multi${n}_r02v.csv python tools/multi_probe.py $n 3 > gpurun_out/multi${n}_r02v.log 2>&1
  tail -2 gpurun_out/multi${n}_r02v.log
  python tools/launch_summary.py gpurun_out/launches_multi${n}_r02v.csv
done
