#!/bin/bash
# usage: tools/gpu_r02q.sh <tag> : A/B of PM_LIST_PREFETCH (+ parity under it), ncu of syncmers_rank with and without staging
set -u
TAG=${1:-r02q}
mkdir -p gpurun_out
P="python tools/tune_probe.py 1000000 resident"
$P 2>&1 | tail -1
PM_LIST_PREFETCH=1 $P 2>&1 | tail -1
PM_LIST_PREFETCH=1 PM_RANK_STAGE=1 $P 2>&1 | tail -1
( time PM_LIST_PREFETCH=1 PM_COUNT_WARP_BELOW=0 PM_AGG_MIN_READS=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) 2>&1 | tail -6
for v in 0 1; do
  PM_RANK_STAGE=$v ncu --set full --clock-control none --import-source on -k "regex:syncmers_rank" --launch-skip 3 -c 1 -f -o gpurun_out/prof_${TAG}_stage$v $P > gpurun_out/ncu_${TAG}_stage$v.log 2>&1
done
PM_LIST_PREFETCH=1 ncu --set full --clock-control none --import-source on -k "regex:count_seeds_lane" --launch-skip 3 -c 1 -f -o gpurun_out/prof_${TAG}_pf1 $P > gpurun_out/ncu_${TAG}_pf1.log 2>&1
ls -la gpurun_out/*.ncu-rep
