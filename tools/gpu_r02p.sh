#!/bin/bash
# usage: tools/gpu_r02p.sh <tag> : parity tests (default build, then with the staged syncmers_rank on every s = 8 launch), default bench,
# A/B probes of the experiment switches (PM_RANK_STAGE, PM_RESIDENT_SLICES, PM_TABLE_FIT_LO)
set -u
TAG=${1:-r02p}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
( time PM_RANK_STAGE=1 PM_RANK_MIN_READS=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['wall_ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['kernel_ms'], d['stage_ms'])"
P="python tools/tune_probe.py 1000000 resident"
$P 2>&1 | tail -1
PM_RANK_STAGE=1 $P 2>&1 | tail -1
PM_TABLE_FIT_LO=3 $P 2>&1 | tail -1
for s in 2 3; do PM_RESIDENT_SLICES=$s $P 2>&1 | tail -1; done
for s in 4 8; do PM_RESIDENT_SLICES=$s PM_AGG_MIN_READS=60000 PM_COUNT_WARP_BELOW=50000 $P 2>&1 | tail -1; done
PM_RANK_STAGE=1 PM_RESIDENT_SLICES=4 PM_AGG_MIN_READS=60000 PM_COUNT_WARP_BELOW=50000 $P 2>&1 | tail -1
PM_RANK_STAGE=1 python tools/tune_probe.py 1000000 e2e 2>&1 | tail -2
