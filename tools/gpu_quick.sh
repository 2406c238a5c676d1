# usage: tools/gpu_quick.sh <tag> : parity tests, bench (no CPU baseline), launch list
set -u
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_${TAG}.json').read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_packed',{}).get('ms_per_step'), d['kernel_ms'], d['stage_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv
