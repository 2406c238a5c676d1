set -u
mkdir -p gpurun_out
PM_AGG_MIN_READS=100 PM_COUNT_WARP_BELOW=0 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q -k "full_size" 2>&1 | tail -3
for t in 0 1; do
  PM_MISS_QUEUE=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-file-span 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('queue $t', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['stage_ms'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02f.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/ncu_launch_r02f.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02f.csv
