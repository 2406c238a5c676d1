set -u
mkdir -p gpurun_out
PM_RANK_MIN_READS=0 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
for t in 512 768; do
  PM_RANK_THREADS=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-file-span 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('threads $t', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'])"
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/plain_r02e.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:syncmers_rank|count_seeds" --launch-skip 4 -c 4 -f -o gpurun_out/prof_r02e python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-span > gpurun_out/ncu_full_r02e.log 2>&1
ls -la gpurun_out/prof_r02e.ncu-rep
