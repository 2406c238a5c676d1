python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['stage_ms'])"
python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -2
