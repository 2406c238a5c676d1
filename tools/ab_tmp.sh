set -u
for v in "4 8" "3 6" "2 4"; do
  set -- $v
  echo "== PM_FIT_LO=$1 PM_FIT_HI=$2"
  PM_FIT_LO=$1 PM_FIT_HI=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['stage_ms']['table finalize'])"
done
