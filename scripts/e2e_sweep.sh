# e2e vs sub-batch size of the pipelined align_batch: bash scripts/e2e_sweep.sh 131072 262144 ...
for sp in "$@"; do
  python bench.py --no-cpu-baseline --steps 2 --warmup 3 --split $sp 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('split $sp value ms', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1))
"
done
