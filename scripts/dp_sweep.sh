# A/B of compile-time knobs on a GPU box: bash scripts/dp_sweep.sh "<nvcc -D flags>" ...
B="python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline"
for cfg in "$@"; do
  make -s -B -C ma_b200/csrc EXTRA="$cfg" > /dev/null 2>&1
  $B 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$cfg', 'ms/step',round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items()})
"
done
