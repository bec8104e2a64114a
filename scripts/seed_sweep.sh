# A/B sweep of the seeding kernel's tile parameters on a GPU box: bash scripts/seed_sweep.sh "K BLOCK MINB" ...
B="python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline"
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'ms/step',round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items()})
"; }
for cfg in "$@"; do set -- $cfg
  make -s -B -C ma_b200/csrc EXTRA="-DMA_SEED_K=$1 -DMA_SEED_BLOCK=$2 -DMA_SEED_MINB=$3" > /dev/null 2>&1
  $B 2>&1 | summ K$1_B$2_MB$3
done
