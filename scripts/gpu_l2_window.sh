cd /root/repo
for w in 0 100 60; do
echo "== MA_B200_L2_WINDOW=$w"
MA_B200_L2_WINDOW=$w python bench.py --pairs 500000 --steps 3 --warmup 2 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), {k: round(v['ms'],2) for k,v in d['kernels'].items()})"
grep "L2 window" /tmp/err.txt | head -1
done
echo "== config 2 (2 M pairs)"
for w in 0 30; do
MA_B200_L2_WINDOW=$w python bench.py --config 2 --pairs 1000000 --steps 2 --warmup 2 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), {k: round(v['ms'],2) for k,v in d['kernels'].items()})"
grep "L2 window" /tmp/err.txt | head -1
done
