cd /root/repo
python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-200
echo "one kernel:"; MA_B200_SOC_ONE_KERNEL=1 python bench.py --pairs 500000 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), {k: round(v['ms'],2) for k,v in d['kernels'].items()})"
bash scripts/gpu_variants.sh --pairs 500000 --steps 3 --warmup 2
