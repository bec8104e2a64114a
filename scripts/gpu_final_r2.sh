#!/bin/bash
# GPU box: the whole GPU suite, then the evidence of the round's last state: configs[1] (ours + reference arm), launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_full.json 2> gpurun_out/r2g_bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2g_launches.log 2>&1
python - <<'PY'
import json
for f in ("r2g_bench_full", "r2g_bench_reference"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"),
              (d.get("roofline") or {}).get("frac"), d.get("gpu_launches"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
