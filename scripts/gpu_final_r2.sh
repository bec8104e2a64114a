#!/bin/bash
# GPU box: bench lines after the packed banded DP modes (ksw_bx.cuh / ksw_bn.cuh): configs[4] sweep, configs[3], configs[1]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python bench.py --config dp_sweep --steps 1 --warmup 1 > gpurun_out/r2e_dp_sweep_line.json 2> gpurun_out/r2e_dp_sweep.err
ls -la gpurun_out/*.json | tail -3
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r2e_bench_config3_n1.json 2> gpurun_out/r2e_bench_config3.err
python - <<'PY'
import json
for f in ("r2e_dp_sweep_line", "r2e_bench_config3_n1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"),
              (d.get("roofline") or {}).get("band_ge_128"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
