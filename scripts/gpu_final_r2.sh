#!/bin/bash
# GPU box: the whole GPU suite, then the bench lines of the round's last state: configs[1], configs[3], configs[4] sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_full.json 2> gpurun_out/r2f_bench_full.err
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r2f_bench_config3_n1.json 2> gpurun_out/r2f_bench_config3.err
timeout 1500 python bench.py --config dp_sweep --steps 1 --warmup 1 > gpurun_out/r2f_dp_sweep.json 2> gpurun_out/r2f_dp_sweep.err
python - <<'PY'
import json
for f in ("r2f_bench_full", "r2f_bench_config3_n1", "r2f_dp_sweep"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"),
              (d.get("roofline") or {}).get("band_ge_128"), (d.get("cpu_baseline") or {}).get("value"),
              {k: round(v["ms"], 1) for k, v in (d.get("kernels") or {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
