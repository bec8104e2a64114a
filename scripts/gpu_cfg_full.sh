#!/bin/bash
# GPU box: bench.py on BASELINE configs[2] and configs[3] at full size (N = 1) + the reference arm on configs[2]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --config 2 --steps 2 --warmup 1 > gpurun_out/bench_c2_full.json 2> gpurun_out/bench_c2_full.err; tail -2 gpurun_out/bench_c2_full.err
python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/bench_c3_full.json 2> gpurun_out/bench_c3_full.err; tail -2 gpurun_out/bench_c3_full.err
python bench.py --config 2 --impl reference --steps 1 --warmup 1 > gpurun_out/bench_c2_full_ref.json 2> gpurun_out/bench_c2_full_ref.err; tail -2 gpurun_out/bench_c2_full_ref.err
python bench.py --config 3 --impl reference --steps 1 --warmup 1 > gpurun_out/bench_c3_full_ref.json 2> gpurun_out/bench_c3_full_ref.err; tail -2 gpurun_out/bench_c3_full_ref.err
python - <<'PY'
import json
for c in ("c2_full", "c3_full", "c2_full_ref", "c3_full_ref"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % c).read().strip().splitlines()[-1])
    except Exception as e:
        print(c, "no line", e); continue
    print(c, d.get("value"), d.get("ms_per_step"), d.get("e2e"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"),
          {k: round(v["ms"], 2) for k, v in (d.get("kernels") or {}).items()}, d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"), d.get("unavailable"), d.get("workload_generation_s"))
PY
