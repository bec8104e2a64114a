#!/bin/bash
# GPU box: bench.py on the three alignment configs (reduced sets for configs[2] / [3] unless FULL=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -3 gpurun_out/bench_c1.err
if [ "$FULL" = "1" ]; then P2=""; P3=""; else P2="--pairs 2000000"; P3="--long-reads 10000 --long-batch 5000"; fi
python bench.py --config 2 $P2 --steps 2 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python bench.py --config 3 $P3 --steps 2 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err
python bench.py --config 2 --impl reference --steps 1 --warmup 1 --cpu-sample 20000 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; tail -3 gpurun_out/bench_c2_ref.err
python - <<'PY'
import json
for c in ("c1", "c2", "c3", "c2_ref"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % c).read().strip().splitlines()[-1])
    except Exception as e:
        print(c, "no line", e); continue
    print(c, d.get("value"), d.get("ms_per_step"), d.get("e2e"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"),
          {k: round(v["ms"], 2) for k, v in (d.get("kernels") or {}).items()}, d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"), d.get("unavailable"))
PY
