#!/bin/bash
# GPU box: pipeline + boundary parity, then the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_pipeline_gpu.py tests/test_reference_boundary_gpu.py -q -m gpu 2>&1 | tail -12 | cut -c1-400
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -3 gpurun_out/bench_c1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], {k: round(v["ms"], 2) for k, v in d["kernels"].items()}, d["cpu_baseline"]["value"])
PY
