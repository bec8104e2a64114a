#!/bin/bash
# GPU box: DP + pipeline parity, then the bench with the time of every DP launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_ksw_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -5 | cut -c1-300
MA_B200_DP_BINS=1 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_qs.json
grep "dp bin" gpurun_out/bench_err.txt | tail -14 | sort | uniq
python - <<PY
import json
d = json.load(open("gpurun_out/bench_qs.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels"]["ksw_kernels"])
PY
