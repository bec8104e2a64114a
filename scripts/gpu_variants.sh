#!/bin/bash
# GPU box: the bench with the time of every DP launch, for the default library and the builds in ma_b200/variants/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_ksw_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -3
for lib in ma_b200/libma_b200.so ma_b200/variants/*.so; do
  echo "=== $lib"
  MA_B200_LIB=$PWD/$lib MA_B200_DP_BINS=1 python bench.py --steps 3 --warmup 3 2>gpurun_out/err_$(basename $lib).txt | tail -1 > gpurun_out/bench_$(basename $lib).json
  grep "dp bin" gpurun_out/err_$(basename $lib).txt | tail -12 | sort | uniq
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_$(basename $lib).json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels"]["ksw_kernels"])
PY
done
