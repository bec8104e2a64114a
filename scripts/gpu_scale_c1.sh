#!/bin/bash
# GPU box with N GPUs: the driver's line (configs[1], weak scaling) under torchrun
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_c1_n$N.json 2> gpurun_out/r2g_bench_c1_n$N.err
tail -2 gpurun_out/r2g_bench_c1_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2g_bench_c1_n$N.json").read().strip().splitlines()[-1])
    print("configs[1] N=$N", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["scaling"], d["clocks"])
except Exception as e:
    print("no line", e)
PY
