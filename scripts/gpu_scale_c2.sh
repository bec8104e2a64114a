#!/bin/bash
# GPU box with N GPUs: bench.py --config 2 (the fixed 10 M-pair set split over the ranks) and --config 3 under torchrun
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
for c in 2 3; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/scale_c${c}_n$N.json 2> gpurun_out/scale_c${c}_n$N.err
tail -2 gpurun_out/scale_c${c}_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_c${c}_n$N.json").read().strip().splitlines()[-1])
    print("config $c N=$N", d["value"], d["ms_per_step"], d["e2e"], d["scaling"], d["config"]["reads_per_step"], d["config"]["reads_per_step_per_gpu"], d["clocks"])
except Exception as e:
    print("no line", e)
PY
done
