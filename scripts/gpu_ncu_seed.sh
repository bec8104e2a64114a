#!/bin/bash
# GPU box: ncu --set full (with source counters) of seed_kernel / locate_kernel / socharm_kernel on a reduced batch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_seed.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"seed_kernel|locate_kernel|socharm_kernel" -c 3 -o gpurun_out/prof_seed \
  python bench.py --pairs 250000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seed.log 2>&1
tail -2 gpurun_out/ncu_seed.log | cut -c1-300
