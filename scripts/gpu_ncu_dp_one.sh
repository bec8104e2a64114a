#!/bin/bash
# GPU box: ncu --set full with source counters of ksw_batch_kernel on one DP-only point; args as scripts/dp_one.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_dp_one.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ksw_batch_kernel" -s 1 -c 1 -o gpurun_out/prof_dp_one \
  python scripts/dp_one.py "$@" > gpurun_out/ncu_dp_one.log 2>&1
tail -1 gpurun_out/ncu_dp_one.log | cut -c1-200; ls -la gpurun_out/prof_dp_one.ncu-rep
