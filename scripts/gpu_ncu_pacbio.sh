#!/bin/bash
# GPU box: ncu --set full with source counters of the W=1024 / W=512 DP launches of a small PacBio batch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_pacbio.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ksw_batch_kernel" -s 10 -c 10 -o gpurun_out/prof_pacbio \
  python bench.py --config 3 --long-reads 3000 --long-batch 3000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pacbio.log 2>&1
tail -2 gpurun_out/ncu_pacbio.log | cut -c1-200; ls -la gpurun_out/prof_pacbio.ncu-rep
