#!/bin/bash
# GPU box: ncu --set full of the six ksw_qs_kernel launches of one step (reduced batch), and the e2e line with --split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_qs.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ksw_qs_kernel -c 6 -o gpurun_out/prof_qs \
  python bench.py --pairs 250000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_qs.log 2>&1
tail -2 gpurun_out/ncu_qs.log | cut -c1-300
python bench.py --split 500000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split 500k', d['value'], d['ms_per_step'], d['e2e'])"
