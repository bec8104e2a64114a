#!/bin/bash
# GPU box: the evidence bundle of the round — bench lines, launch list, ncu --set full summaries of the kernels of a step
# (the .ncu-rep of a full-size step is too large to bring back: it is summarised on the box and removed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != "ncu" ]; then
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
MA_B200_DP_BINS=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2> gpurun_out/r2_dp_bins.err > /dev/null
grep "dp bin" gpurun_out/r2_dp_bins.err | sort | uniq > gpurun_out/r2_dp_bins.txt
MA_B200_DP_BINS=1 python bench.py --config 2 --pairs 2000000 --steps 1 --warmup 2 --no-cpu-baseline 2> gpurun_out/r2_dp_bins_c2.err > /dev/null
grep "dp bin" gpurun_out/r2_dp_bins_c2.err | sort | uniq > gpurun_out/r2_dp_bins_config2.txt
# launch list of the same command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
python - <<'PY'
import json
for f in ("r2_bench_full", "r2_bench_reference"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e"), (d.get("roofline") or {}).get("frac"))
PY
cat gpurun_out/r2_dp_bins_config2.txt
else
# ncu --set full at full size: every kernel of the second step (the slabs are sized by then)
timeout 1500 ncu --set full --clock-control none \
  -k regex:"seed_kernel|locate_kernel|socharm_kernel|ksw_qs_kernel|ksw_tiny_kernel|ksw_batch_kernel|nwasm_kernel|mapq_kernel|pair_kernel" \
  -s 17 -c 17 -o /tmp/r2_full_size python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1
python scripts/ncu_summary.py /tmp/r2_full_size.ncu-rep gpurun_out/r2_full_size_ncu.txt > /dev/null
ls -la /tmp/r2_full_size.ncu-rep; wc -l gpurun_out/r2_full_size_ncu.txt
# the human-sized configuration (2 M reads): seeding / locate against the HBM-resident table
timeout 900 ncu --set full --clock-control none -k regex:"seed_kernel|locate_kernel" -s 2 -c 2 -o /tmp/r2_config2 \
  python bench.py --config 2 --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_config2.log 2>&1
python scripts/ncu_summary.py /tmp/r2_config2.ncu-rep gpurun_out/r2_config2_seed_locate_ncu.txt > /dev/null
wc -l gpurun_out/r2_config2_seed_locate_ncu.txt
fi
