#!/bin/bash
# GPU box: parity of the packed banded exact DP mode (ksw_bx.cuh), A/B against the scalar exact mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ksw_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -6 | cut -c1-400
echo "--- dp_quick bx"; timeout 300 python scripts/dp_quick.py 2>&1 | tail -10
if [ -n "$BX_AB" ]; then echo "--- dp_quick NO_BX=$BX_AB"; MA_B200_NO_BX=$BX_AB timeout 300 python scripts/dp_quick.py 2>&1 | tail -10; fi
for m in 0 $BX_AB; do
  echo "--- pacbio 3000 reads NO_BX=$m"
  MA_B200_NO_BX=$m MA_B200_DP_BINS=1 timeout 600 python bench.py --config 3 --long-reads 3000 --long-batch 3000 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/bx_pacbio_err_$m.txt | tail -1 > gpurun_out/bx_pacbio_$m.json
  grep "dp bin W" gpurun_out/bx_pacbio_err_$m.txt | awk -F'[:,]' '{k=$1; ms=$4; gsub(/ ms/,"",ms); if(!(k in best)||ms+0<best[k]+0){best[k]=ms; line[k]=$0}} END{for(k in line) print line[k]}' | sort
  python - <<PY
import json
d = json.load(open("gpurun_out/bx_pacbio_$m.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels"]["ksw_kernels"])
PY
done
