#!/bin/bash
# GPU box (debugging aid): one early-stop extension problem with the debug builds in ma_b200/variants/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in ma_b200/variants/*.so; do
echo "=== $lib"
MA_B200_LIB=$PWD/$lib python - 2>&1 <<'PY' | grep "QSDBG\|RESULT" | sort | uniq -c | head -30
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_ksw_gpu as T
from ma_b200 import api
ctx = api.Context(0)
pairs = T._ext_pairs(400, 77, 150, 950, 1060, 0.03)
for sel in (pairs[8:9],):
    tasks, seq = api.pack_ksw_tasks(sel)
    ctx.ksw_set_extension_only(True)
    res, cig = ctx.ksw_batch(tasks, seq)
    import torch; torch.cuda.synchronize()
    print("RESULT", {k: int(res[k][0]) for k in res.dtype.names}, flush=True)
PY
done
