#!/bin/bash
# GPU box: a reduced bench for the default library and the builds in ma_b200/variants/ ; args: extra bench.py options
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in ma_b200/libma_b200.so ma_b200/variants/*.so; do
  MA_B200_LIB=$PWD/$lib python bench.py "$@" --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$(basename $lib)', round(d['ms_per_step'],2), {k: round(v['ms'],2) for k,v in d['kernels'].items()})"
done
