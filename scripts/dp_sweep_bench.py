#!/usr/bin/env python
"""BASELINE.json configs[4] — the DP-only sweep: kswcpp-equivalent banded DP on synthetic query/ref pairs.

  python scripts/dp_sweep_bench.py [--out profiles/dp_sweep.json] [--cpu]

For every point (length 100 .. 20 000, band 16 .. 512, mode global / extension / reversed extension; query = copy of
the target with 5 % divergence) it reports the GPU's GCUPS through ma_b200_ksw_upload/run
(device-resident inputs, CUDA-event time, all kswcpp_extz_t fields + CIGAR = the exact mode) and, for extensions,
also the extension-only mode (max / position / CIGAR only, early termination). The unit is the band cell the
reference processes (Σ_r en0 - st0 + 1 up to the row where it stops), so both modes are divided by the SAME cell
count: the oracle-defined cells of the exact computation. With --cpu the unmodified reference's kswcpp_dispatch
(oracle/_ref/ref_dump kswbench) is timed on the same pairs with 1 and all host threads.
Roofline: 148 SMs x 128 INT32 lanes x 1.965 GHz / 44 integer ops per cell = 846 GCUPS.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import dpgen  # noqa: E402
from ma_b200 import api  # noqa: E402

REF_DUMP = bench.REF_DUMP
MODES = {"global": dpgen.GLOBAL, "ext": dpgen.EXT, "ext_right": dpgen.EXT_RIGHT}


def write_pairs(path, pairs):
    with open(path, "w") as f:
        for w, zd, fl, q, t in pairs:
            f.write("%d %d %d %s %s\n" % (w, zd, fl, "".join(str(int(c)) for c in q) or "-",
                                           "".join(str(int(c)) for c in t) or "-"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "dp_sweep.json"))
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--cells", type=float, default=2e9, help="target band cells per point")
    args = ap.parse_args()
    ctx = api.Context(0)
    threads = os.cpu_count() or 1
    points = []
    for length in (100, 300, 1000, 3000, 10000, 20000):
        for w in (16, 32, 64, 128, 256, 512):
            if w > 2 * length:
                continue
            for mode in ("global", "ext", "ext_right"):
                base = dpgen.sweep_pairs(64, length, w, MODES[mode], 0.05, seed=length * 31 + w)
                per = sum(min(len(q), len(t), 2 * w + 1) * (len(q) + len(t)) for _, _, _, q, t in base) / len(base)
                # enough problems to fill the GPU twice over (148 SMs x 16 warps, one warp per problem), bounded total work
                reps = int(max(1, min(512, args.cells / per / len(base))))
                reps = max(reps, min(74, int(8e10 / per / len(base)) or 1))
                pairs = base * reps
                tasks, seq = api.pack_ksw_tasks(pairs)
                row = {"len": length, "w": w, "mode": mode, "tasks": len(pairs)}
                ctx.ksw_set_extension_only(False)
                ctx.ksw_upload(tasks, seq)
                ctx.ksw_run()
                ms = min(ctx.ksw_run() for _ in range(2))
                res, _ = ctx.ksw_download()
                cells = int(res["cells"].sum())
                row.update(cells=cells, ms=ms, gcups=cells / ms / 1e6, int16_mode=bool(length <= 1365))
                if mode != "global":
                    ctx.ksw_set_extension_only(True)
                    ctx.ksw_upload(tasks, seq)
                    ctx.ksw_run()
                    ms2 = min(ctx.ksw_run() for _ in range(2))
                    res2, _ = ctx.ksw_download()
                    ctx.ksw_set_extension_only(False)
                    row.update(ext_only_ms=ms2, ext_only_gcups_equiv=cells / ms2 / 1e6,
                               ext_only_cells_done=int(res2["cells"].sum()))
                if args.cpu and os.path.exists(REF_DUMP):
                    with tempfile.TemporaryDirectory() as d:
                        pf = os.path.join(d, "p.txt")
                        write_pairs(pf, base)
                        rep = max(1, int(2e8 / (cells / reps)))
                        for th in (1, threads):
                            o = bench.run_reference_ksw(pf, th, rep * (th if th > 1 else 1))  # bench.py's cpu_baseline leg
                            row["cpu_gcups_%dt" % th] = (cells / reps) * (o["calls"] / len(base)) / o["seconds"] / 1e9
                points.append(row)
                print(json.dumps(row), flush=True)
    out = {"metric": "DP GCUPS (band cells of the reference / s)", "roofline_gcups_at_44_ops_per_cell": 846.0,
           "host_threads": threads, "points": points}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
