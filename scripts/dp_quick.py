"""Quick DP-only throughput probe (GCUPS) for the dominant Illumina extension shape and a few sweep points."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dpgen
from ma_b200 import api

ctx = api.Context(0)
rng = np.random.Generator(np.random.PCG64(1))
def illumina_like(n):
    pairs = []
    base = [rng.integers(0, 4, size=1100, dtype=np.uint8) for _ in range(64)]
    for i in range(n):
        t = base[i % 64][: int(950 + (i * 7) % 100)]
        ql = 20 + (i * 13) % 60
        q = t[:ql].copy()
        q[ql // 2] = (q[ql // 2] + 1) & 3
        pairs.append((512, 200, dpgen.EXT if i % 2 else dpgen.EXT_RIGHT, q, t))
    return pairs
for name, pairs in [("illumina_ext_50x1000_w512", illumina_like(100000)),
                    ("global_300_w64", dpgen.sweep_pairs(3000, 300, 64, dpgen.GLOBAL, 0.05, 3) * 10),
                    ("ext_3000_w256", dpgen.sweep_pairs(300, 3000, 256, dpgen.EXT, 0.05, 4) * 10),
                    ("global_1000_w128", dpgen.sweep_pairs(600, 1000, 128, dpgen.GLOBAL, 0.05, 5) * 10),
                    ("ext_10000_w512", dpgen.sweep_pairs(60, 10000, 512, dpgen.EXT, 0.12, 6) * 20),
                    ("ext_20000_w32", dpgen.sweep_pairs(60, 20000, 32, dpgen.EXT_RIGHT, 0.05, 7) * 20),
                    ("ext_3000_w16", dpgen.sweep_pairs(500, 3000, 16, dpgen.EXT, 0.05, 8) * 20),
                    ("global_3000_w32", dpgen.sweep_pairs(500, 3000, 32, dpgen.GLOBAL, 0.05, 9) * 20),
                    ("ext_3000_w64", dpgen.sweep_pairs(500, 3000, 64, dpgen.EXT_RIGHT, 0.05, 10) * 20),
                    ("global_3000_w128", dpgen.sweep_pairs(400, 3000, 128, dpgen.GLOBAL, 0.05, 11) * 20)]:
    tasks, seq = api.pack_ksw_tasks(pairs)
    ctx.ksw_upload(tasks, seq)
    ctx.ksw_run()
    ms = min(ctx.ksw_run() for _ in range(3))
    res, cig = ctx.ksw_download()
    cells = int(res["cells"].sum())
    print(json.dumps({"case": name, "tasks": len(pairs), "ms": ms, "cells": cells, "gcups": cells / ms / 1e6}))
