"""One DP-only point (configs[4] grid): python scripts/dp_one.py <n_distinct> <repeat> <length> <band> <mode> <div>"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dpgen
from ma_b200 import api

n, rep, length, w = (int(x) for x in sys.argv[1:5])
mode = {"global": dpgen.GLOBAL, "ext": dpgen.EXT, "ext_right": dpgen.EXT_RIGHT}[sys.argv[5]]
div = float(sys.argv[6])
ctx = api.Context(0)
pairs = dpgen.sweep_pairs(n, length, w, mode, div, 8) * rep
tasks, seq = api.pack_ksw_tasks(pairs)
ctx.ksw_upload(tasks, seq)
ctx.ksw_run()
ms = min(ctx.ksw_run() for _ in range(2))
res, cig = ctx.ksw_download()
cells = int(res["cells"].sum())
print(json.dumps({"tasks": len(pairs), "len": length, "w": w, "ms": ms, "cells": cells, "gcups": cells / ms / 1e6}))
