import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, helpers as H
from ma_b200 import api
g = np.load(os.path.join(H.GOLDEN, "ksw_golden.npz"))
d = {"ksw_calls": g["calls"].astype(np.int64), "ksw_seq": g["seq"], "ksw_cigar": g["cigar"]}
calls = list(H.split_ksw_dump(d))
ctx = api.Context(0)
tasks, seq = api.pack_ksw_tasks([(f["w"], f["zdrop"], f["flag"], q, t) for f, q, t, c in calls])
res, cig = ctx.ksw_batch(tasks, seq)
FIELDS = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end"]
bad = 0
for i, (f, q, t, c) in enumerate(calls):
    diffs = [(k, int(res[k][i]), f[k]) for k in FIELDS if int(res[k][i]) != f[k]]
    got = cig[res["cigar_off"][i]:res["cigar_off"][i] + res["n_cigar"][i]]
    if diffs or not np.array_equal(got, c):
        bad += 1
        if bad <= 12:
            print(i, "qlen", f["qlen"], "tlen", f["tlen"], "w", f["w"], "zd", f["zdrop"], "flag", hex(f["flag"]), diffs[:6], "cigar_eq", np.array_equal(got, c))
print("bad", bad, "of", len(calls))
