"""Hot SASS regions of one kernel from an .ncu-rep source page: python scripts/ncu_hot.py rep launch_index [top]"""
import csv
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                              stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(raw.splitlines()))
print(rows[0][:2])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    if r[ix["Address"]] == "Address":
        break  # a second view of the same kernel follows
    data.append(r)
tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("total warp instructions", tot)
groups = []
for r in data:
    n = int(r[ix["Instructions Executed"]] or 0)
    t = int(r[ix["Thread Instructions Executed"]] or 0)
    a = r[ix["Address"]]
    if groups and abs(groups[-1]["n"] - n) <= 0.03 * max(n, 1):
        g = groups[-1]
        g["cnt"] += 1
        g["sum"] += n
        g["tsum"] += t
        g["end"] = a
    else:
        groups.append(dict(start=a, end=a, n=n, cnt=1, sum=n, tsum=t, first=r[ix["Source"]][:40]))
for g in sorted(groups, key=lambda g: -g["sum"])[:top]:
    print(g["start"][-5:], g["end"][-5:], "n=%d cnt=%d share=%.1f%% lanes=%.1f" % (g["n"], g["cnt"], 100 * g["sum"] / tot, g["tsum"] / max(g["sum"], 1)), g["first"])
