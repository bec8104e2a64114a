"""Warp instructions per CUDA source line of one kernel launch: joins the SASS page of an .ncu-rep with the line table of
the cubin (nvdisasm -g).  python scripts/ncu_lines.py rep launch_index lib.so kernel_symbol_substring [top]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, skip, lib, sym = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                              stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    if r[ix["Address"]] == "Address":
        break
    sass.append((int(r[ix["Instructions Executed"]] or 0), int(r[ix["Thread Instructions Executed"]] or 0), r[ix["Source"]]))
with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(d, cub)], stderr=subprocess.DEVNULL).decode()
lines, on, cur = [], False, ("?", 0)
for l in dis.splitlines():
    if l.startswith(".text."):
        on = sym in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
agg = {}
for (n, t, _), key in zip(sass, lines):
    a = agg.setdefault(key, [0, 0])
    a[0] += n
    a[1] += t
tot = sum(v[0] for v in agg.values())
print("total warp instructions", tot)
for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %5d  %5.1f%%  lanes %.1f" % (key[0], key[1], 100.0 * n / tot, t / max(n, 1)))
