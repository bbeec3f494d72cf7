"""Executed warp instructions and stall samples of one kernel per CUDA SOURCE LINE: joins the SASS page of an ncu report with
the line table of the object file (nvdisasm -g), instruction by instruction.

    python tools/ncu_lines.py <file.ncu-rep> <object.o> <kernel-name-substring> [top]
"""
import csv
import os
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    import glob
    cubin = glob.glob(tmp + "/*.cubin")[0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# instructions of the wanted function, in order, each with the (file, line) in force
table, cur, inside = [], ("?", 0), False
for l in sass:
    if l.startswith("//---") and ".text." in l:
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        table.append((int(m.group(1), 16), m.group(2).strip(), cur))
if len(table) != len(rows):
    print(f"warning: {len(table)} instructions in the object, {len(rows)} in the report (name filter?)")
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
ti = ts = 0
for (off, text, where), r in zip(table, rows):
    n = int(r["Instructions Executed"] or 0)
    s = int(r["Warp Stall Sampling (All Samples)"] or 0)
    a = agg[where]
    a[0] += n
    a[1] += s
    for k, v in r.items():
        if k.startswith("stall_") and "Not Issued" not in k and v and int(v):
            a[2][k] += int(v)
    ti += n
    ts += s
print(f"{rep}: {ti} warp instructions, {ts} stall samples")
print(f"{'file:line':32s} {'inst':>11s} {'inst%':>6s} {'smp%':>6s}  top stall reasons")
for where, (n, s, why) in sorted(agg.items(), key=lambda kv: -(kv[1][1]))[:top]:
    reasons = " ".join(f"{k[6:]}:{v}" for k, v in sorted(why.items(), key=lambda kv: -kv[1])[:3])
    print(f"{where[0] + ':' + str(where[1]):32s} {n:11d} {100 * n / max(ti, 1):6.2f} {100 * s / max(ts, 1):6.2f}  {reasons}")
