"""Per-opcode executed-instruction and stall-sample totals of one kernel from an ncu report (SASS source page).

    python tools/ncu_hot.py <file.ncu-rep> [top]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ops = defaultdict(lambda: [0, 0])
total_i = total_s = 0
for r in rows:
    src = r["Source"].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0]
    n = int(r["Instructions Executed"] or 0)
    s = int(r["Warp Stall Sampling (All Samples)"] or 0)
    ops[op][0] += n
    ops[op][1] += s
    total_i += n
    total_s += s
print(f"{path}: {total_i} warp instructions, {total_s} stall samples, {len(rows)} SASS lines")
print(f"{'opcode':12s} {'inst':>12s} {'inst%':>7s} {'samples%':>9s}")
for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{op:12s} {n:12d} {100 * n / total_i:7.2f} {100 * s / max(total_s, 1):9.2f}")
