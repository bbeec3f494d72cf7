"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, total time, share.

    python tools/launch_shares.py gpurun_out/<tag>_launches.csv [--skip N] > profiles/<tag>_launch_shares.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)                       # drop the argument list
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    name = name.replace("void ", "")
    return name[:110]


def main(path, skip=0):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        a = agg[short(name)]
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("afsl::") or "_kernel<" in k and "afsl" in k)
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms of kernel time (serialised, cold cache); "
          f"libafsl kernels {100 * ours / total:.1f} %")
    print(f"{'share%':>7s} {'ms':>10s} {'launches':>9s} {'avg_us':>9s}  kernel")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * ns / total:7.2f} {ns / 1e6:10.3f} {n:9d} {ns / n / 1e3:9.1f}  {k}")


if __name__ == "__main__":
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    main(sys.argv[1], skip)
