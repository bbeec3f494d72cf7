"""Write profiles/ncu_traffic.json: DRAM bytes (read + write) per launch of the kernels bench.py times,
taken from `ncu --set full` reports of tools/kernels_bench.py.

    python tools/ncu_traffic.py name=report.ncu-rep [name=a.ncu-rep+b.ncu-rep ...]
A `+` list sums several kernels into one entry (e.g. the forward/backward pair).
"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    total = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(key)
        total += float(vals[i].replace(",", "")) * UNIT[units[i]]
    return total, vals[hdr.index("Kernel Name")]


def main(args):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_path = os.path.join(root, "profiles", "ncu_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    src = out.setdefault("_source", {})
    for a in args:
        name, files = a.split("=")
        total, kernels = 0.0, []
        for f in files.split("+"):
            b, k = dram_bytes(f)
            total += b
            kernels.append(f"{os.path.basename(f)}: {k}")
        out[name] = total
        src[name] = kernels
    with open(out_path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: v for k, v in out.items() if not k.startswith('_')}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
