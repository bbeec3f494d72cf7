"""Run the stand-alone kernel roofline measurements of bench.py (also the ncu target for single kernels)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
pk, src = bench.peaks()
out = bench.kernel_rooflines(dev, pk["hbm_gbs"], 32)
for k, v in out.items():
    print(f"{k:22s} {v['ms']:8.3f} ms  {v['achieved']:8.1f} GB/s  frac {v['frac']:.3f}  ({v['units']})")
print(json.dumps(out))
