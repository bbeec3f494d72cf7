"""Kernel-time breakdown of one training step (torch.profiler / CUPTI), for optimisation work.

    python tools/profile_step.py [--episodes E] [--cuda-profiler]   # --cuda-profiler: bracket one step for ncu
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from afsl_b200.episodes import EpisodeRunner, synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=8)
ap.add_argument("--cuda-profiler", action="store_true")
ap.add_argument("--rows", type=int, default=40)
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = bench.build_model(dev)
opt = torch.optim.Adam(model.parameters(), lr=7e-4)
runner = EpisodeRunner(model, bench.EXPERIMENT_CONFIG, opt)      # eager launches: the profiler attributes time per op
batch = synthetic_batch(args.episodes, 5, 5, 5, 157).to(dev)
for _ in range(3):
    runner.train_step(batch)
torch.cuda.synchronize()
if args.cuda_profiler:
    torch.cuda.cudart().cudaProfilerStart()
    runner.train_step(batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
else:
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        runner.train_step(batch)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=70))
