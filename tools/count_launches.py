import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from afsl_b200.episodes import EpisodeRunner, synthetic_batch
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = bench.build_model(dev)
opt = torch.optim.Adam(model.parameters(), lr=7e-4)
runner = EpisodeRunner(model, bench.EXPERIMENT_CONFIG, opt)
batch = synthetic_batch(32, 5, 5, 5, 157).to(dev)
for _ in range(3):
    runner.train_step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    runner.train_step(batch)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev)
small = [e for e in ev if e.device_time < 20]
print("kernels:", len(ev), "total us:", round(tot), "| kernels < 20 us:", len(small), "their us:", round(sum(e.device_time for e in small)))
from collections import Counter
c = Counter()
for e in small:
    c[e.name[:170]] += 1
for k, v in c.most_common(40):
    print(v, k)
big = Counter()
for e in ev:
    big[e.name[:90]] += e.device_time
print("--- top kernels by device time (us) ---")
for k, v in big.most_common(28):
    print(f"{v:9.0f}  {k}")
