"""Stand-alone timing / ncu target of the fused encoder stage-1 kernels at the training shape (G groups x 25 x [1,128,157])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import afsl_b200.ops as ops

dev = torch.device("cuda", 0)
g, grp, h, w = int(os.environ.get("S1_GROUPS", 64)), 25, 128, 157
x = torch.randn(g * grp, 1, h, w, device=dev)
conv, bn = torch.nn.Conv2d(1, 64, 3, padding=1).to(dev), torch.nn.BatchNorm2d(64).to(dev)
gy = None


def step():
    global gy
    y = ops.stage1_conv_bn_relu_pool(x, conv, bn, grp)
    if gy is None:
        gy = torch.randn_like(y)
    y.backward(gy)


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record()
torch.cuda.synchronize()
print(f"stage1 fwd+bwd, {g} groups x {grp}: {a.elapsed_time(b) / 5:.3f} ms per step")
