"""ncu target: the grouped BatchNorm+ReLU+MaxPool kernels on a stage-1 shaped tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from afsl_b200._lib import call, ptr, stream_ptr
dev = torch.device("cuda", 0)
g, grp, c, h, w = 16, 25, 64, 128, 157
x = torch.randn(g * grp, c, h, w, device=dev)
mean, rstd, var = (torch.empty(g, c, device=dev) for _ in range(3))
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
y = torch.empty(g * grp, c, h // 3, w // 3, device=dev)
dy = torch.randn_like(y); dx = torch.empty_like(x); sums = torch.empty(g, c, 2, device=dev)
st = stream_ptr()
for _ in range(3):
    call("afsl_gbn_stats_f32", ptr(x), ptr(mean), ptr(rstd), ptr(var), g, grp, c, h, w, 1e-5, st)
    call("afsl_gbn_relu_pool_fwd_f32", ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(y), g, grp, c, h, w, 1, st)
    call("afsl_gbn_relu_pool_bwd_f32", ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(dy), ptr(dx), ptr(sums), g, grp, c, h, w, 1, st)
torch.cuda.synchronize()
print("ok")
