import sys; sys.path.insert(0, "/root/repo")
import torch
from afsl_b200._lib import call, ptr, stream_ptr
dev = torch.device("cuda", 0)
g, grp, h, w = 64, 25, 128, 157
n = g * grp
x = torch.randn(n, 1, h, w, device=dev)
w9 = torch.randn(64, 9, device=dev) * 0.3
a, b = torch.rand(g, 64, device=dev) + 0.5, torch.randn(g, 64, device=dev) * 0.1
y = torch.empty(n, 42, 52, 64, device=dev)
arg = torch.empty(n, 42, 52, 64, device=dev, dtype=torch.uint8)
st = stream_ptr()
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
print("codes   ", timed(lambda: call("afsl_stage1_fwd_f32", ptr(x), ptr(w9), ptr(a), ptr(b), ptr(y), ptr(arg), g, grp, h, w, 1, 1, st)))
print("no codes", timed(lambda: call("afsl_stage1_fwd_f32", ptr(x), ptr(w9), ptr(a), ptr(b), ptr(y), None, g, grp, h, w, 1, 1, st)))
