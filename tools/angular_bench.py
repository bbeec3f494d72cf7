"""Time the angular-loss kernels (config 3: prototypes as anchors, Dp = 64, 5-way, 25 queries) through the C ABI:
tensor-core kernel (AFSL_ANGULAR_TC=1, default) against the warp-per-episode fp32 kernel (AFSL_ANGULAR_TC=0).
Also the ncu target for these kernels.

    python tools/angular_bench.py [tc|warp ...] [--angle DEG] [--episodes E]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from afsl_b200._lib import call, ptr, stream_ptr

args = sys.argv[1:]
angle, e = 0.0, 16384
if "--angle" in args:
    i = args.index("--angle"); angle = float(args[i + 1]); del args[i:i + 2]
if "--episodes" in args:
    i = args.index("--episodes"); e = int(args[i + 1]); del args[i:i + 2]
timeline = "--timeline" in args
if timeline:
    args.remove("--timeline")
variants = args or ["tc", "warp"]
dev = torch.device("cuda", 0)
pk = bench.peaks()[0]["hbm_gbs"]
st = stream_ptr()
ways, nq, d = 5, 25, 64
gen = torch.Generator(device="cpu").manual_seed(0)
pa = torch.nn.functional.normalize(torch.randn(e, ways, d, generator=gen), dim=-1).to(dev)
qa = torch.nn.functional.normalize(torch.randn(e, nq, d, generator=gen), dim=-1).to(dev)
ql = torch.arange(ways, dtype=torch.int32).repeat_interleave(nq // ways).expand(e, -1).contiguous().to(dev)
loss = torch.empty(e, device=dev)
dl = torch.full((e,), 1.0 / e, device=dev)
dpa, dqa = torch.empty_like(pa), torch.empty_like(qa)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


for v in variants:
    os.environ["AFSL_ANGULAR_TC"] = "1" if v == "tc" else "0"
    if timeline and v == "tc":          # one instrumented launch of each kernel: CTA 0's pipeline events on stderr
        os.environ["AFSL_ANGULAR_DBG"] = "1"
        call("afsl_angular_fwd_f32", ptr(pa), ptr(qa), ptr(ql), angle, 40.0, 1, 0, ptr(loss), e, nq, ways, d, st)
        call("afsl_angular_bwd_f32", ptr(pa), ptr(qa), ptr(ql), angle, 40.0, 1, 0, ptr(dl), ptr(dpa), ptr(dqa), e, nq, ways, d, st)
        del os.environ["AFSL_ANGULAR_DBG"]
    af = lambda: call("afsl_angular_fwd_f32", ptr(pa), ptr(qa), ptr(ql), angle, 40.0, 1, 0, ptr(loss), e, nq, ways, d, st)
    ab = lambda: call("afsl_angular_bwd_f32", ptr(pa), ptr(qa), ptr(ql), angle, 40.0, 1, 0, ptr(dl), ptr(dpa), ptr(dqa), e, nq,
                      ways, d, st)
    tf, tb = timed(af), timed(ab)
    bf = (4.0 * d * (nq + ways) + 4 * nq + 4) * e
    bb = (4.0 * d * 2 * (nq + ways) + 4 * nq + 4) * e
    bp = 4.0 * d * 3 * (nq + ways) * e
    print(f"{v:5s} angle {angle:4.1f}  fwd {tf:7.4f} ms frac {bf / tf / 1e6 / pk:5.3f}   bwd {tb:7.4f} ms frac {bb / tb / 1e6 / pk:5.3f}"
          f"   pair {tf + tb:7.4f} ms frac {bp / (tf + tb) / 1e6 / pk:5.3f}   ({e} episodes, loss[0] = {float(loss[0]):.6f})", flush=True)
