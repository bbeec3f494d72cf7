import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from afsl_b200._lib import call, ptr, stream_ptr
pk = bench.peaks()[0]["hbm_gbs"]
dev = torch.device("cuda", 0); st = stream_ptr()
e, ways, nq, d = 16384, 5, 25, 256
p = torch.randn(e, ways, d, device=dev); q = torch.randn(e, nq, d, device=dev)
ql = torch.arange(ways, dtype=torch.int32, device=dev).repeat_interleave(5).expand(e, -1).contiguous()
loss = torch.empty(e, device=dev); dl = torch.full((e,), 1.0 / e, device=dev)
dp, dq = torch.empty_like(p), torch.empty_like(q)
sim, qinv = torch.empty(e, ways, nq, device=dev), torch.empty(e, nq, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / reps
cf = lambda: call("afsl_cpl_fwd_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(loss), e, nq, ways, d, st)
cb = lambda: call("afsl_cpl_bwd_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(dl), ptr(dp), ptr(dq), e, nq, ways, d, st)
cfs = lambda: call("afsl_cpl_fwd_save_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(loss), ptr(sim), ptr(qinv), e, nq, ways, d, st)
cbs = lambda: call("afsl_cpl_bwd_saved_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(sim), ptr(qinv), ptr(dl), ptr(dp), ptr(dq), e, nq, ways, d, st)
bf = (4.0 * d * (nq + ways) + 4 * nq + 4) * e; bb = (4.0 * d * 2 * (nq + ways) + 4 * nq + 4) * e
for name, f, b in (("recompute", cf, cb), ("saved", cfs, cbs)):
    tf, tb = timed(f), timed(b)
    print(f"{name:10s} fwd {tf:.4f} ms frac {bf/tf/1e6/pk:.3f}  bwd {tb:.4f} ms frac {bb/tb/1e6/pk:.3f}  pair {tf+tb:.4f} frac {(bf+bb)/(tf+tb)/1e6/pk:.3f}")
