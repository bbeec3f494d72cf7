"""Time the many-way evaluation head (config 5 shapes) through the C ABI: TMA-fed tcgen05 kernel (default), its LDG-fed
predecessor (AFSL_HEAD_MMA=2) and the fp32-pipe kernel (AFSL_HEAD_MMA=0).  Also the ncu target for these kernels.

    python tools/head_many_way_bench.py [variant ...]        # variants: 1 2 0 (default: all three)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from afsl_b200._lib import call, ptr, stream_ptr

dev = torch.device("cuda", 0)
pk, _ = bench.peaks()
st = stream_ptr()
variants = sys.argv[1:] or ["1", "2", "0"]
out = {}
for w_, k_, d_ in ((20, 5, 256), (20, 1, 256), (20, 5, 128), (20, 5, 64), (20, 1, 64)):
    ns_, nq_ = w_ * k_, w_ * 5
    e_ = min(65536, max(2048, int(1.5e9 // (4 * d_ * (ns_ + nq_)))))
    s_ = torch.randn(e_, ns_, d_, device=dev)
    q_ = torch.randn(e_, nq_, d_, device=dev)
    sl_ = torch.arange(w_, device=dev, dtype=torch.int32).repeat_interleave(k_).expand(e_, -1).contiguous()
    ql_ = torch.arange(w_, device=dev, dtype=torch.int32).repeat_interleave(5).expand(e_, -1).contiguous()
    pred_ = torch.empty(e_ * nq_, device=dev, dtype=torch.int32)
    post_ = torch.empty(e_ * nq_, device=dev)
    corr_ = torch.empty(e_, device=dev, dtype=torch.int32)
    nbytes = (4.0 * d_ * (ns_ + nq_) + 4 * (ns_ + nq_) + 8 * nq_ + 4) * e_
    for v in variants:
        os.environ["AFSL_HEAD_MMA"] = v
        fn = lambda: call("afsl_proto_head_fwd_f32", ptr(s_), ptr(sl_), ptr(q_), ptr(ql_), None, None, None, None, ptr(pred_),
                          ptr(post_), ptr(corr_), e_, ns_, nq_, w_, d_, st)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        frac = nbytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]
        out[f"{w_}w{k_}s_d{d_}_v{v}"] = {"ms": ms, "frac": frac, "tasks": e_, "us_per_task_per_sm": ms * 1e3 * 148 / e_}
        print(f"{w_}w{k_}s D={d_} variant {v}: {ms:8.3f} ms  frac {frac:.3f}  ({e_} tasks, {ms * 1e3 * 148 / e_:.2f} us per task per SM)")
    del s_, q_
print(json.dumps(out))
