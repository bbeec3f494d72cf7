import os, sys
sys.path.insert(0, '/root/repo')
import torch, bench
from afsl_b200._lib import call, ptr, stream_ptr
dev = torch.device("cuda", 0); st = stream_ptr()
for w_, k_, d_ in ((20, 5, 256), (20, 1, 256), (20, 5, 64)):
    ns_, nq_ = w_ * k_, w_ * 5
    e_ = 148 * 40
    s_ = torch.randn(e_, ns_, d_, device=dev); q_ = torch.randn(e_, nq_, d_, device=dev)
    sl_ = torch.arange(w_, device=dev, dtype=torch.int32).repeat_interleave(k_).expand(e_, -1).contiguous()
    ql_ = torch.arange(w_, device=dev, dtype=torch.int32).repeat_interleave(5).expand(e_, -1).contiguous()
    pred_ = torch.empty(e_ * nq_, device=dev, dtype=torch.int32); post_ = torch.empty(e_ * nq_, device=dev)
    corr_ = torch.empty(e_, device=dev, dtype=torch.int32)
    fn = lambda: call("afsl_proto_head_fwd_f32", ptr(s_), ptr(sl_), ptr(q_), ptr(ql_), None, None, None, None, ptr(pred_),
                      ptr(post_), ptr(corr_), e_, ns_, nq_, w_, d_, st)
    fn(); torch.cuda.synchronize()
    os.environ["AFSL_HEAD_DBG"] = "1"
    print(f"=== {w_}w{k_}s D={d_}", file=sys.stderr, flush=True)
    fn(); torch.cuda.synchronize()
    del os.environ["AFSL_HEAD_DBG"]
