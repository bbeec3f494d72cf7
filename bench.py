#!/usr/bin/env python
"""Benchmark of the episodic prototypical-network hot path (BASELINE.json metric: episodes/s, train fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--episodes E] [--impl b200|reference]

Workload = BASELINE.json configs[1]: Hybrid encoder + self-attention view fusion + CPL loss
(M=5, T=9.2361, l=2.022308) with SpecAugment support and query views, FSD2018-shaped 5-way 5-shot
5-query episodes of [1,128,157] synthetic N(0,1) log-mel spectrograms, random-init weights.
A step = one optimizer step over E episodes per GPU (SpecAugment views -> encoder -> view fusion ->
fused prototype head + CPL -> backward -> Adam).  ``value`` is measured with inputs resident in HBM,
``e2e`` through EpisodeRunner.train_step with pinned HOST inputs (H2D copy of the spectrograms and D2H
read of the loss inside the timed region).  N > 1: one process per GPU (torchrun), episodes sharded
across ranks (weak scaling), one NCCL all-reduce of the flat gradient per step.

``--impl reference`` times the reference's algorithm on the host cores: the CPU oracle port
(oracle/episode.py - a torch-CPU restatement of loops/loops.py:26-61 pinned to the real reference by
tests/golden), one episode per optimizer step exactly as the reference does, all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = ("config2: Hybrid encoder + self-attention view fusion + CPL(M=5,T=9.2361,l=2.022308) + SpecAugment "
            "support/query views, FSD2018-shaped 5-way 5-shot 5-query, spectrograms [1,128,157]")
N_WAY, K_SHOT, K_QUERY, T_LEN, MELS = 5, 5, 5, 157, 128

EXPERIMENT_CONFIG = {
    "encoder_name": "Hybrid", "use_attention": True, "use_contrastive": True, "input_type": "spec",
    "train_query_augmentations": True, "project_prototypes": True, "normalize_prototypes": True,
    "loss": {"l_param": 2.022308, "cpl": {"use": True, "m_param": 5, "t_param": 9.2361},
             "angular": {"use": False, "angle": 0, "prototypes_as_anchors": True}},
    "specaug_params": {"use": True, "mask_param": 16, "W": 22, "num_mask": 1, "mask_value": 0, "p": 0.282},
    "lr": 0.0007,
}
MODEL_CONFIG = {
    "Hybrid": {"in_channels": 1, "seq_layers": 1, "seq_type": "RNN", "bidirectional": False, "hidden_channels": 64,
               "pool_dim": [3, 3], "out_dim": 64},
    "Attention": {"embed_dim": 64, "num_heads": 1, "ffn_dim": 256, "dropout": 0.1},
    "Projection": {"input_dim": 256, "hidden_dim": 512, "output_dim": 256},
}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Print the one JSON line on the real stdout (see main)."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 500 ms while the timed region runs (rank 0's GPU)."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def __enter__(self):
        if self.index is None:            # ranks other than 0: no sampler (N nvidia-smi pollers contend for the driver)
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.summary = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            self.summary = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                            "samples": len(sm)}


def build_model(device):
    from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, SelfAttention
    from afsl_b200.models.prototypical import ContrastivePrototypicalNetworks
    import contextlib
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):          # the encoder prints its parameter count like the reference
        model = ContrastivePrototypicalNetworks(EncoderModule(EXPERIMENT_CONFIG, MODEL_CONFIG),
                                                SelfAttention(MODEL_CONFIG), ProjectionHead(MODEL_CONFIG)).to(device)
    return model


def peaks_clock_mhz():
    """Maximum SM clock (MEASURED_PEAKS.json, else 1965 MHz) for the nominal fp32 FFMA peak."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh).get("sm_max_mhz", 1965.0))
    except (OSError, ValueError):
        return 1965.0


def kernel_rooflines(device, peak_gbs, episodes):
    """CUDA-event timing of each libafsl kernel alone (direct C-ABI launches on preallocated buffers larger
    than the 126 MB L2, current stream) against its algorithmic bytes (SURVEY 8d / DESIGN.md)."""
    from afsl_b200._lib import call, ptr, stream_ptr
    from afsl_b200.ops import _row_tables
    out = {}

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(reps):
            fn()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / reps * 1e-3

    traffic = ncu_traffic()

    def entry(name, bytes_per_launch, sec, units, launches=1, **extra):
        out[name] = {"bound": "hbm", "achieved": bytes_per_launch / sec / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": bytes_per_launch / sec / 1e9 / peak_gbs, "traffic": traffic.get(name), "ms": sec * 1e3,
                     "units": units, "bytes_per_launch": bytes_per_launch, "launches": launches, **extra}

    sms = torch.cuda.get_device_properties(device).multi_processor_count
    fp32_peak = sms * 128 * 2 * peaks_clock_mhz() * 1e6 / 1e12

    def entry_fp32(name, flops, sec, units):
        out[name] = {"bound": "fp32", "achieved": flops / sec / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": flops / sec / 1e12 / fp32_peak, "traffic": traffic.get(name), "ms": sec * 1e3, "units": units,
                     "flops_per_launch": flops, "launches": 1}

    st = stream_ptr()
    # ---- SpecAugment: 4 views written; algorithmic bytes 4*N*F*T*(1+V) per launch (10.05 MB / 25-sample set)
    sets, n = 256, 256 * 25
    x = torch.randn(n, 1, MELS, T_LEN, device=device)
    views = torch.empty(4, n, 1, MELS, T_LEN, device=device)
    wp = torch.randint(22, T_LEN - 22, (n,), device=device, dtype=torch.int32)
    wd = torch.randint(-22, 22, (n,), device=device, dtype=torch.int32)
    tm = torch.tensor([[[40, 12]]], device=device, dtype=torch.int32).repeat(sets, 1, 1).contiguous()
    fm = torch.tensor([[[60, 9]]], device=device, dtype=torch.int32).repeat(sets, 1, 1).contiguous()
    lo, w = _row_tables(MELS, device)
    # the warp spline evaluated on the host with the reference's op sequence, as EpisodeRunner passes it (exact path)
    from afsl_b200.utils.augmentations import warp_source_x
    sx = warp_source_x(wp.cpu().long(), wd.cpu().long(), T_LEN).to(device).contiguous()
    sec = timed(lambda: call("afsl_specaug_views_f32", ptr(x), ptr(views), ptr(wp), ptr(wd), ptr(sx), ptr(lo), ptr(w), None,
                             ptr(tm), ptr(fm), 1, 0.0, n, 25, MELS, T_LEN, 15, st))
    entry("specaug_views", 4.0 * n * MELS * T_LEN * 5 + 4.0 * n * T_LEN, sec, f"{sets} sets x 25 samples [1,128,157], host spline")
    del x, views
    # ---- fused head, D=256, 5w5s5q
    e, d, ns, nq, ways = 16384, 256, 25, 25, N_WAY
    s = torch.randn(e, ns, d, device=device)
    q = torch.randn(e, nq, d, device=device)
    sl = torch.arange(ways, device=device, dtype=torch.int32).repeat_interleave(K_SHOT).expand(e, -1).contiguous()
    ql = torch.arange(ways, device=device, dtype=torch.int32).repeat_interleave(K_QUERY).expand(e, -1).contiguous()
    protos = torch.empty(e, ways, d, device=device)
    loss = torch.empty(e, device=device)
    correct = torch.empty(e, device=device, dtype=torch.int32)
    pred = torch.empty(e * nq, device=device, dtype=torch.int32)
    post = torch.empty(e * nq, device=device)
    dl = torch.full((e,), 1.0 / e, device=device)
    ds, dq = torch.empty_like(s), torch.empty_like(q)
    fwd = lambda: call("afsl_proto_head_fwd_f32", ptr(s), ptr(sl), ptr(q), ptr(ql), None, ptr(protos), None, ptr(loss), None,
                       None, ptr(correct), e, ns, nq, ways, d, st)
    bwd = lambda: call("afsl_proto_head_bwd_f32", ptr(s), ptr(protos), ptr(sl), ptr(q), ptr(ql), None, ptr(dl), None, ptr(ds),
                       ptr(dq), e, ns, nq, ways, d, st)       # protos = the forward's output, as ops._ProtoHead passes it
    t_f, t_b = timed(fwd), timed(bwd)
    # compulsory traffic of the kernels as built: the backward takes the forward's prototypes (W rows) instead of
    # re-reading the Ns support rows, so it moves 4D(W+Nq) in and 4D(Ns+Nq) out
    b_f = (4.0 * d * (ns + nq + ways) + 4 * (ns + nq) + 8) * e            # read S,Q ; write prototypes, loss, correct
    b_b = (4.0 * d * (ways + nq) + 4.0 * d * (ns + nq) + 4 * (ns + nq) + 4) * e   # read P,Q ; write dS,dQ
    entry("proto_head_fwd", b_f, t_f, f"{e} episodes 5w5s5q D=256")
    entry("proto_head_bwd", b_b, t_b, f"{e} episodes 5w5s5q D=256")
    entry("proto_head_fwd_bwd", b_f + b_b, t_f + t_b, f"{e} episodes 5w5s5q D=256", launches=2,
          episodes_per_s=e / (t_f + t_b),
          # SURVEY 8d's unit counts a re-read of the support block in the backward: 158.9 KB per episode
          survey_unit_bytes_per_launch=(4.0 * d * (3 * (ns + nq) + ways) + 4 * (ns + nq) + 4) * e,
          survey_unit_frac=(4.0 * d * (3 * (ns + nq) + ways) + 4 * (ns + nq) + 4) * e / (t_f + t_b) / 1e9 / peak_gbs)
    # ---- evaluation head (prototypes + distances + argmax + accuracy): 4*D*(Ns+Nq) + 4*(Ns+Nq) + 8 B per task
    ev = lambda: call("afsl_proto_head_fwd_f32", ptr(s), ptr(sl), ptr(q), ptr(ql), None, None, None, None, ptr(pred), ptr(post),
                      ptr(correct), e, ns, nq, ways, d, st)
    t_e = timed(ev)
    entry("eval_head", (4.0 * d * (ns + nq) + 4 * (ns + nq) + 8) * e, t_e, f"{e} tasks 5w5s5q D=256", tasks_per_s=e / t_e)
    # ---- CPL fwd+bwd, Dp=256, Nq=25: 4*Dp*3*(Nq+W) + Nq^2/8 B per episode (92.2 KB)
    p = torch.randn(e, ways, d, device=device)
    dp = torch.empty_like(p)
    cf = lambda: call("afsl_cpl_fwd_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(loss), e, nq, ways, d, st)
    cb = lambda: call("afsl_cpl_bwd_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(dl), ptr(dp), ptr(dq), e, nq, ways, d, st)
    t_cf, t_cb = timed(cf), timed(cb)
    entry("cpl_fwd", (4.0 * d * (nq + ways) + 4 * nq + 4) * e, t_cf, f"{e} episodes Nq=25 Dp=256")
    entry("cpl_bwd", (4.0 * d * 2 * (nq + ways) + 4 * nq + 4) * e, t_cb, f"{e} episodes Nq=25 Dp=256")
    entry("cpl_fwd_bwd", (4.0 * d * 3 * (nq + ways) + nq * nq / 8) * e, t_cf + t_cb, f"{e} episodes Nq=25 Dp=256", launches=2)
    # the pair the autograd op launches: the forward also writes C [E,W,Nq] and 1/|q| [E,Nq], the backward reads them back
    # and walks the query rows once
    sim, qinv = torch.empty(e, ways, nq, device=device), torch.empty(e, nq, device=device)
    cfs = lambda: call("afsl_cpl_fwd_save_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(loss), ptr(sim), ptr(qinv), e, nq,
                       ways, d, st)
    cbs = lambda: call("afsl_cpl_bwd_saved_f32", ptr(p), ptr(q), ptr(ql), None, 9.2361, ptr(sim), ptr(qinv), ptr(dl), ptr(dp),
                       ptr(dq), e, nq, ways, d, st)
    t_cfs, t_cbs = timed(cfs), timed(cbs)
    saved_bytes = 4.0 * (ways * nq + nq)
    entry("cpl_fwd_save", (4.0 * d * (nq + ways) + 4 * nq + 4 + saved_bytes) * e, t_cfs, f"{e} episodes Nq=25 Dp=256")
    entry("cpl_bwd_saved", (4.0 * d * 2 * (nq + ways) + 4 * nq + 4 + saved_bytes) * e, t_cbs, f"{e} episodes Nq=25 Dp=256")
    entry("cpl_save_fwd_bwd", (4.0 * d * 3 * (nq + ways) + nq * nq / 8 + 2 * saved_bytes) * e, t_cfs + t_cbs,
          f"{e} episodes Nq=25 Dp=256", launches=2)
    del s, ds, dq, p, dp
    # ---- angular loss (config 3: prototypes as anchors, miner angle 0, alpha 40 deg), Dp=64, Nq=25:
    #      4*Dp*3*(Nq+W) B per episode fwd+bwd (23 KB); Gram-matrix / mining arithmetic dominates
    da = 64
    pa = torch.nn.functional.normalize(torch.randn(e, ways, da, device=device), dim=-1)
    qa = torch.nn.functional.normalize(torch.randn(e, nq, da, device=device), dim=-1)
    dpa, dqa = torch.empty_like(pa), torch.empty_like(qa)
    af = lambda: call("afsl_angular_fwd_f32", ptr(pa), ptr(qa), ptr(ql), 0.0, 40.0, 1, 0, ptr(loss), e, nq, ways, da, st)
    ab = lambda: call("afsl_angular_bwd_f32", ptr(pa), ptr(qa), ptr(ql), 0.0, 40.0, 1, 0, ptr(dl), ptr(dpa), ptr(dqa), e, nq,
                      ways, da, st)
    t_af, t_ab = timed(af, reps=5), timed(ab, reps=5)
    entry("angular_fwd", (4.0 * da * (nq + ways) + 4 * nq + 4) * e, t_af, f"{e} episodes Nq=25 Dp=64 anchors angle=0")
    entry("angular_bwd", (4.0 * da * 2 * (nq + ways) + 4 * nq + 4) * e, t_ab, f"{e} episodes Nq=25 Dp=64 anchors angle=0")
    entry("angular_fwd_bwd", 4.0 * da * 3 * (nq + ways) * e, t_af + t_ab, f"{e} episodes Nq=25 Dp=64 anchors angle=0", launches=2)
    del pa, qa, dpa, dqa
    # ---- evaluation sweep (config 5): (W, K) x D, Q = 5, tasks per launch sized to stay above the 126 MB L2
    for w_, k_, d_ in ((5, 1, 64), (5, 5, 64), (20, 5, 64), (5, 1, 256), (20, 1, 256), (20, 5, 256)):
        ns_, nq_ = w_ * k_, w_ * 5
        e_ = min(65536, max(2048, int(1.5e9 // (4 * d_ * (ns_ + nq_)))))
        s_ = torch.randn(e_, ns_, d_, device=device)
        q_ = torch.randn(e_, nq_, d_, device=device)
        sl_ = torch.arange(w_, device=device, dtype=torch.int32).repeat_interleave(k_).expand(e_, -1).contiguous()
        ql_ = torch.arange(w_, device=device, dtype=torch.int32).repeat_interleave(5).expand(e_, -1).contiguous()
        pred_ = torch.empty(e_ * nq_, device=device, dtype=torch.int32)
        post_ = torch.empty(e_ * nq_, device=device)
        corr_ = torch.empty(e_, device=device, dtype=torch.int32)
        t_ = timed(lambda: call("afsl_proto_head_fwd_f32", ptr(s_), ptr(sl_), ptr(q_), ptr(ql_), None, None, None, None,
                                ptr(pred_), ptr(post_), ptr(corr_), e_, ns_, nq_, w_, d_, st))
        entry(f"eval_head_{w_}w{k_}s_d{d_}", (4.0 * d_ * (ns_ + nq_) + 4 * (ns_ + nq_) + 8 * nq_ + 4) * e_, t_,
              f"{e_} tasks {w_}w{k_}s5q D={d_}", tasks_per_s=e_ / t_)
        del s_, q_
    del q
    # ---- grouped BN + ReLU + MaxPool, stage-1 shape [G*25, 64, 128, 157]
    g, grp, c, h, wd_ = 16, 25, 64, MELS, T_LEN
    xx = torch.randn(g * grp, c, h, wd_, device=device)
    mean, rstd, var = (torch.empty(g, c, device=device) for _ in range(3))
    gamma, beta = torch.ones(c, device=device), torch.zeros(c, device=device)
    yy = torch.empty(g * grp, c, h // 3, wd_ // 3, device=device)
    dyy = torch.randn_like(yy)
    dxx = torch.empty_like(xx)
    sums = torch.empty(g, c, 2, device=device)
    t_s = timed(lambda: call("afsl_gbn_stats_f32", ptr(xx), ptr(mean), ptr(rstd), ptr(var), g, grp, c, h, wd_, 1e-5, st), reps=5)
    t_gf = timed(lambda: call("afsl_gbn_relu_pool_fwd_f32", ptr(xx), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(yy), g, grp,
                              c, h, wd_, 1, st), reps=5)
    t_gb = timed(lambda: call("afsl_gbn_relu_pool_bwd_f32", ptr(xx), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(dyy),
                              ptr(dxx), ptr(sums), g, grp, c, h, wd_, 1, st), reps=5)
    nx, ny = xx.numel() * 4.0, yy.numel() * 4.0
    entry("gbn_stats", nx, t_s, f"{g} groups x 25 x [64,128,157]")
    entry("gbn_relu_pool_fwd", nx + ny, t_gf, f"{g} groups x 25 x [64,128,157]")
    entry("gbn_relu_pool_bwd", 2 * (nx + ny) + nx, t_gb, f"{g} groups x 25 x [64,128,157]", launches=2)
    del xx, yy, dyy, dxx
    # ---- the same three passes channels-last, stage-2 shape [G*25, 42, 52, 64] (the layout the encoder runs in)
    g, grp, c, h, wd_ = 32, 25, 64, 42, 52
    xc = torch.randn(g * grp, c, h, wd_, device=device).contiguous(memory_format=torch.channels_last)
    yc = torch.empty(g * grp, c, h // 3, wd_ // 3, device=device).contiguous(memory_format=torch.channels_last)
    dyc, dxc = torch.randn_like(yc), torch.empty_like(xc)
    mean, rstd, var = (torch.empty(g, c, device=device) for _ in range(3))
    sums = torch.empty(g, c, 2, device=device)
    from afsl_b200 import _lib
    parts = int(_lib.load().afsl_gbn_nhwc_parts(g))
    ws = torch.empty(g, parts, c, 2, device=device, dtype=torch.float64)
    t_s = timed(lambda: call("afsl_gbn_stats_nhwc_f32", ptr(xc, True), ptr(ws), parts, ptr(mean), ptr(rstd), ptr(var), g, grp, c, h,
                             wd_, 1e-5, st), reps=10)
    t_gf = timed(lambda: call("afsl_gbn_relu_pool_nhwc_fwd_f32", ptr(xc, True), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta),
                              ptr(yc, True), g, grp, c, h, wd_, 1, st), reps=10)
    t_gb = timed(lambda: call("afsl_gbn_relu_pool_nhwc_bwd_f32", ptr(xc, True), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta),
                              ptr(dyc, True), ptr(dxc, True), ptr(ws), parts, ptr(sums), g, grp, c, h, wd_, 1, st), reps=10)
    nx, ny = xc.numel() * 4.0, yc.numel() * 4.0
    entry("nhwc_gbn_stats", nx, t_s, f"{g} groups x 25 x [42,52,64] channels-last", launches=2)
    entry("nhwc_gbn_relu_pool_fwd", nx + ny, t_gf, f"{g} groups x 25 x [42,52,64] channels-last")
    entry("nhwc_gbn_relu_pool_bwd", 2 * (nx + ny) + nx, t_gb, f"{g} groups x 25 x [42,52,64] channels-last", launches=3)
    del xc, yc, dyc, dxc
    # ---- fused encoder stage 1 (conv1 + grouped BN + ReLU + MaxPool 3) at the training shape, channels-last output.
    # The forward is fp32-FMA bound (90 multiply-adds per pooled output per channel): reported against the nominal FFMA
    # peak 148 SMs x 128 lanes x 2 flop x max SM clock.  The moments pass (one read of the 1-channel input) and the
    # backward (dy + argmax codes + input, winners only recomputed) are reported against HBM.
    g, grp, h, wd_ = 64, 25, MELS, T_LEN
    ph, pw = h // 3, wd_ // 3
    n1 = g * grp
    x1 = torch.randn(n1, 1, h, wd_, device=device)
    w9 = torch.randn(64, 9, device=device) * 0.3
    a1, b1 = torch.rand(g, 64, device=device) + 0.5, torch.randn(g, 64, device=device) * 0.1
    m1, r1 = torch.zeros(g, 64, device=device), torch.ones(g, 64, device=device)
    y1 = torch.empty(n1, ph, pw, 64, device=device)
    arg1 = torch.empty(n1, ph, pw, 64, device=device, dtype=torch.uint8)
    dy1 = torch.randn_like(y1)
    parts1 = max(1, (4 * sms + g - 1) // g)
    mom = torch.empty(g, parts1, 54, device=device, dtype=torch.float64)
    parts_b = max(1, (4 * sms + g - 1) // g)
    partial = torch.empty(g, parts_b, 64, 11, device=device)
    t_m = timed(lambda: call("afsl_stage1_moments_f64", ptr(x1), ptr(mom), parts1, g, grp, h, wd_, st), reps=10)
    t_f = timed(lambda: call("afsl_stage1_fwd_f32", ptr(x1), ptr(w9), ptr(a1), ptr(b1), ptr(y1), ptr(arg1), g, grp, h, wd_, 1, 1,
                             st), reps=10)
    t_b = timed(lambda: call("afsl_stage1_bwd_f32", ptr(x1), ptr(w9), ptr(a1), ptr(b1), ptr(m1), ptr(r1), ptr(dy1), ptr(arg1),
                             ptr(partial), parts_b, g, grp, h, wd_, 1, 1, st), reps=10)
    fp32_peak = sms * 128 * 2 * peaks_clock_mhz() * 1e6 / 1e12

    # ---- view fusion (one post-norm encoder layer over V = 4 views, d = 64, FFN 256): the one head kernel that is fp32-FMA
    #      bound (397 312 flop per sample forward; backward ~2x forward + the weight gradients), at the bench step's
    #      size (N = 800 samples: 32 episodes x 25 rows) and at a size that fills the GPU (N = 409 600)
    lib = _lib.load()
    wf, pf = int(lib.afsl_view_fusion_weight_floats()), int(lib.afsl_view_fusion_param_floats())
    layer = torch.nn.TransformerEncoderLayer(d_model=64, nhead=1, dim_feedforward=256, dropout=0.0, batch_first=True).to(device)
    from afsl_b200.ops import _FUSION_PARAMS
    named = dict(layer.named_parameters())
    prm = [named[k].detach() for k in _FUSION_PARAMS]
    packed = torch.cat([t_.reshape(-1) for t_ in prm] + [prm[0].t().reshape(-1), prm[2].t().reshape(-1), prm[4].t().reshape(-1),
                                                         prm[6].t().reshape(-1)]).float().contiguous()
    assert packed.numel() == wf
    for nf in (800, 409600):
        xf = torch.randn(nf, 4, 64, device=device)
        yf, dyf, dxf = torch.empty_like(xf), torch.randn_like(xf), torch.empty_like(xf)
        gridf = int(lib.afsl_view_fusion_grid(nf, 4))
        partial_f = torch.zeros(gridf, pf, device=device)
        t_ff = timed(lambda: call("afsl_view_fusion_fwd_f32", ptr(xf), ptr(packed), ptr(yf), None, None, None, None, nf, 4, 64, 256, st))
        t_fb = timed(lambda: call("afsl_view_fusion_bwd_f32", ptr(xf), ptr(packed), ptr(dyf), None, None, None, None, ptr(dxf),
                                  ptr(partial_f), nf, 4, 64, 256, st))
        entry_fp32(f"fusion_fwd_n{nf}", 397312.0 * nf, t_ff, f"{nf} samples x 4 views x 64")
        entry_fp32(f"fusion_bwd_n{nf}", 3.0 * 397312.0 * nf, t_fb, f"{nf} samples x 4 views x 64 (dx + weight gradients: ~3x the forward flops)")
        del xf, yf, dyf, dxf
    # ---- majority vote (E2): 16 bytes per query segment (pred, clip id, label, posterior) + 8 per task out
    tasks_v, clips_v = 65536, 25
    seg = torch.randint(1, 9, (tasks_v, clips_v), device=device)
    rows_v = int(seg.sum())
    offs_v = torch.cat([torch.zeros(1, dtype=torch.int64, device=device), seg.sum(1).cumsum(0)]).to(torch.int32)
    cid_v = torch.repeat_interleave(torch.arange(clips_v, device=device).repeat(tasks_v), seg.reshape(-1)).to(torch.int32)
    lab_v = (cid_v // 5).to(torch.int32)
    pred_v = torch.randint(0, 5, (rows_v,), device=device, dtype=torch.int32)
    post_v = -torch.rand(rows_v, device=device)
    corr_v, ncl_v = torch.empty(tasks_v, device=device, dtype=torch.int32), torch.empty(tasks_v, device=device, dtype=torch.int32)
    t_v = timed(lambda: call("afsl_eval_vote_i32", ptr(pred_v), ptr(cid_v), ptr(lab_v), ptr(post_v), ptr(offs_v), 1, ptr(corr_v),
                             ptr(ncl_v), tasks_v, st))
    entry("eval_vote", 16.0 * rows_v + 12.0 * tasks_v, t_v, f"{tasks_v} tasks x 25 clips x U{{1..8}} segments ({rows_v} rows), min_label",
          tasks_per_s=tasks_v / t_v)
    shape1 = f"{g} groups x 25 x [1,128,157] -> [42,52,64] channels-last"
    entry_fp32("stage1_fwd", 2.0 * 90 * n1 * ph * pw * 64, t_f, shape1)
    entry("stage1_moments", 4.0 * x1.numel(), t_m, shape1)       # one pass over the 1-channel input, 14 FMA per pixel
    entry("stage1_bwd", 4.0 * dy1.numel() + arg1.numel() + 4.0 * x1.numel(), t_b, shape1)
    return out


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels above, from the committed
    `ncu --set full` captures of tools/kernels_bench.py (profiles/ncu_traffic.json; same shapes as timed here)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as fh:
        return {k: v for k, v in json.load(fh).items() if not k.startswith("_")}


def cpu_episode_runner(threads):
    """The CPU oracle port of one reference training episode (config 2)."""
    from oracle import episode as oep
    from oracle import modules as om
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    net = om.FusedViewsNet(om.ViewEncoder(om.build_encoder("Hybrid", T_LEN)), om.ViewFusion(64, 1, 256, 0.1),
                           om.Projection(256, 512, 256))
    opt = torch.optim.Adam(net.parameters(), lr=EXPERIMENT_CONFIG["lr"])
    gen = torch.Generator().manual_seed(99)
    sl = torch.arange(N_WAY).repeat_interleave(K_SHOT)
    ql = torch.arange(N_WAY).repeat_interleave(K_QUERY)

    def one_episode():
        s = torch.randn(N_WAY * K_SHOT, 1, MELS, T_LEN, generator=gen)
        q = torch.randn(N_WAY * K_QUERY, 1, MELS, T_LEN, generator=gen)
        return oep.train_step(net, opt, s, sl, q, ql, EXPERIMENT_CONFIG)
    return one_episode


def head_baselines(device, threads):
    """SURVEY 8d's other baselines for the head alone, given embeddings (5-way 5-shot 5-query, D = 256):
    (i) the CPU oracle port, episode by episode like the reference (prototypes, -cdist, log-softmax/NLL, backward);
    (ii) the same op chain in eager PyTorch ON THE B200, episode by episode (what the reference itself would launch);
    (iii) libafsl's fused head on the same episodes in one launch pair."""
    import afsl_b200.ops as ops
    from oracle import head as ohead
    torch.set_num_threads(threads)
    e, d, ways = 256, 256, N_WAY
    gen = torch.Generator().manual_seed(5)
    s = torch.randn(e, ways * K_SHOT, d, generator=gen)
    q = torch.randn(e, ways * K_QUERY, d, generator=gen)
    sl = torch.arange(ways).repeat_interleave(K_SHOT)
    ql = torch.arange(ways).repeat_interleave(K_QUERY)

    def chain(si, qi, sli, qli):
        si, qi = si.clone().requires_grad_(True), qi.clone().requires_grad_(True)
        loss = ohead.fsl_loss(ohead.prototypes(si, sli), qi, qli)
        loss.backward()
        return loss
    n_cpu = 64
    chain(s[0], q[0], sl, ql)
    t0 = time.perf_counter()
    for i in range(n_cpu):
        chain(s[i], q[i], sl, ql)
    cpu_eps = n_cpu / (time.perf_counter() - t0)
    sg, qg, slg, qlg = s.to(device), q.to(device), sl.to(device), ql.to(device)
    for i in range(8):
        chain(sg[i], qg[i], slg, qlg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e):
        chain(sg[i], qg[i], slg, qlg)
    torch.cuda.synchronize()
    eager_eps = e / (time.perf_counter() - t0)
    big = 16384
    sb = torch.randn(big, ways * K_SHOT, d, device=device).requires_grad_(True)
    qb = torch.randn(big, ways * K_QUERY, d, device=device).requires_grad_(True)
    slb, qlb = slg.expand(big, -1).contiguous(), qlg.expand(big, -1).contiguous()
    for _ in range(3):
        ops.proto_head(sb, slb, qb, qlb, n_way=ways)[0].sum().backward()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        sb.grad = qb.grad = None
        ops.proto_head(sb, slb, qb, qlb, n_way=ways)[0].sum().backward()
    b.record()
    torch.cuda.synchronize()
    fused_eps = 5 * big / (a.elapsed_time(b) * 1e-3)
    return {"workload": "head only on given embeddings: prototypes + -cdist + log-softmax/NLL, forward + backward, 5w5s5q D=256",
            "unit": "episodes/s",
            "cpu_oracle_port": {"value": cpu_eps, "cores": threads, "sample": f"{n_cpu} episodes, one at a time"},
            "eager_pytorch_on_b200": {"value": eager_eps, "sample": f"{e} episodes, one at a time (the reference's launch pattern)"},
            "libafsl_autograd": {"value": fused_eps, "sample": f"{big} episodes per launch pair through ops.proto_head (autograd wrapper included)"}}


def cpu_eval_baselines(threads):
    """CPU oracle port of evaluate_single_segment / evaluate_multisegment_loop per task (loops/loops.py:66-121,250-283) on
    the bench model (Hybrid + SpecAugment support/query views + view fusion, eval mode): tasks/s on a bounded sample."""
    import numpy as np
    from oracle import episode as oep
    from oracle import modules as om
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    net = om.FusedViewsNet(om.ViewEncoder(om.build_encoder("Hybrid", T_LEN)), om.ViewFusion(64, 1, 256, 0.1),
                           om.Projection(256, 512, 256)).eval()
    gen = torch.Generator().manual_seed(17)
    sl = torch.arange(N_WAY).repeat_interleave(K_SHOT)
    ql = torch.arange(N_WAY).repeat_interleave(K_QUERY)
    rng = np.random.RandomState(3)

    def single():
        s = torch.randn(N_WAY * K_SHOT, 1, MELS, T_LEN, generator=gen)
        q = torch.randn(N_WAY * K_QUERY, 1, MELS, T_LEN, generator=gen)
        return oep.eval_task(net, oep.make_views(s, EXPERIMENT_CONFIG, True), sl, oep.make_views(q, EXPERIMENT_CONFIG, True), ql)

    def multi():
        seg = rng.randint(1, 9, size=N_WAY * K_QUERY)
        rows = int(seg.sum())
        s = torch.randn(N_WAY * K_SHOT, 1, MELS, T_LEN, generator=gen)
        q = torch.randn(rows, 1, MELS, T_LEN, generator=gen)
        ids = torch.from_numpy(np.repeat(np.arange(N_WAY * K_QUERY), seg))
        labels = torch.from_numpy(np.repeat(np.arange(N_WAY).repeat(K_QUERY), seg))
        return oep.eval_task(net, oep.make_views(s, EXPERIMENT_CONFIG, True), sl, oep.make_views(q, EXPERIMENT_CONFIG, True), labels,
                             clip_ids=ids, tie_strategy="min_label")
    out = {}
    for name, fn, n in (("single_segment", single, 4), ("multi_segment", multi, 2)):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        out[name] = {"value": n / (time.perf_counter() - t0), "unit": "tasks/s", "cores": threads, "kind": "port",
                     "sample": f"{n} tasks after 1 warm-up, oracle/episode.py::eval_task"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step = cpu_episode_runner(threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    line = {"impl": "reference", "metric": "episodes/sec (train fwd+bwd)", "value": value, "unit": "episodes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "episodes_per_step": 1, "device": "cpu"},
            "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} episodes, one optimizer step each (reference schedule), "
                                       "oracle/episode.py torch-CPU port of loops/loops.py:26-61"},
            "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def eval_throughput(runner, device, world, rank, tasks_per_step, steps):
    """Test tasks / s through EpisodeRunner.eval_step (BASELINE metric's second half), device-resident inputs:
    (a) single-segment 5-way 5-shot 5-query tasks (loops/loops.py:84-121), (b) multi-segment tasks with
    S ~ U{1..8} segments per query clip and the majority vote (loops/loops.py:250-279).  Ranks own contiguous task
    blocks; the per-task accuracies are all-gathered at the end (inside the timed region)."""
    import numpy as np
    import torch.distributed as dist
    from afsl_b200 import parallel
    from afsl_b200.episodes import EpisodeBatch, synthetic_batch
    out = {}
    e = tasks_per_step
    single = [synthetic_batch(e, N_WAY, K_SHOT, K_QUERY, T_LEN, seed=7000 + 10 * rank + i, device=str(device)) for i in range(2)]
    # multi-segment: 25 query clips per task, 1..8 segments each, packed rows + CSR offsets
    rng = np.random.RandomState(4321 + rank)
    seg = rng.randint(1, 9, size=(e, N_WAY * K_QUERY))
    rows_per_task = seg.sum(1)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(rows_per_task)]).astype(np.int64))
    rows = int(offsets[-1])
    clip_ids = torch.from_numpy(np.concatenate([np.repeat(np.arange(N_WAY * K_QUERY), s) for s in seg]).astype(np.int64))
    clip_label = np.arange(N_WAY).repeat(K_QUERY)
    row_labels = torch.from_numpy(np.concatenate([np.repeat(clip_label, s) for s in seg]).astype(np.int64))
    gen = torch.Generator(device=device).manual_seed(99 + rank)
    multi = EpisodeBatch(single[0].support, single[0].support_labels,
                         torch.randn(1, rows, 1, MELS, T_LEN, generator=gen, device=device), row_labels.view(1, -1).to(device), N_WAY)
    clip_ids, offsets_dev = clip_ids.to(device), offsets.to(device)

    def timed(fn, units):
        for _ in range(2):
            fn(0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        accs = [fn(i) for i in range(steps)]                                  # device tensors: no host sync per step
        acc = parallel.gather_accuracies(torch.cat(accs).cpu().numpy(), world * steps * e)
        b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        return {"value": world * steps * units / (ms * 1e-3), "ms_per_step": ms / steps, "tasks_per_step_per_gpu": e,
                "mean_accuracy": float(acc.mean())}

    # test_query_augmentations = true (README.md:92): the view-fusion model needs the four query views at test time too
    out["test_tasks_per_s"] = dict(timed(lambda i: runner.eval_step(single[i % 2], augment_query=True, as_tensor=True), e), unit="tasks/s",
                                   workload="5-way 5-shot 5-query single-segment tasks, SpecAugment support + query views, "
                                            "eval-mode encoder")
    out["multiseg_tasks_per_s"] = dict(timed(lambda i: runner.eval_step(multi, augment_query=True, clip_ids=clip_ids,
                                                                        seg_offsets=offsets_dev, tie_strategy="min_label", as_tensor=True), e),
                                       unit="tasks/s",
                                       rows_per_step_per_gpu=rows,
                                       workload="5-way 5-shot, 25 query clips x U{1..8} segments per task, majority vote (min_label)")
    return out


def conv_variant(args, device, world, rank, episodes, host, barrier, kind):
    """A second, LABELLED line: the same step with another convolution arithmetic than the fp32 channels-last headline.
    ``kind`` "tf32": cuDNN TF32 convolutions allowed (PyTorch's default, i.e. what the reference itself would run on this
    GPU).  ``kind`` "nchw": still fp32, but blocks 2-4 on NCHW tensors, where cuDNN picks faster, less exact kernels
    (embeddings 1e-5 from the channels-last ones, projection-head gradients 1e-2 from the fp32 oracle: outside the parity
    suite's bounds, hence not the headline).  Fresh model with the headline's initial weights; reports throughput with
    device-resident inputs and the first step's loss next to the headline arithmetic's for the same batch and seeds."""
    import random
    import numpy as np
    import torch.distributed as dist
    import afsl_b200.models.main_modules as mm
    from afsl_b200 import parallel
    from afsl_b200.episodes import EpisodeRunner

    def arithmetic(on):
        torch.backends.cudnn.allow_tf32 = on and kind == "tf32"
        mm.NCHW_FP32_CONVS = on and kind == "nchw"

    first = {}
    for on in (False, True):
        arithmetic(on)
        model = build_model(device)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        torch.manual_seed(4321); np.random.seed(4321); random.seed(4321)
        runner = EpisodeRunner(model, EXPERIMENT_CONFIG, None, replay_reference_rng=False, use_cuda_graph=False)
        first[on] = runner.train_step(host[0])["loss"].double().cpu()
    dev_loss = float(((first[True] - first[False]).abs() / first[False].abs()).max())
    arithmetic(True)
    model = build_model(device)
    opt = torch.optim.Adam(model.parameters(), lr=EXPERIMENT_CONFIG["lr"])
    runner = EpisodeRunner(model, EXPERIMENT_CONFIG, opt, replay_reference_rng=False, use_cuda_graph=not args.no_graph)
    dp = parallel.EpisodeDataParallel(model)
    runner.grad_sync = dp.sync_gradients if world > 1 else None
    resident = [b.to(device) for b in host]
    steps = min(args.steps, 30)
    for i in range(args.warmup):
        runner.train_step(resident[i % len(resident)])
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        runner.train_step(resident[i % len(resident)])
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    arithmetic(False)
    label = {"tf32": "cuDNN TF32 convolutions allowed (NOT the headline; parity suite validates fp32 only)",
             "nchw": "fp32, convolution blocks 2-4 on NCHW tensors: faster, less exact cuDNN kernels (NOT the headline: "
                     "gradients leave the parity suite's bounds)"}[kind]
    return {"label": label, "dtype": "tf32-conv" if kind == "tf32" else "f32-nchw-conv",
            "value": world * episodes * steps / (ms * 1e-3), "unit": "episodes/s", "ms_per_step": ms / steps, "steps": steps,
            "max_rel_loss_deviation_vs_fp32_first_step": dev_loss}


def run_b200(args):
    import torch.distributed as dist
    import afsl_b200.ops as ops
    from afsl_b200 import parallel
    from afsl_b200.episodes import EpisodeRunner, synthetic_batch

    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    # torchrun pins OMP_NUM_THREADS=1; the host side of a step (randomness draws, the reference-exact warp spline) is
    # vectorised torch-CPU work: give every rank its share of the host cores
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, world)))
    device = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True
    # the headline is fp32 end to end: cuDNN's TF32 convolution / RNN kernels are switched off (PyTorch's default allows
    # them), matching the fp32 arithmetic every parity test checks; matmuls are fp32 by PyTorch's own default.
    # `--tf32` measures the TF32-convolution variant as a second, labelled line under extra_metrics.
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    pk, pk_src = peaks()
    E = args.episodes

    model = build_model(device)
    opt = torch.optim.Adam(model.parameters(), lr=EXPERIMENT_CONFIG["lr"])
    runner = EpisodeRunner(model, EXPERIMENT_CONFIG, opt, replay_reference_rng=False, use_cuda_graph=not args.no_graph)
    dp = parallel.EpisodeDataParallel(model)
    runner.grad_sync = dp.sync_gradients if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # rotate over several distinct input batches; per-step activations (GBs) exceed the 126 MB L2 anyway
    n_rot = 3
    host = [synthetic_batch(E, N_WAY, K_SHOT, K_QUERY, T_LEN, seed=1234 + rank * 100 + i).pin() for i in range(n_rot)]
    resident = [b.to(device) for b in host]
    torch.cuda.synchronize()

    # ---- device-resident timing
    for i in range(args.warmup):
        runner.train_step(resident[i % n_rot])
    barrier()
    launches0 = ops.launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local if rank == 0 else None) as clocks:
        start.record()
        for i in range(args.steps):
            runner.train_step(resident[i % n_rot])
        end.record()
        barrier()
    ms = start.elapsed_time(end)
    # kernels launched eagerly + those inside the replayed CUDA graph (counted once, at capture)
    launches = ops.launch_count() - launches0 + (runner.launches_per_replay * args.steps if runner.use_cuda_graph else 0)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * E * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the loss
    del resident
    for i in range(max(1, args.warmup // 2)):
        float(runner.train_step(host[i % n_rot])["loss"].mean().item())
    barrier()
    t0 = time.perf_counter()
    start.record()
    for i in range(args.steps):
        # the following batch is handed over too: its H2D copy overlaps this step's compute (inside the timed region)
        out = runner.train_step(host[i % n_rot], next_batch=host[(i + 1) % n_rot] if i + 1 < args.steps else None)
        loss_host = out["loss"].cpu()              # D2H of the per-episode losses
    end.record()
    barrier()
    e2e_ms = start.elapsed_time(end)
    t = torch.tensor([e2e_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * E * args.steps / (e2e_ms * 1e-3)
    h2d_bytes = host[0].nbytes() * world            # whole job, like `value`

    tf32_line = nchw_line = None
    if not args.skip_tf32:
        tf32_line = conv_variant(args, device, world, rank, E, host, barrier, "tf32")
        nchw_line = conv_variant(args, device, world, rank, E, host, barrier, "nchw")
    del host
    evals = eval_throughput(runner, device, world, rank, args.eval_tasks, max(2, args.steps // 2)) if not args.skip_eval else {}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    roofs = kernel_rooflines(device, pk["hbm_gbs"], E) if not args.skip_kernels else {}
    cpu = None
    if world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        step = cpu_episode_runner(threads)
        step()
        t0 = time.perf_counter()
        n_cpu = 4
        for _ in range(n_cpu):
            step()
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": "episodes/s", "cores": threads, "kind": "port",
               "sample": f"{n_cpu} episodes after 1 warm-up, one optimizer step each; oracle/episode.py (torch-CPU port "
                         "of loops/loops.py:26-61, pinned to the reference by tests/golden)"}
    baselines = None
    if world == 1 and not args.skip_cpu:
        baselines = {"head_only": head_baselines(device, os.cpu_count() or 1), "cpu_eval_per_task": cpu_eval_baselines(os.cpu_count() or 1)}
    line = {
        "metric": "episodes/sec (train fwd+bwd)", "value": value, "unit": "episodes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "episodes_per_step_per_gpu": E, "global_episodes_per_step": world * E,
                   "parallelism": f"episode-sharded dp{world}", "cudnn_conv_tf32": torch.backends.cudnn.allow_tf32,
                   "cuda_graph": runner.use_cuda_graph,
                   "l2": f"inputs rotate over {n_rot} batches; per-step activations exceed the 126 MB L2",
                   "peaks": pk_src},
        "e2e": {"value": e2e_value, "unit": "episodes/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(loss_host.numel() * 4) * world, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks.summary,
        # the fused loss-head pair (prototypes + distances + log-softmax/NLL forward, and its backward): the
        # kernels BASELINE.json's "head HBM GB/s" names; every other libafsl kernel is listed under "kernels"
        "roofline": roofs.get("proto_head_fwd_bwd"),
        "kernels": roofs,
        "cpu_baseline": cpu,
        "baselines": baselines,
        "extra_metrics": dict(evals, **({"tf32_conv_variant": tf32_line, "nchw_conv_variant": nchw_line} if tf32_line else {})),
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # stdout carries exactly one JSON line: everything else - Python prints and C-level writes such as NCCL's version banner,
    # which ignores NCCL_DEBUG_FILE - is sent to stderr by pointing file descriptor 1 at it until the line is printed
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--episodes", type=int, default=32, help="episodes per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--eval-tasks", type=int, default=64, help="test tasks per eval step per GPU")
    ap.add_argument("--skip-eval", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-kernels", action="store_true")
    ap.add_argument("--skip-tf32", action="store_true", help="skip the labelled TF32 / NCHW convolution variants")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback "
                             "(use --impl reference for the CPU oracle port)")
        run_b200(args)


if __name__ == "__main__":
    main()
