"""torch.autograd bindings of the libafsl kernels.

Every function takes CUDA fp32 tensors, optionally with a leading episode
dimension E (the reference always has E = 1 and no such dimension), converts
labels to int32 and calls the C ABI on the current stream.  There is no CPU
implementation here on purpose (see _lib.py).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

TIE_STRATEGIES = {"min_label": 1, "max_posterior": 2}   # anything else -> first seen (loops/loops.py:233-234)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32(t: torch.Tensor) -> torch.Tensor:
    return t.to(dtype=torch.int32).contiguous()


def _same_dim(what: str, a: torch.Tensor, b: torch.Tensor) -> None:
    """The kernels take ONE embedding dimension D for both operands; a mismatch (e.g. fused 4-view support
    features against single-view query features) would read out of bounds, so it is refused here."""
    if a.shape[-1] != b.shape[-1]:
        raise ValueError(f"{what}: embedding dimensions differ ({tuple(a.shape)} vs {tuple(b.shape)})")


def _batched(feats: torch.Tensor, labels: Optional[torch.Tensor]):
    """-> (feats [E,N,D], labels [E,N] or None, had_batch_dim)."""
    if feats.dim() == 2:
        return feats.unsqueeze(0), (labels.unsqueeze(0) if labels is not None else None), False
    if feats.dim() == 3:
        if labels is not None and labels.dim() == 1:
            labels = labels.unsqueeze(0).expand(feats.shape[0], -1)
        return feats, labels, True
    raise ValueError("Illegal backbone or feature shape. Expected output for an image is a 1-dim tensor.")


def infer_n_way(labels: torch.Tensor) -> int:
    """len(torch.unique(labels)) as in models/util_functions.py:17 (synchronises when labels are on the GPU)."""
    return int(torch.unique(labels).numel())


# ---------------------------------------------------------------------------------- prototypes
class _Prototypes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support, labels, n_way):
        e, ns, d = support.shape
        protos = torch.empty(e, n_way, d, device=support.device, dtype=torch.float32)
        call("afsl_prototypes_fwd_f32", ptr(support), ptr(labels), ptr(protos), e, ns, n_way, d, stream_ptr())
        ctx.save_for_backward(labels)
        ctx.shape = (e, ns, n_way, d)
        return protos

    @staticmethod
    def backward(ctx, d_protos):
        (labels,) = ctx.saved_tensors
        e, ns, w, d = ctx.shape
        d_protos = _f32(d_protos)
        d_support = torch.empty(e, ns, d, device=d_protos.device, dtype=torch.float32)
        call("afsl_prototypes_bwd_f32", ptr(d_protos), ptr(labels), ptr(d_support), e, ns, w, d, stream_ptr())
        return d_support, None, None


def prototypes(support: torch.Tensor, labels: torch.Tensor, n_way: Optional[int] = None) -> torch.Tensor:
    """Per-label mean of support rows: [Ns,D],[Ns] -> [W,D] or [E,Ns,D],[E,Ns] -> [E,W,D]."""
    s, l, had = _batched(support, labels)
    if n_way is None:
        n_way = infer_n_way(l[0])
    out = _Prototypes.apply(_f32(s), _i32(l), int(n_way))
    return out if had else out[0]


# ---------------------------------------------------------------------------------- scores / proto loss
class _ProtoScores(torch.autograd.Function):
    """(-dist, loss) given prototypes.  loss is NaN-free only when labels are given."""

    @staticmethod
    def forward(ctx, protos, queries, labels, offsets, want_scores, want_loss):
        e, w, d = protos.shape
        rows = queries.shape[0] * queries.shape[1] if queries.dim() == 3 else queries.shape[0]
        nq = queries.shape[1] if queries.dim() == 3 else ctx_max_rows(offsets)
        dev = protos.device
        scores = torch.empty(rows, w, device=dev, dtype=torch.float32) if want_scores else None
        loss = torch.empty(e, device=dev, dtype=torch.float32) if want_loss else None
        call("afsl_proto_scores_fwd_f32", ptr(protos), ptr(queries), ptr(labels), ptr(offsets), ptr(scores), ptr(loss),
             None, None, None, e, nq, w, d, stream_ptr())
        ctx.save_for_backward(protos, queries, labels, offsets)
        ctx.dims = (e, nq, w, d)
        ctx.flags = (want_scores, want_loss)
        if want_scores and queries.dim() == 3:
            scores = scores.view(e, nq, w)
        empty = torch.empty(0, device=dev)
        return (scores if want_scores else empty), (loss if want_loss else empty)

    @staticmethod
    def backward(ctx, d_scores, d_loss):
        protos, queries, labels, offsets = ctx.saved_tensors
        e, nq, w, d = ctx.dims
        want_scores, want_loss = ctx.flags
        d_scores = _f32(d_scores).reshape(-1, w) if want_scores else None
        d_loss = _f32(d_loss) if want_loss else None
        d_protos = torch.empty_like(protos)
        d_queries = torch.empty_like(queries)
        call("afsl_proto_scores_bwd_f32", ptr(protos), ptr(queries), ptr(labels), ptr(offsets), ptr(d_loss),
             ptr(d_scores), ptr(d_protos), ptr(d_queries), e, nq, w, d, stream_ptr())
        return d_protos, d_queries, None, None, None, None


def ctx_max_rows(offsets: torch.Tensor) -> int:
    """Largest per-task row count of a CSR offsets tensor (host value; offsets are built on the host)."""
    return int((offsets[1:] - offsets[:-1]).max().item())


def l2_scores(queries: torch.Tensor, protos: torch.Tensor) -> torch.Tensor:
    """-cdist(queries, protos): [Nq,D],[W,D] -> [Nq,W] (or with a leading E)."""
    q, _, had = _batched(queries, None)
    p = protos if protos.dim() == 3 else protos.unsqueeze(0)
    _same_dim("l2_scores", q, p)
    scores, _ = _ProtoScores.apply(_f32(p), _f32(q), None, None, True, False)
    return scores if had else scores[0]


def proto_loss(protos: torch.Tensor, queries: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """FSL loss per episode: mean NLL of log_softmax(-cdist).  Returns a 0-dim tensor without E."""
    q, l, had = _batched(queries, labels)
    p = protos if protos.dim() == 3 else protos.unsqueeze(0)
    _same_dim("proto_loss", q, p)
    _, loss = _ProtoScores.apply(_f32(p), _f32(q), _i32(l), None, False, True)
    return loss if had else loss[0]


# ---------------------------------------------------------------------------------- fused head
class _ProtoHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support, s_labels, queries, q_labels, n_way):
        e, ns, d = support.shape
        nq = queries.shape[1]
        dev = support.device
        protos = torch.empty(e, n_way, d, device=dev, dtype=torch.float32)
        loss = torch.empty(e, device=dev, dtype=torch.float32)
        correct = torch.empty(e, device=dev, dtype=torch.int32)
        call("afsl_proto_head_fwd_f32", ptr(support), ptr(s_labels), ptr(queries), ptr(q_labels), None, ptr(protos), None,
             ptr(loss), None, None, ptr(correct), e, ns, nq, n_way, d, stream_ptr())
        ctx.save_for_backward(support, s_labels, queries, q_labels, protos)
        ctx.dims = (e, ns, nq, n_way, d)
        ctx.mark_non_differentiable(correct)
        return loss, protos, correct

    @staticmethod
    def backward(ctx, d_loss, d_protos, _d_correct):
        support, s_labels, queries, q_labels, protos = ctx.saved_tensors
        e, ns, nq, w, d = ctx.dims
        d_loss = _f32(d_loss) if d_loss is not None else torch.zeros(e, device=support.device)
        d_protos = _f32(d_protos) if d_protos is not None else None
        d_support = torch.empty_like(support)
        d_queries = torch.empty_like(queries)
        # the saved prototypes stand in for the support block: the backward reads P and Q only
        call("afsl_proto_head_bwd_f32", ptr(support), ptr(protos), ptr(s_labels), ptr(queries), ptr(q_labels), None,
             ptr(d_loss), ptr(d_protos), ptr(d_support), ptr(d_queries), e, ns, nq, w, d, stream_ptr())
        return d_support, None, d_queries, None, None


def proto_head(support, s_labels, queries, q_labels, n_way: Optional[int] = None):
    """Fused prototypes + FSL loss (+ #correct): returns (loss [E], protos [E,W,D], correct [E])."""
    s, sl, had = _batched(support, s_labels)
    q, ql, _ = _batched(queries, q_labels)
    _same_dim("proto_head", s, q)
    if q.shape[0] != s.shape[0]:
        raise ValueError(f"proto_head: {s.shape[0]} support episodes vs {q.shape[0]} query episodes")
    if n_way is None:
        n_way = infer_n_way(sl[0])
    loss, protos, correct = _ProtoHead.apply(_f32(s), _i32(sl), _f32(q), _i32(ql), int(n_way))
    return (loss, protos, correct) if had else (loss[0], protos[0], correct[0])


@torch.no_grad()
def proto_eval(support, s_labels, queries, q_labels=None, n_way: Optional[int] = None,
               q_offsets: Optional[torch.Tensor] = None, max_rows: Optional[int] = None, want_scores: bool = False):
    """Evaluation head: prototypes + scores -> (pred, posterior, correct, scores?).

    ``queries`` is [E,Nq,D], or packed [rows,D] with CSR ``q_offsets`` [E+1] for ragged
    multi-segment tasks (``max_rows`` = largest task, computed on the host when omitted).
    """
    s, sl, _ = _batched(support, s_labels)
    e, ns, d = s.shape
    _same_dim("proto_eval", s, queries)
    if n_way is None:
        n_way = infer_n_way(sl[0])
    dev = s.device
    if q_offsets is None:
        q = queries if queries.dim() == 3 else queries.unsqueeze(0)
        nq, rows = q.shape[1], q.shape[0] * q.shape[1]
        off = None
    else:
        q = queries
        rows = q.shape[0]
        nq = int(max_rows) if max_rows is not None else ctx_max_rows(q_offsets)
        off = _i32(q_offsets)
    ql = _i32(q_labels.reshape(-1)) if q_labels is not None else None
    pred = torch.empty(rows, device=dev, dtype=torch.int32)
    post = torch.empty(rows, device=dev, dtype=torch.float32)
    correct = torch.empty(e, device=dev, dtype=torch.int32) if ql is not None else None
    scores = torch.empty(rows, n_way, device=dev, dtype=torch.float32) if want_scores else None
    s32, sl32, q32 = _f32(s), _i32(sl), _f32(q)          # keep converted temporaries alive across the launch
    call("afsl_proto_head_fwd_f32", ptr(s32), ptr(sl32), ptr(q32), ptr(ql), ptr(off), None, ptr(scores), None,
         ptr(pred), ptr(post), ptr(correct), e, ns, nq, int(n_way), d, stream_ptr())
    return pred, post, correct, scores


# ---------------------------------------------------------------------------------- normalise
class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        y = torch.empty_like(x)
        rows, d = x.numel() // x.shape[-1], x.shape[-1]
        call("afsl_l2_normalize_fwd_f32", ptr(x), ptr(y), rows, d, float(eps), stream_ptr())
        ctx.save_for_backward(x)
        ctx.eps = float(eps)
        return y

    @staticmethod
    def backward(ctx, d_y):
        (x,) = ctx.saved_tensors
        d_y = _f32(d_y)
        d_x = torch.empty_like(x)
        rows, d = x.numel() // x.shape[-1], x.shape[-1]
        call("afsl_l2_normalize_bwd_f32", ptr(x), ptr(d_y), ptr(d_x), rows, d, ctx.eps, stream_ptr())
        return d_x, None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, p=2, dim=-1, eps) on the last dimension."""
    return _L2Normalize.apply(_f32(x), eps)


# ---------------------------------------------------------------------------------- Linear (projection head)
class _Linear(torch.autograd.Function):
    """y = x w^T + b (optionally ReLU) on the libafsl SGEMM; x [M,K], w [N,K], b [N]."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        m, k = x.shape
        n = weight.shape[0]
        y = torch.empty(m, n, device=x.device, dtype=torch.float32)
        call("afsl_linear_fwd_f32", ptr(x), ptr(weight), ptr(bias), ptr(y), m, n, k, int(relu), stream_ptr())
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, d_y):
        x, weight, y_relu = ctx.saved_tensors
        m, k = x.shape
        n = weight.shape[0]
        d_y = _f32(d_y)
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        d_x = torch.empty_like(x) if need_x else None
        d_w = torch.empty_like(weight) if need_w else None
        d_b = torch.empty(n, device=x.device, dtype=torch.float32) if need_b else None
        if m == 0:
            return (torch.zeros_like(x) if need_x else None, torch.zeros_like(weight) if need_w else None,
                    torch.zeros(n, device=x.device) if need_b else None, None)
        ws = torch.empty(int(_lib.load().afsl_linear_bwd_workspace_floats(m, n, k)), device=x.device, dtype=torch.float32)
        call("afsl_linear_bwd_f32", ptr(x), ptr(weight), ptr(y_relu), ptr(d_y), ptr(d_x), ptr(d_w), ptr(d_b), ptr(ws), m, n, k,
             stream_ptr())
        return d_x, d_w, d_b, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """F.linear(x, weight, bias) (then ReLU when ``relu``) over the last dimension, any leading dimensions."""
    lead = x.shape[:-1]
    if x.shape[-1] != weight.shape[1]:
        raise ValueError(f"linear: input features {x.shape[-1]} vs weight {tuple(weight.shape)}")
    y = _Linear.apply(_f32(x.reshape(-1, x.shape[-1])), _f32(weight), _f32(bias) if bias is not None else None, relu)
    return y.view(*lead, weight.shape[0])


# ---------------------------------------------------------------------------------- CPL
class _Cpl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, protos, queries, labels, keep, temperature):
        e, w, d = protos.shape
        nq = queries.shape[1]
        loss = torch.empty(e, device=protos.device, dtype=torch.float32)
        # when a backward will follow and the warp kernels take the shape, the forward hands its similarity matrix and
        # reciprocal norms to the backward (0.6 KB per episode), which then walks the query rows once instead of twice
        needs_grad = protos.requires_grad or queries.requires_grad
        saved = (needs_grad and os.environ.get("AFSL_CPL_SAVE", "1") != "0"       # the parity tests run both backwards
                 and bool(_lib.load().afsl_cpl_saved_supported(nq, w, d)))
        sim = qinv = None
        if saved:
            sim = torch.empty(e, w, nq, device=protos.device, dtype=torch.float32)
            qinv = torch.empty(e, nq, device=protos.device, dtype=torch.float32)
            call("afsl_cpl_fwd_save_f32", ptr(protos), ptr(queries), ptr(labels), ptr(keep), float(temperature), ptr(loss),
                 ptr(sim), ptr(qinv), e, nq, w, d, stream_ptr())
        else:
            call("afsl_cpl_fwd_f32", ptr(protos), ptr(queries), ptr(labels), ptr(keep), float(temperature), ptr(loss), e, nq,
                 w, d, stream_ptr())
        ctx.save_for_backward(protos, queries, labels, keep, sim, qinv)
        ctx.temperature = float(temperature)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        protos, queries, labels, keep, sim, qinv = ctx.saved_tensors
        e, w, d = protos.shape
        nq = queries.shape[1]
        d_protos, d_queries = torch.empty_like(protos), torch.empty_like(queries)
        d_loss = _f32(d_loss)
        if sim is not None:
            call("afsl_cpl_bwd_saved_f32", ptr(protos), ptr(queries), ptr(labels), ptr(keep), ctx.temperature, ptr(sim),
                 ptr(qinv), ptr(d_loss), ptr(d_protos), ptr(d_queries), e, nq, w, d, stream_ptr())
        else:
            call("afsl_cpl_bwd_f32", ptr(protos), ptr(queries), ptr(labels), ptr(keep), ctx.temperature, ptr(d_loss),
                 ptr(d_protos), ptr(d_queries), e, nq, w, d, stream_ptr())
        return d_protos, d_queries, None, None, None


def pack_keep(keep: torch.Tensor) -> torch.Tensor:
    """bool [.., Nq, Nq] -> little-endian bit words int32 [.., Nq, ceil(Nq/32)] (host or device)."""
    nq = keep.shape[-1]
    words = (nq + 31) // 32
    pad = words * 32 - nq
    k = torch.nn.functional.pad(keep.to(torch.int64), (0, pad)).view(*keep.shape[:-1], words, 32)
    weights = (1 << torch.arange(32, dtype=torch.int64, device=keep.device))
    packed = (k * weights).sum(-1)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)
    return packed.to(torch.int32)


def cpl_loss(protos, queries, labels, temperature: float, keep: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CPL loss per episode.  ``keep``: bool [E,Nq,Nq] / [Nq,Nq] sampled-negative mask or None for
    "all queries of the other classes" (M >= per-class count)."""
    q, l, had = _batched(queries, labels)
    p = protos if protos.dim() == 3 else protos.unsqueeze(0)
    _same_dim("cpl_loss", q, p)
    kp = None
    if keep is not None:
        kp = keep if keep.dim() == 3 else keep.unsqueeze(0)
        if kp.dtype == torch.bool:
            kp = pack_keep(kp)
        kp = kp.to(device=q.device, dtype=torch.int32).contiguous()
    loss = _Cpl.apply(_f32(p), _f32(q), _i32(l), kp, temperature)
    return loss if had else loss[0]


# ---------------------------------------------------------------------------------- angular
class _Angular(torch.autograd.Function):
    @staticmethod
    def forward(ctx, protos, queries, labels, angle, alpha, anchors, normalize_ref):
        e, w, d = protos.shape
        nq = queries.shape[1]
        loss = torch.empty(e, device=protos.device, dtype=torch.float32)
        call("afsl_angular_fwd_f32", ptr(protos), ptr(queries), ptr(labels), float(angle), float(alpha), int(anchors),
             int(normalize_ref), ptr(loss), e, nq, w, d, stream_ptr())
        ctx.save_for_backward(protos, queries, labels)
        ctx.args = (float(angle), float(alpha), int(anchors), int(normalize_ref))
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        protos, queries, labels = ctx.saved_tensors
        e, w, d = protos.shape
        nq = queries.shape[1]
        d_loss = _f32(d_loss)
        d_protos, d_queries = torch.empty_like(protos), torch.empty_like(queries)
        call("afsl_angular_bwd_f32", ptr(protos), ptr(queries), ptr(labels), *ctx.args, ptr(d_loss), ptr(d_protos),
             ptr(d_queries), e, nq, w, d, stream_ptr())
        return d_protos, d_queries, None, None, None, None, None


def angular_loss(protos, queries, labels, miner_angle_deg: float, alpha_deg: float = 40.0,
                 prototypes_as_anchors: bool = True, normalize_ref: bool = False) -> torch.Tensor:
    """Angular loss with angular mining per episode (AngularLossClass, loops/loss.py:39-97)."""
    q, l, had = _batched(queries, labels)
    p = protos if protos.dim() == 3 else protos.unsqueeze(0)
    _same_dim("angular_loss", q, p)
    loss = _Angular.apply(_f32(p), _f32(q), _i32(l), miner_angle_deg, alpha_deg, prototypes_as_anchors, normalize_ref)
    return loss if had else loss[0]


# ---------------------------------------------------------------------------------- SpecAugment
_ROW_TABLES = {}


def _row_tables(rows: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Source row / blend weight of grid_sample's y axis for grid y = linspace(-1, 1, rows)
    (utils/augmentations.py:142-146), computed once on the host with the same fp32 ops."""
    key = (rows, str(device))
    if key not in _ROW_TABLES:
        gy = torch.linspace(-1, 1, rows)
        iy = ((gy + 1) / 2) * (rows - 1)
        lo = torch.floor(iy)
        _ROW_TABLES[key] = (lo.to(torch.int32).to(device), (iy - lo).to(device))
    return _ROW_TABLES[key]


@torch.no_grad()
def specaug_views(x: torch.Tensor, warp_p: torch.Tensor, warp_d: torch.Tensor, time_masks: torch.Tensor,
                  freq_masks: torch.Tensor, mask_value: float, set_size: int, src_x: Optional[torch.Tensor] = None,
                  views_mask: int = 0b1111, out: Optional[torch.Tensor] = None,
                  set_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [N,1,F,T] -> views [4,N,1,F,T] (copy, time-warp, time-mask, freq-mask).

    warp_p/warp_d: int [N]; time_masks/freq_masks: int [sets,num_mask,2] = (start, length) with
    sets = N / set_size, or ``set_ids`` int [N] naming each sample's set (ragged sets).  Views whose bit is clear
    in ``views_mask`` are left untouched.
    """
    n, c, f, t = x.shape
    if c != 1:
        raise ValueError("SpecAugment expects [batch, 1, freq, time]")
    x = _f32(x)
    dev = x.device
    views = out if out is not None else torch.empty(4, n, 1, f, t, device=dev, dtype=torch.float32)
    lo, w = _row_tables(f, dev)
    tm, fm = _i32(time_masks.to(dev)), _i32(freq_masks.to(dev))
    num_mask = tm.shape[-2] if tm.numel() else 0
    # converted temporaries must stay referenced until the launch has been issued
    wp, wd = _i32(warp_p.to(dev)), _i32(warp_d.to(dev))
    sx = _f32(src_x.to(dev)) if src_x is not None else None
    sid = _i32(set_ids.to(dev)) if set_ids is not None else None
    if sid is not None and sid.numel() != n:
        raise ValueError(f"set_ids has {sid.numel()} entries for {n} samples")
    call("afsl_specaug_views_f32", ptr(x), ptr(views), ptr(wp), ptr(wd), ptr(sx), ptr(lo), ptr(w), ptr(sid), ptr(tm), ptr(fm),
         int(num_mask), float(mask_value), n, int(set_size), f, t, int(views_mask), stream_ptr())
    return views


# ---------------------------------------------------------------------------------- view fusion
_FUSION_PARAMS = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                  "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias", "norm1.weight", "norm1.bias",
                  "norm2.weight", "norm2.bias")


class _ViewFusion(torch.autograd.Function):
    """x [N,V,64] -> y [N,V,64] through the packed encoder-layer weights (see include/afsl.h)."""

    @staticmethod
    def forward(ctx, x, masks, *params):
        n, v, d = x.shape
        ffn = params[4].shape[0]
        flat = [p.detach().reshape(-1) for p in params]
        transposed = [params[0].detach().t().reshape(-1), params[2].detach().t().reshape(-1),
                      params[4].detach().t().reshape(-1), params[6].detach().t().reshape(-1)]
        packed = torch.cat(flat + transposed).float().contiguous()
        y = torch.empty_like(x)
        m = [None if t is None else _f32(t) for t in masks]
        call("afsl_view_fusion_fwd_f32", ptr(x), ptr(packed), ptr(y), ptr(m[0]), ptr(m[1]), ptr(m[2]), ptr(m[3]), n, v, d, ffn,
             stream_ptr())
        ctx.save_for_backward(x, packed, *[t for t in m if t is not None])
        ctx.mask_present = [t is not None for t in m]
        ctx.shapes = [p.shape for p in params]
        ctx.dims = (n, v, d, ffn)
        return y

    @staticmethod
    def backward(ctx, d_y):
        x, packed, *present = ctx.saved_tensors
        it = iter(present)
        m = [next(it) if has else None for has in ctx.mask_present]
        n, v, d, ffn = ctx.dims
        lib = _lib.load()
        count = int(lib.afsl_view_fusion_param_floats())
        grid = int(lib.afsl_view_fusion_grid(n, v))
        d_y = _f32(d_y)
        d_x = torch.empty_like(x)
        partial = torch.zeros(grid, count, device=x.device, dtype=torch.float32)
        call("afsl_view_fusion_bwd_f32", ptr(x), ptr(packed), ptr(d_y), ptr(m[0]), ptr(m[1]), ptr(m[2]), ptr(m[3]), ptr(d_x),
             ptr(partial), n, v, d, ffn, stream_ptr())
        total = partial.sum(0)
        grads, off = [], 0
        for shape in ctx.shapes:
            k = int(torch.Size(shape).numel())
            grads.append(total[off:off + k].view(shape))
            off += k
        return (d_x, None, *grads)


def view_fusion_supported(layer: torch.nn.TransformerEncoderLayer, views: int) -> bool:
    att = layer.self_attn
    return (att.embed_dim == 64 and att.num_heads == 1 and layer.linear1.out_features == 256 and not layer.norm_first
            and views in (1, 2, 4, 8) and getattr(layer, "activation_relu_or_gelu", 1) == 1
            and abs(layer.norm1.eps - 1e-5) < 1e-12 and att.in_proj_weight is not None)


def view_fusion(x: torch.Tensor, layer: torch.nn.TransformerEncoderLayer, masks=None) -> torch.Tensor:
    """One post-norm encoder layer over the views: x [.., V, 64] -> [.., V, 64] on the libafsl kernel.

    ``masks``: optional 4-tuple of keep masks scaled by 1/(1-p) for the dropout sites (attention
    probabilities [N,V,V], after out_proj [N,V,64], after ReLU [N,V,256], after linear2 [N,V,64]).
    When omitted they are drawn with torch if the layer is in training mode with p > 0."""
    lead = x.shape[:-2]
    v, d = x.shape[-2:]
    xf = _f32(x.reshape(-1, v, d))
    n = xf.shape[0]
    if masks is None:
        masks = (None, None, None, None)
        if layer.training:
            def draw(p, *shape):
                if p <= 0.0:
                    return None
                return (torch.rand(*shape, device=xf.device) >= p).float() / (1.0 - p)
            masks = (draw(layer.self_attn.dropout, n, v, v), draw(layer.dropout1.p, n, v, d),
                     draw(layer.dropout.p, n, v, layer.linear1.out_features), draw(layer.dropout2.p, n, v, d))
    named = dict(layer.named_parameters())
    params = [named[k] for k in _FUSION_PARAMS]
    y = _ViewFusion.apply(xf, tuple(masks), *params)
    return y.view(*lead, v, d)


# ---------------------------------------------------------------------------------- grouped BN + ReLU + pool
def _is_nhwc(t: torch.Tensor) -> bool:
    """Dense channels-last 4-d tensor that is not also NCHW-contiguous (1x1 planes count as NCHW)."""
    return (t.dim() == 4 and t.shape[1] % 4 == 0 and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last))


class _GbnReluPool(torch.autograd.Function):
    """y = MaxPool3(ReLU(BN(u + conv_bias))) with u the bias-free convolution output.

    mean/rstd refer to u (training: batch statistics of u per group; eval: running_mean - conv_bias),
    so the convolution bias never has to be added to the full-resolution tensor."""

    @staticmethod
    def forward(ctx, u, gamma, beta, conv_bias, mean, rstd, groups, group, per_group):
        n, c, h, w = u.shape
        nhwc = _is_nhwc(u)
        if nhwc:
            y = torch.empty(n, c, h // 3, w // 3, device=u.device, dtype=torch.float32, memory_format=torch.channels_last)
            call("afsl_gbn_relu_pool_nhwc_fwd_f32", ptr(u, True), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(y, True),
                 groups, group, c, h, w, int(per_group), stream_ptr())
        else:
            y = torch.empty(n, c, h // 3, w // 3, device=u.device, dtype=torch.float32)
            call("afsl_gbn_relu_pool_fwd_f32", ptr(u), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(y), groups, group, c,
                 h, w, int(per_group), stream_ptr())
        ctx.save_for_backward(u, gamma, beta, mean, rstd)
        ctx.dims = (groups, group, int(per_group))
        ctx.has_bias = conv_bias is not None
        ctx.nhwc = nhwc
        return y

    @staticmethod
    def backward(ctx, d_y):
        u, gamma, beta, mean, rstd = ctx.saved_tensors
        groups, group, per_group = ctx.dims
        n, c, h, w = u.shape
        sums = torch.empty(groups, c, 2, device=u.device, dtype=torch.float32)
        if ctx.nhwc:
            d_y = d_y.float().contiguous(memory_format=torch.channels_last)
            d_u = torch.empty_like(u, memory_format=torch.channels_last)
            parts = int(_lib.load().afsl_gbn_nhwc_parts(groups))
            ws = torch.empty(groups, parts, c, 2, device=u.device, dtype=torch.float64)
            call("afsl_gbn_relu_pool_nhwc_bwd_f32", ptr(u, True), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(d_y, True),
                 ptr(d_u, True), ptr(ws), parts, ptr(sums), groups, group, c, h, w, per_group, stream_ptr())
        else:
            d_y = _f32(d_y)
            d_u = torch.empty_like(u)
            call("afsl_gbn_relu_pool_bwd_f32", ptr(u), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(d_y), ptr(d_u),
                 ptr(sums), groups, group, c, h, w, per_group, stream_ptr())
        totals = sums.sum(0)
        d_gamma, d_beta = totals[:, 1].contiguous(), totals[:, 0].contiguous()
        d_bias = None
        if ctx.has_bias:
            # batch statistics remove any per-channel constant: the gradient is exactly zero (the eager
            # chain obtains round-off noise by summing d_u); with running statistics it is gamma*rstd*sum(dz)
            d_bias = torch.zeros_like(gamma) if per_group else gamma * rstd * d_beta
        return d_u, d_gamma, d_beta, d_bias, None, None, None, None, None


def gbn_relu_pool(u: torch.Tensor, bn: torch.nn.BatchNorm2d, group_size: Optional[int] = None,
                  conv_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """MaxPool3(ReLU(BatchNorm(u + conv_bias))) for a convolution output ``u [N,C,H,W]``.

    Training mode: batch statistics per group of ``group_size`` consecutive samples (default: the
    whole batch), running statistics updated group by group like that many separate module calls.
    Eval mode: running statistics.  ``conv_bias`` (per channel) is folded into the statistics instead
    of being added to the tensor.  One libafsl launch per pass instead of the eager chain.
    """
    nhwc = _is_nhwc(u)
    u = u.float() if nhwc else _f32(u)                   # channels-last activations stay channels-last
    n, c, h, w = u.shape
    gamma, beta = _f32(bn.weight), _f32(bn.bias)
    use_batch_stats = bn.training or not bn.track_running_stats
    if use_batch_stats:
        group = int(group_size) if group_size else n
        if n % group:
            raise ValueError(f"batch of {n} samples is not a whole number of groups of {group}")
        groups = n // group
        mean = torch.empty(groups, c, device=u.device, dtype=torch.float32)
        rstd, var = torch.empty_like(mean), torch.empty_like(mean)
        if nhwc:
            parts = int(_lib.load().afsl_gbn_nhwc_parts(groups))
            ws = torch.empty(groups, parts, c, 2, device=u.device, dtype=torch.float64)
            call("afsl_gbn_stats_nhwc_f32", ptr(u, True), ptr(ws), parts, ptr(mean), ptr(rstd), ptr(var), groups, group, c, h, w,
                 float(bn.eps), stream_ptr())
        else:
            call("afsl_gbn_stats_f32", ptr(u), ptr(mean), ptr(rstd), ptr(var), groups, group, c, h, w, float(bn.eps),
                 stream_ptr())
        if bn.training and bn.track_running_stats:
            bn_running_update(bn, mean, var, float(group * h * w), shift=conv_bias)
        return _GbnReluPool.apply(u, gamma, beta, conv_bias, mean, rstd, groups, group, True)
    mean = _f32(bn.running_mean) if conv_bias is None else _f32(bn.running_mean - conv_bias)
    rstd = torch.rsqrt(bn.running_var.float() + bn.eps)
    return _GbnReluPool.apply(u, gamma, beta, conv_bias, mean, rstd, 1, n, False)


# ---------------------------------------------------------------------------------- layout switch NHWC -> NCHW
class _Transpose(torch.autograd.Function):
    """[N, A, B] -> [N, B, A] (contiguous), one libafsl launch; the backward is the same kernel with A and B swapped."""

    @staticmethod
    def forward(ctx, x):
        n, a, b = x.shape
        x = _f32(x)
        y = torch.empty(n, b, a, device=x.device, dtype=torch.float32)
        call("afsl_transpose_f32", ptr(x), ptr(y), n, a, b, stream_ptr())
        return y

    @staticmethod
    def backward(ctx, dy):
        n, b, a = dy.shape
        dy = _f32(dy)
        dx = torch.empty(n, a, b, device=dy.device, dtype=torch.float32)
        call("afsl_transpose_f32", ptr(dy), ptr(dx), n, b, a, stream_ptr())
        return dx


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    """A channels-last activation ``x [N,C,H,W]`` as an NCHW-contiguous tensor (and, in the backward, the NCHW gradient
    back to channels-last).  cuDNN's fp32 convolutions of encoder stages 2-4 run ~25 % faster on NCHW tensors on sm_100
    (tools/conv_fp32_probe.py); the TF32 kernels are channels-last native, so only the fp32 path switches."""
    n, c, h, w = x.shape
    rows = x.permute(0, 2, 3, 1).reshape(n, h * w, c)            # a view: channels-last memory is [N][H*W][C]
    return _Transpose.apply(rows).view(n, c, h, w)


# ---------------------------------------------------------------------------------- BatchNorm running statistics
@torch.no_grad()
def bn_running_update(bn, mean: torch.Tensor, var_biased: torch.Tensor, count: float, shift: Optional[torch.Tensor] = None) -> None:
    """The momentum updates that one module call per group would make, in group order (one launch).
    ``mean`` / ``var_biased`` [G,C] batch statistics, ``count`` elements per (group, channel), ``shift`` [C] an optional
    convolution bias that was folded out of ``mean``."""
    if bn.momentum is None:
        raise NotImplementedError("momentum=None (cumulative average) is not used by the reference encoders")
    groups, c = mean.shape
    sh = _f32(shift.detach()) if shift is not None else None
    call("afsl_bn_running_update_f32", ptr(_f32(mean)), ptr(_f32(var_biased)), ptr(sh), ptr(bn.running_mean), ptr(bn.running_var),
         ptr(bn.num_batches_tracked), float(bn.momentum), float(count / max(count - 1, 1)), groups, c, stream_ptr())


# ---------------------------------------------------------------------------------- fused encoder stage 1
# stage 1 writes its pooled output channels-last, which keeps stages 2-4 (cuDNN's NHWC-native sm_100 convolutions +
# the channels-last BatchNorm/ReLU/pool kernels) free of layout-conversion kernels; False = NCHW everywhere (A/B switch)
STAGE1_CHANNELS_LAST = True




class _Stage1(torch.autograd.Function):
    """y = MaxPool3(ReLU(BN(conv3x3(x) + bias))) for a 1-channel input, batch statistics per group."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, a, b, mean_u, rstd, s_mom, r_mom, groups, group, per_group):
        n, _, h, w = x.shape
        c = weight.shape[0]
        nhwc = STAGE1_CHANNELS_LAST and h // 3 > 1 and w // 3 > 1
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        y = torch.empty(n, c, h // 3, w // 3, device=x.device, dtype=torch.float32, memory_format=fmt)
        w9 = weight.detach().reshape(c, 9).float().contiguous()
        # window argmax codes (1 byte per pooled output) let the backward skip recomputing the 3x3 windows
        need_grad = any(ctx.needs_input_grad)
        arg = torch.empty(n, c, h // 3, w // 3, device=x.device, dtype=torch.uint8, memory_format=fmt) if need_grad else None
        call("afsl_stage1_fwd_f32", ptr(x), ptr(w9), ptr(a), ptr(b), ptr(y, nhwc), ptr(arg, nhwc), groups, group, h, w,
             int(per_group), int(nhwc), stream_ptr())
        ctx.save_for_backward(x, w9, gamma, a, b, mean_u, rstd, s_mom, r_mom, arg)
        ctx.dims = (groups, group, int(per_group), bias is not None)
        ctx.nhwc = nhwc
        return y

    @staticmethod
    def backward(ctx, d_y):
        x, w9, gamma, a, b, mean_u, rstd, s_mom, r_mom, arg = ctx.saved_tensors
        groups, group, per_group, has_bias = ctx.dims
        n, _, h, w = x.shape
        c = w9.shape[0]
        nhwc = ctx.nhwc
        d_y = d_y.float().contiguous(memory_format=torch.channels_last) if nhwc else _f32(d_y)
        sms = torch.cuda.get_device_properties(x.device).multi_processor_count
        tiles = group * ((h // 3 + 7) // 8)
        if nhwc:
            # two 512-thread CTAs per SM: pick the split whose G*parts CTAs fill whole waves of 2*SMs best
            slots = 2 * sms
            eff = lambda pr: groups * pr / (-(-groups * pr // slots) * slots)
            cand = range(1, min(tiles, 32) + 1)
            parts = next((pr for pr in cand if pr * groups >= slots and eff(pr) >= 0.95), max(cand, key=eff))
        else:
            parts = max(1, min(tiles, (2 * sms + groups - 1) // groups))
        partial = torch.empty(groups, parts, c, 11, device=x.device, dtype=torch.float32)
        call("afsl_stage1_bwd_f32", ptr(x), ptr(w9), ptr(a), ptr(b), ptr(mean_u), ptr(rstd), ptr(d_y, nhwc), ptr(arg, nhwc),
             ptr(partial), parts, groups, group, h, w, per_group, int(nhwc), stream_ptr())
        # dW_c[k] = sum_g a_gc [ T_gck - m1 S_gk - m2 rstd (sum_l w_cl R_glk - mean S_gk) ], d_gamma = sum s2, d_beta = sum s1
        # (one glue kernel; d_bias is exactly zero under batch statistics, gamma*rstd*sum(dz) under running ones)
        d_w = torch.empty(c, 9, device=x.device, dtype=torch.float32)
        d_gamma, d_beta = torch.empty(c, device=x.device), torch.empty(c, device=x.device)
        d_bias = torch.empty(c, device=x.device) if has_bias else None
        call("afsl_stage1_dw_f32", ptr(partial), parts, groups, ptr(s_mom) if per_group else None, ptr(r_mom) if per_group else None,
             ptr(w9), ptr(a), ptr(mean_u), ptr(rstd), float(group * h * w), per_group, ptr(d_w), ptr(d_gamma), ptr(d_beta),
             ptr(d_bias), stream_ptr())
        return (None, d_w.view(c, 1, 3, 3), d_bias, d_gamma, d_beta) + (None,) * 9


def stage1_supported(conv: torch.nn.Conv2d, x: torch.Tensor) -> bool:
    return (conv.in_channels == 1 and conv.out_channels == 64 and tuple(conv.kernel_size) == (3, 3)
            and tuple(conv.padding) == (1, 1) and tuple(conv.stride) == (1, 1) and tuple(conv.dilation) == (1, 1)
            and conv.groups == 1 and not x.requires_grad and x.shape[-1] >= 3 and x.shape[-2] >= 3 and x.shape[-1] <= 1024)


def stage1_conv_bn_relu_pool(x: torch.Tensor, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d,
                             group_size: Optional[int] = None) -> torch.Tensor:
    """First encoder block on a 1-channel spectrogram batch ``x [N,1,H,W]`` in two libafsl launches
    (input moments, fused conv+BN+ReLU+pool), the 64-channel full-resolution tensor never hits HBM."""
    x = _f32(x)
    n, _, h, w = x.shape
    c = conv.out_channels
    gamma, beta = _f32(bn.weight), _f32(bn.bias)
    bias = conv.bias
    use_batch_stats = bn.training or not bn.track_running_stats
    if use_batch_stats:
        group = int(group_size) if group_size else n
        if n % group:
            raise ValueError(f"batch of {n} samples is not a whole number of groups of {group}")
        groups = n // group
        sms = torch.cuda.get_device_properties(x.device).multi_processor_count
        parts = max(1, (4 * sms + groups - 1) // groups)
        mom = torch.empty(groups, parts, 54, device=x.device, dtype=torch.float64)
        call("afsl_stage1_moments_f64", ptr(x), ptr(mom), parts, groups, group, h, w, stream_ptr())
        # mean_c = w_c.S / m, E[u_c^2] = w_c^T R w_c / m -> mean, variance, rstd and the folded affine, one glue kernel
        m = float(group * h * w)
        w9 = conv.weight.detach().reshape(c, 9).float().contiguous()
        s_mom = torch.empty(groups, 9, device=x.device, dtype=torch.float64)
        r_mom = torch.empty(groups, 9, 9, device=x.device, dtype=torch.float64)
        mean_u, var, rstd, a, b = (torch.empty(groups, c, device=x.device, dtype=torch.float32) for _ in range(5))
        call("afsl_stage1_finalize_f64", ptr(mom), parts, ptr(w9), ptr(gamma), ptr(beta), float(bn.eps), m, ptr(s_mom),
             ptr(r_mom), ptr(mean_u), ptr(var), ptr(rstd), ptr(a), ptr(b), groups, stream_ptr())
        if bn.training and bn.track_running_stats:
            bn_running_update(bn, mean_u, var, m, shift=bias)
        return _Stage1.apply(x, conv.weight, bias, gamma, beta, a, b, mean_u, rstd, s_mom, r_mom, groups, group, True)
    rstd = torch.rsqrt(bn.running_var.double() + bn.eps)
    mean_u = bn.running_mean.double() - (bias.detach().double() if bias is not None else 0.0)
    a = gamma.double() * rstd
    b = beta.double() - mean_u * a
    empty = torch.empty(0, device=x.device, dtype=torch.float64)
    return _Stage1.apply(x, conv.weight, bias, gamma, beta, a.float().contiguous(), b.float().contiguous(),
                         mean_u.float().contiguous(), rstd.float().contiguous(), empty, empty, 1, n, False)


# ---------------------------------------------------------------------------------- waveform front end
def log_mel_supported(mel_transform) -> bool:
    """True for a torchaudio MelSpectrogram configured as the reference's (src/train_test.py:123-129)."""
    spec, scale = getattr(mel_transform, "spectrogram", None), getattr(mel_transform, "mel_scale", None)
    if spec is None or scale is None:
        return False
    return (spec.n_fft == 1024 and spec.win_length == 1024 and spec.hop_length > 0 and spec.power == 2.0 and spec.center
            and spec.pad_mode == "reflect" and not spec.normalized and spec.onesided and spec.pad == 0
            and tuple(scale.fb.shape) == (513, 128))


def _mel_bands(mel_transform, device):
    """The transform's dense [513,128] filterbank as bands (first bin, length, offset, packed weights), cached on it."""
    cache = getattr(mel_transform, "_afsl_bands", None)
    if cache is None or cache[0].device != torch.device(device):
        fb = mel_transform.mel_scale.fb.detach().float().cpu()
        starts, lens, offs, weights, off = [], [], [], [], 0
        for m in range(fb.shape[1]):
            nz = torch.nonzero(fb[:, m]).flatten()
            lo, hi = (int(nz[0]), int(nz[-1])) if nz.numel() else (0, -1)
            starts.append(lo); lens.append(hi - lo + 1); offs.append(off)
            weights.append(fb[lo:hi + 1, m])
            off += hi - lo + 1
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=device)
        cache = (i32(starts), i32(lens), i32(offs), torch.cat(weights + [torch.zeros(4)]).to(device).contiguous(),
                 mel_transform.spectrogram.window.detach().float().to(device).contiguous())
        mel_transform._afsl_bands = cache
    return cache


@torch.no_grad()
def log_mel(wave: torch.Tensor, mel_transform, mean: float = 0.0, std: float = 1.0) -> torch.Tensor:
    """wave [N,L] -> normalised log-mel spectrogram in dB [N,1,128,T] in one libafsl launch:
    ``((20/2 * log10(mel_transform(wave) + eps)) - mean) / std`` with eps = float32 machine epsilon
    (datasets/batch_creation.py:138-143,215-218)."""
    if not log_mel_supported(mel_transform):
        raise NotImplementedError("log_mel needs the reference's MelSpectrogram (n_fft 1024, 128 mels, power 2, centred, reflect)")
    wave = _f32(wave)
    n, length = wave.shape
    hop = int(mel_transform.spectrogram.hop_length)
    t = length // hop + 1
    starts, lens, offs, weights, window = _mel_bands(mel_transform, wave.device)
    out = torch.empty(n, 1, 128, t, device=wave.device, dtype=torch.float32)
    call("afsl_logmel_f32", ptr(wave), ptr(window), ptr(starts), ptr(lens), ptr(offs), ptr(weights), ptr(out), n, length, t, 1024,
         hop, 128, float(torch.finfo(torch.float32).eps), float(mean), float(std), stream_ptr())
    return out


# ---------------------------------------------------------------------------------- majority vote
@torch.no_grad()
def eval_vote(pred, clip_ids, labels, posterior, seg_offsets, tie_strategy: str = "min_label"):
    """Per task (#correct clips, #clips) of the multi-segment majority vote; all inputs packed [rows]."""
    dev = pred.device
    off = _i32(seg_offsets.to(dev))
    e = off.numel() - 1
    correct = torch.empty(e, device=dev, dtype=torch.int32)
    clips = torch.empty(e, device=dev, dtype=torch.int32)
    pr, ci, lb, po = _i32(pred), _i32(clip_ids.to(dev)), _i32(labels.to(dev)), _f32(posterior)
    call("afsl_eval_vote_i32", ptr(pr), ptr(ci), ptr(lb), ptr(po), ptr(off), TIE_STRATEGIES.get(tie_strategy, 0),
         ptr(correct), ptr(clips), e, stream_ptr())
    return correct, clips


def launch_count() -> int:
    return _lib.launch_count()
