"""Oracle (TEST INFRASTRUCTURE): angular loss.  **PARITY UNPINNED.**

The reference's ``AngularLossClass`` (loops/loss.py:39-97) delegates all
arithmetic to ``pytorch_metric_learning.{miners.AngularMiner, losses.AngularLoss}``.
That package is imported at loops/loss.py:5-6 but is not listed in the
reference's requirements.txt, not vendored, has no pinned version, and is not
installed in this image - so it cannot be executed here and no fixture can be
generated from it.  This file restates

* the reference wrapper line by line (anchor duplication, positive/negative
  concatenation, no ``indices_tuple`` in the anchors branch), and
* the published algorithm of pytorch-metric-learning 2.x:
  ``miners/angular_miner.py`` (all (a,p,n) triplets with
  ``atan(|a^-p^| / (2 |n^ - c^|)) > angle``, ``c = (a+p)/2``, every vector
  L2-normalised inside ``LpDistance.pairwise_distance``, which also adds
  ``eps=1e-6`` to the difference like ``F.pairwise_distance``),
  ``losses/angular_loss.py`` (``log(1 + sum_k exp(4 tan^2(alpha) (a^+p^).x_k
  - 2 (1+tan^2(alpha)) a^.p^))`` over all other-label reference rows, alpha = 40
  degrees by default, anchors/positives normalised, reference rows as given),
  ``utils/loss_and_miner_utils.py`` (``get_all_triplets_indices``,
  ``get_all_pairs_indices``, ``convert_to_pairs``, masked ``logsumexp`` with an
  appended zero), and ``reducers/MeanReducer`` (plain mean over pair rows).

Whether PML normalises the reference rows inside the loss is the one point that
could not be checked; ``normalize_ref`` exposes both readings.  They coincide
whenever the inputs are unit-norm, which holds for every BASELINE config
(queries leave ProjectionHead normalised, models/main_modules.py:253;
prototypes are normalised or projected, loops/loops.py:44-48).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

DEFAULT_ALPHA_DEG = 40.0     # AngularLoss() default; loops/loss.py:43 never overrides it


def _unit(x):
    return F.normalize(x, p=2, dim=1)


def all_triplets(labels, ref_labels, same_set: bool):
    """(a, p, n) index triples, lexicographic (PML get_all_triplets_indices)."""
    match = labels.unsqueeze(1) == ref_labels.unsqueeze(0)
    diff = ~match
    if same_set:
        match = match.clone()
        match.fill_diagonal_(False)
    trip = match.unsqueeze(2) & diff.unsqueeze(1)
    return torch.where(trip)


def mine(emb, labels, ref_emb, ref_labels, angle_deg: float, same_set: bool):
    """AngularMiner.mine: keep triplets whose angle exceeds ``angle_deg``."""
    with torch.no_grad():
        a, p, n = all_triplets(labels, ref_labels, same_set)
        anchors, positives, negatives = emb[a], ref_emb[p], ref_emb[n]
        centers = (anchors + positives) / 2
        ap = F.pairwise_distance(_unit(anchors), _unit(positives), p=2)
        nc = F.pairwise_distance(_unit(negatives), _unit(centers), p=2)
        ang = torch.atan(ap / (2 * nc))
        ok = ang > math.radians(angle_deg)
    return a[ok], p[ok], n[ok]


def _masked_lse_plus_one(x, keep):
    """PML logsumexp(keep_mask, add_one=True): LSE over kept entries and an extra 0."""
    x = x.masked_fill(~keep, torch.finfo(x.dtype).min)
    x = torch.cat([x, torch.zeros(x.shape[0], 1, dtype=x.dtype)], dim=1)
    out = torch.logsumexp(x, dim=1, keepdim=True)
    return out.masked_fill(~keep.any(dim=1, keepdim=True), 0)


def angular_loss_pairs(emb, labels, ref_emb, ref_labels, a1, p, alpha_deg=DEFAULT_ALPHA_DEG,
                       normalize_ref: bool = False):
    """AngularLoss.compute_loss for explicit (anchor, positive) index pairs."""
    if a1.numel() == 0:
        return emb.sum() * 0
    anchors, positives = _unit(emb[a1]), _unit(ref_emb[p])
    keep = labels[a1].unsqueeze(1) != ref_labels.unsqueeze(0)
    if not keep.any():
        # PML get_pairs returns None when there is no (anchor, negative) pair at all
        return emb.sum() * 0
    ref = _unit(ref_emb) if normalize_ref else ref_emb
    t2 = math.tan(math.radians(alpha_deg)) ** 2
    ap_dot = (anchors * positives).sum(dim=1, keepdim=True)
    proj = torch.matmul(anchors + positives, ref.unsqueeze(2)).squeeze(2).t()    # [pairs, R]
    form = 4 * t2 * proj - 2 * (1 + t2) * ap_dot
    return _masked_lse_plus_one(form, keep).mean()


def angular_loss_class(protos, queries, query_labels, angle_deg: float, prototypes_as_anchors: bool,
                       alpha_deg=DEFAULT_ALPHA_DEG, normalize_ref: bool = False):
    """The reference wrapper, loops/loss.py:48-97."""
    ways = protos.shape[0]
    assert ways == torch.unique(query_labels).numel()                      # loss.py:65
    proto_labels = torch.arange(ways)
    if prototypes_as_anchors:
        a, p, n = mine(protos, proto_labels, queries, query_labels, angle_deg, same_set=False)
        anchor_labels = proto_labels[a]
        emb = protos[a]                                                    # duplicated anchors, :75
        ref = torch.cat([queries[p], queries[n]])                          # :79
        ref_labels = torch.cat([query_labels[p], query_labels[n]])         # :80-82
        # no indices_tuple -> every (row, same-label ref) pair, PML get_all_pairs_indices
        match = anchor_labels.unsqueeze(1) == ref_labels.unsqueeze(0)
        a1, pp = torch.where(match)
        if a1.numel() == 0 or not (~match).any():
            return protos.sum() * 0
        return angular_loss_pairs(emb, anchor_labels, ref, ref_labels, a1, pp, alpha_deg, normalize_ref)
    emb = torch.cat([protos, queries], dim=0)                              # :89
    labels = torch.cat([proto_labels, query_labels], dim=0)                # :92
    a, p, n = mine(emb, labels, emb, labels, angle_deg, same_set=True)
    if a.numel() == 0:
        return emb.sum() * 0
    # triplets -> pairs (a, p) once per mined negative (convert_to_pairs)
    return angular_loss_pairs(emb, labels, emb, labels, a, p, alpha_deg, normalize_ref)


def anchors_closed_form(protos, queries, query_labels, m_a, w_q, alpha_deg=DEFAULT_ALPHA_DEG,
                        normalize_ref: bool = False):
    """Multiplicity-weighted closed form of the anchors branch.

    ``m_a`` = mined triplets per prototype, ``w_q`` = times query q occurs in the
    concatenated reference rows.  Used to check the algebra the CUDA kernel uses.
    """
    t2 = math.tan(math.radians(alpha_deg)) ** 2
    ph, qh = _unit(protos), _unit(queries)
    qr = qh if normalize_ref else queries
    c_pos = ph @ qh.T                         # a^.p^           [W, Nq]
    c_ref = ph @ qr.T                         # a^.x_k          [W, Nq]
    s_ref = qh @ qr.T                         # p^.x_k          [Nq, Nq]
    num = protos.new_zeros(())
    den = protos.new_zeros(())
    for a in range(protos.shape[0]):
        pos = torch.where(query_labels == a)[0]
        neg = torch.where(query_labels != a)[0]
        for q in pos:
            if w_q[q] == 0 or m_a[a] == 0:
                continue
            expo = 4 * t2 * (c_ref[a, neg] + s_ref[q, neg]) - 2 * (1 + t2) * c_pos[a, q]
            inner = (w_q[neg].to(expo.dtype) * torch.exp(expo)).sum()
            num = num + m_a[a] * w_q[q] * torch.log1p(inner)
            den = den + m_a[a] * w_q[q]
    return num / den if den > 0 else num
