"""Oracle (TEST INFRASTRUCTURE): prototypes, distances, proto loss, CPL loss.

One episode at a time, torch-CPU fp32, written to follow the reference op by
op so that rounding matches it; gradients come from torch autograd on these
restatements.  Pinned by tests/golden/head_*.npz and cpl_*.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def prototypes(support_features: torch.Tensor, support_labels: torch.Tensor) -> torch.Tensor:
    """Per-label mean of the support rows.

    Follows models/util_functions.py:6-19: the way count is the number of
    distinct labels and prototype ``w`` averages the rows whose label equals
    ``w`` (labels are assumed to be 0..W-1).
    """
    ways = int(torch.unique(support_labels).numel())
    rows = []
    for w in range(ways):
        members = (support_labels == w).nonzero()          # [K, 1]
        rows.append(support_features[members].mean(0))     # [1, D]
    return torch.cat(rows)


def l2_scores(queries: torch.Tensor, protos: torch.Tensor) -> torch.Tensor:
    """Negated Euclidean (not squared) distance, models/few_shot_classifier.py:108-116."""
    return -torch.cdist(queries, protos)


def cosine_scores(queries: torch.Tensor, protos: torch.Tensor) -> torch.Tensor:
    """Cosine logits, models/few_shot_classifier.py:118-126 (eps 1e-12 per F.normalize)."""
    return F.normalize(queries, dim=1) @ F.normalize(protos, dim=1).T


def fsl_loss(protos: torch.Tensor, queries: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Prototypical loss, loops/loss.py:24-37: mean NLL of log-softmax(-cdist)."""
    logits = (-1) * torch.cdist(x1=queries, x2=protos, p=2.0)
    return F.nll_loss(torch.log_softmax(logits, dim=-1), labels)


def evaluate_task(scores: torch.Tensor, labels: torch.Tensor):
    """(#correct, #queries) with first-index argmax on ties, loops/loops.py:66-81."""
    predicted = torch.max(scores, 1)[1]
    return int((predicted == labels).sum().item()), int(labels.numel())


def normalize_prototypes(protos: torch.Tensor) -> torch.Tensor:
    """Row L2 normalisation used on prototypes, loops/loops.py:47-48."""
    return F.normalize(protos, p=2.0, dim=1, eps=1e-12)


# --------------------------------------------------------------------------
# CPL
# --------------------------------------------------------------------------
def cpl_draw_keep(labels: torch.Tensor, m: int) -> torch.Tensor:
    """Replay the negative sampling of loops/loss.py:134-165 on the torch CPU RNG.

    For query i (ascending) and every *other* label c in ascending
    ``labels.unique()`` order the reference draws ``randperm(n_c)[:M]`` over the
    queries of class c (loss.py:149).  Returns ``keep[i, j]`` = query j is one
    of query i's sampled negatives, or j == i (the positive, loss.py:153).
    Consumes the global torch generator exactly as the reference does.
    """
    labels = labels.cpu()
    uniq = labels.unique()
    groups = {int(c): torch.where(labels == c)[0] for c in uniq}
    nq = labels.numel()
    keep = torch.zeros(nq, nq, dtype=torch.bool)
    for i in range(nq):
        for c, members in groups.items():
            if c == int(labels[i]):
                continue
            chosen = members[torch.randperm(len(members))[:m]]
            keep[i, chosen] = True
        keep[i, i] = True
    return keep


def cpl_loss_loop(protos, queries, labels, temperature: float, m: int) -> torch.Tensor:
    """CPL loss in the reference's own loop form, loops/loss.py:118-165.

    Row i of the similarity matrix holds cos(proto[y_i], sampled negatives...)
    followed by cos(proto[y_i], q_i), divided by T; the target is the last
    column; the batch-mean NLL is then divided by Nq a second time (loss.py:131).
    """
    uniq = labels.unique()
    ways = len(uniq)
    groups = {int(c): torch.where(labels == c)[0] for c in uniq}
    rows, targets = [], []
    for i in range(queries.shape[0]):
        picked = []
        for c, members in groups.items():
            if c != int(labels[i]):
                picked.append(queries[members[torch.randperm(len(members))[:m]]])
        block = torch.vstack([torch.cat(picked, dim=0), queries[i].unsqueeze(0)])
        targets.append((ways - 1) * m)
        rows.append(F.cosine_similarity(x1=protos[labels[i]], x2=block) / temperature)
    sims = torch.stack(rows)
    tgt = torch.tensor(targets)
    return (1 / queries.shape[0]) * F.nll_loss(torch.log_softmax(sims, dim=-1), tgt)


def cpl_loss_closed(protos, queries, labels, keep, temperature: float) -> torch.Tensor:
    """Closed form of the same loss given the keep mask.

    ``loss = 1/Nq^2 * sum_i [ LSE_{j in keep_i} C[y_i, j] - C[y_i, i] ]`` with
    ``C = cos(P, Q) / T`` (each vector divided by max(norm, 1e-8) as in
    F.cosine_similarity).  Summation order inside the LSE differs from the loop
    form, so agreement is to rounding, not to the bit.
    """
    nq = queries.shape[0]
    pn = protos / protos.norm(dim=1, keepdim=True).clamp_min(1e-8)
    qn = queries / queries.norm(dim=1, keepdim=True).clamp_min(1e-8)
    c = (pn @ qn.T) / temperature                       # [W, Nq]
    row = c[labels]                                      # [Nq, Nq]
    masked = row.masked_fill(~keep, float("-inf"))
    lse = torch.logsumexp(masked, dim=1)
    return (lse - row.diagonal()).sum() / (nq * nq)
