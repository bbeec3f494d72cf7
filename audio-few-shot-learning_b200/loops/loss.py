"""Loss callables on libafsl kernels (mirror of loops/loss.py:12-165).

``FSL_Loss()``, ``CPL_Loss(T, M)`` and ``AngularLossClass(angle, prototypes_as_anchors)`` are
parameter-free ``nn.Module``s called as ``loss(prototypes, queries, labels)`` and return a 0-dim
fp32 tensor with grad, like the reference.  With a leading episode dimension
(``[E,W,D], [E,Nq,D], [E,Nq]``) they return the per-episode losses ``[E]``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops


class FSL_Loss(nn.Module):
    """Prototypical loss: mean NLL of log_softmax(-cdist(queries, prototypes)) (loops/loss.py:12-37)."""

    def forward(self, prototypes: torch.Tensor, queries: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return ops.proto_loss(prototypes, queries, labels)


def draw_keep_reference(labels_cpu: torch.Tensor, M: int) -> torch.Tensor:
    """Sampled-negative mask drawn exactly like the reference (loops/loss.py:134-165): for query i
    ascending and every other label c ascending, ``randperm(n_c)[:M]`` on the torch CPU generator."""
    uniq = labels_cpu.unique()
    groups = [(int(c), torch.where(labels_cpu == c)[0]) for c in uniq]
    n = labels_cpu.numel()
    keep = torch.eye(n, dtype=torch.bool)
    lab = labels_cpu.tolist()
    for i in range(n):
        for c, members in groups:
            if c != lab[i]:
                keep[i, members[torch.randperm(len(members))[:M]]] = True
    return keep


def draw_keep_vectorised(labels: torch.Tensor, M: int, n_way: int) -> torch.Tensor:
    """Same distribution as :func:`draw_keep_reference` for ``[E, Nq]`` labels, drawn with one
    ``rand`` + per-class ``topk`` on the labels' device (different random stream)."""
    e, n = labels.shape
    keys = torch.rand(e, n, n, device=labels.device)
    keep = torch.eye(n, dtype=torch.bool, device=labels.device).expand(e, n, n).clone()
    for c in range(n_way):
        member = (labels == c).unsqueeze(1).expand(e, n, n)                 # column j belongs to class c
        wanted = (labels != c).unsqueeze(2)                                 # row i is not of class c
        k = min(M, n)
        masked = torch.where(member, keys, torch.full_like(keys, 2.0))
        idx = masked.topk(k, dim=2, largest=False).indices
        chosen = torch.zeros_like(keep).scatter_(2, idx, True) & member & wanted
        keep |= chosen
    return keep


class CPL_Loss(nn.Module):
    """Contrastive prototype loss (loops/loss.py:99-165).

    ``T``: temperature, ``M``: sampled queries per other class.  ``replay_reference_rng`` (extension,
    default True) draws the negatives with the reference's ``randperm`` sequence so that losses,
    gradients and the state of the global torch generator afterwards are those of the reference;
    set it to False to skip host-side sampling (no randomness is needed when M >= per-class count,
    otherwise a vectorised draw on the device is used).
    """

    def __init__(self, T: float = 1.0, M: int = 5, replay_reference_rng: bool = True):
        super().__init__()
        self.T = T
        self.M = M
        self.replay_reference_rng = replay_reference_rng
        self.device = None

    def forward(self, prototypes: torch.Tensor, queries: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.device = prototypes.device
        keep = self.similarity_sampling(prototypes, queries, labels, self.M)
        return ops.cpl_loss(prototypes, queries, labels, self.T, keep=keep)

    def similarity_sampling(self, prototypes, queries, labels, M) -> Optional[torch.Tensor]:
        """Returns the keep mask (bool ``[.., Nq, Nq]``) or None when every other-class query is kept."""
        batched = labels.dim() == 2
        if self.replay_reference_rng:
            host = labels.detach().cpu()
            if batched:
                return torch.stack([draw_keep_reference(row, M) for row in host])
            return draw_keep_reference(host, M)
        lab2 = labels if batched else labels.unsqueeze(0)
        n_way = prototypes.shape[-2]
        per_class_max = int(torch.stack([(lab2 == c).sum(1).max() for c in range(n_way)]).max())
        if M >= per_class_max:
            return None
        keep = draw_keep_vectorised(lab2, M, n_way)
        return keep if batched else keep[0]


class AngularLossClass(nn.Module):
    """Angular loss with mining (loops/loss.py:39-97): AngularMiner(angle) selects triplets,
    AngularLoss (alpha = 40 degrees, the pytorch_metric_learning default - the config ``angle`` only
    parametrises the miner, loss.py:43-46) scores them; ``prototypes_as_anchors`` switches between
    the two branches of the reference.  Arithmetic restated from pytorch_metric_learning (not
    vendored by the reference; see oracle/angular.py for the provenance note)."""

    def __init__(self, angle, prototypes_as_anchors, alpha: float = 40.0, normalize_ref: bool = False):
        super().__init__()
        self.protoypes_as_anchors = prototypes_as_anchors        # (sic) attribute name of the reference
        self.angle = angle
        self.alpha = alpha
        self.normalize_ref = normalize_ref

    def forward(self, prototypes, queries, query_labels):
        num_prototypes = prototypes.size(-2)
        first = query_labels if query_labels.dim() == 1 else query_labels[0]
        assert num_prototypes == torch.unique(first).size(0)                       # loss.py:65
        return ops.angular_loss(prototypes, queries, query_labels, float(self.angle), float(self.alpha),
                                bool(self.protoypes_as_anchors), self.normalize_ref)
