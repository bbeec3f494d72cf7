"""Oracle (TEST INFRASTRUCTURE): one training / evaluation episode on CPU.

Restates the per-episode body of loops/loops.py:26-61 (train) and :66-81,
:250-277 (evaluation) on top of the oracle modules.  This is what the
``cpu_baseline`` and ``--impl reference`` legs of bench.py time, one episode
per optimizer step exactly as the reference does.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import angular, head, specaug, vote


def make_views(spec: torch.Tensor, cfg: dict, augment: bool):
    """datasets/batch_creation.py:111-121: 4 SpecAugment views or the set itself."""
    if cfg["specaug_params"]["use"] and augment:
        return specaug.apply_augmentations(spec, cfg)[0]
    return [spec]


def extra_loss(cfg: dict, protos, feats, labels):
    """Loss selection of src/train_test.py:69-80 (cpl wins over angular)."""
    loss_cfg = cfg["loss"]
    if loss_cfg["cpl"]["use"]:
        return head.cpl_loss_loop(protos, feats, labels, loss_cfg["cpl"]["t_param"], loss_cfg["cpl"]["m_param"])
    if loss_cfg["angular"]["use"]:
        return angular.angular_loss_class(protos, feats, labels, loss_cfg["angular"]["angle"],
                                          loss_cfg["angular"]["prototypes_as_anchors"])
    raise ValueError("use_contrastive without cpl or angular")


def train_step(model, optimizer, support, support_labels, query, query_labels, cfg: dict):
    """One optimizer step on one episode; returns (loss, fsl, extra) floats.  loops.py:26-61."""
    model.train()
    s_views = make_views(support, cfg, True)
    q_views = make_views(query, cfg, cfg["train_query_augmentations"])
    if type(model).__name__ == "ConcatViewsNet":                       # loops.py:33-37
        support_labels = support_labels.repeat(len(s_views))
        query_labels = query_labels.repeat(len(q_views))
    optimizer.zero_grad()
    model.process_support_set(s_views, support_labels)
    feats = model(q_views)
    fsl = head.fsl_loss(model.prototypes, feats, query_labels)
    extra_val = float("nan")
    total = fsl
    if cfg["use_contrastive"]:
        project = cfg["project_prototypes"]
        cfeats, protos = model.contrastive_forward(project)
        if not project and cfg["normalize_prototypes"]:                # loops.py:45-48
            protos = F.normalize(protos, p=2.0, dim=1, eps=1e-12)
        extra = extra_loss(cfg, protos, cfeats, query_labels)
        total = fsl + cfg["loss"]["l_param"] * extra
        extra_val = extra.item()
    total.backward()
    optimizer.step()
    return total.item(), fsl.item(), extra_val


def eval_task(model, support_views, support_labels, query_views, query_labels,
              clip_ids: Optional[torch.Tensor] = None, tie_strategy: str = ""):
    """Single-segment accuracy (loops.py:66-81,114) or multi-segment vote (:268-277)."""
    model.process_support_set(support_views, support_labels)
    with torch.no_grad():
        scores = model(query_views, inference=True)
    if clip_ids is None:
        correct, total = head.evaluate_task(scores, query_labels)
        return correct / total
    post, pred = torch.max(scores, 1)
    return vote.majority_vote_accuracy(pred, clip_ids, query_labels, post, tie_strategy)
