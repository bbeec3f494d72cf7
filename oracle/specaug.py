"""Oracle (TEST INFRASTRUCTURE): SpecAugment 4-view generation.

Restates utils/augmentations.py:23-157 as (a) an RNG replay that draws the
random parameters in the reference's order and (b) a deterministic apply step,
so the CUDA path can be checked on identical parameters.  torch-CPU fp32.
Pinned by tests/golden/specaug_*.npz.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class ViewParams:
    """Random parameters of one ``apply_augmentations`` call (one 25-sample set)."""
    warp_p: torch.Tensor            # int64 [N]   control point position
    warp_d: torch.Tensor            # int64 [N]   control point displacement
    time_masks: List[Tuple[int, int]]   # (t0, t) per mask, batch shared
    freq_masks: List[Tuple[int, int]]   # (f0, f) per mask, batch shared


def draw_params(n: int, t_len: int, cfg: dict) -> ViewParams:
    """Draw in the order of ``apply_augmentations`` (augmentations.py:148-152).

    time_warp first: ``torch.randint(W, T-W, (N,))`` then ``torch.randint(-W, W, (N,))``
    on the torch global generator (augmentations.py:124,128); then per time mask
    ``np.random.randint(1, min(mask_param, int(p*T)) + 1)`` and
    ``np.random.randint(0, T - t)`` (:77-84); then per frequency mask
    ``np.random.randint(1, mask_param + 1)`` and ``np.random.randint(0, 128 - f)`` (:51-52).
    """
    sp = cfg["specaug_params"]
    w = sp["W"]
    warp_p = torch.randint(w, t_len - w, (n,))
    warp_d = torch.randint(-w, w, (n,))
    cap = int(sp["p"] * t_len)
    tmasks, fmasks = [], []
    for _ in range(sp["num_mask"]):
        t = np.random.randint(1, min(sp["mask_param"], cap) + 1)
        t0 = np.random.randint(0, t_len - t)
        tmasks.append((int(t0), int(t)))
    for _ in range(sp["num_mask"]):
        f = np.random.randint(1, sp["mask_param"] + 1)
        f0 = np.random.randint(0, 128 - f)
        fmasks.append((int(f0), int(f)))
    return ViewParams(warp_p, warp_d, tmasks, fmasks)


def _hermite_basis(u: torch.Tensor) -> torch.Tensor:
    """Cubic Hermite basis values, augmentations.py:91-94 (powers 0..3 times a 4x4 matrix)."""
    powers = u.unsqueeze(-2) ** torch.arange(4).view(-1, 1)
    coeff = torch.tensor([[1, 0, -3, 2], [0, 1, -2, 1], [0, 0, 3, -2], [0, 0, -1, 1]], dtype=u.dtype)
    return coeff @ powers


def warp_source_x(warp_p: torch.Tensor, warp_d: torch.Tensor, t_len: int) -> torch.Tensor:
    """Normalised source x per (sample, output column), augmentations.py:96-108,129-141.

    Three control points ``x=[0, p, T-1]`` -> ``y=[-1, (p-d)*2/(T-1)-1, 1]``; slopes are
    the secants at the ends and their mean in the middle; evaluated at 0..T-1.
    """
    n = warp_p.numel()
    x = torch.stack([torch.tensor([0]).expand(n), warp_p, torch.tensor([t_len - 1]).expand(n)], 1)
    y = torch.stack([torch.tensor([-1.0]).expand(n), (warp_p - warp_d) * 2 / (t_len - 1) - 1,
                     torch.tensor([1]).expand(n)], 1)
    xs = torch.linspace(0, t_len - 1, t_len).unsqueeze(0).expand(n, -1)
    slope = (y[..., 1:] - y[..., :-1]) / (x[..., 1:] - x[..., :-1])
    slope = torch.cat([slope[..., [0]], (slope[..., 1:] + slope[..., :-1]) / 2, slope[..., [-1]]], -1)
    seg = torch.searchsorted(x[..., 1:].contiguous(), xs.contiguous())
    x_lo = x.take_along_dim(seg, dim=-1)
    width = x.take_along_dim(seg + 1, dim=-1) - x_lo
    hb = _hermite_basis((xs - x_lo) / width)
    return (hb[..., 0, :] * y.take_along_dim(seg, dim=-1)
            + hb[..., 1, :] * slope.take_along_dim(seg, dim=-1) * width
            + hb[..., 2, :] * y.take_along_dim(seg + 1, dim=-1)
            + hb[..., 3, :] * slope.take_along_dim(seg + 1, dim=-1) * width)


def time_warp(spec: torch.Tensor, src_x: torch.Tensor) -> torch.Tensor:
    """Bilinear resample along time, augmentations.py:142-146 (grid_sample, zeros padding,
    align_corners=True; grid y is linspace(-1, 1, rows))."""
    n, _, rows, t_len = spec.shape
    gx = src_x.view(n, 1, -1, 1).expand(-1, rows, -1, -1)
    gy = torch.linspace(-1, 1, rows).view(-1, 1, 1).expand(n, -1, t_len, -1)
    return F.grid_sample(spec, torch.cat((gx, gy), -1), align_corners=True)


def apply(spec: torch.Tensor, params: ViewParams, mask_value: float) -> List[torch.Tensor]:
    """[original, time-warped, time-masked, frequency-masked], each from the original
    (augmentations.py:148-157)."""
    t_len = spec.shape[-1]
    warped = time_warp(spec, warp_source_x(params.warp_p, params.warp_d, t_len))
    tmasked = spec.clone()
    for t0, t in params.time_masks:
        tmasked[:, :, :, t0:t0 + t] = mask_value
    fmasked = spec.clone()
    for f0, f in params.freq_masks:
        fmasked[:, :, f0:f0 + f, :] = mask_value
    return [spec.clone(), warped, tmasked, fmasked]


def apply_augmentations(spec: torch.Tensor, cfg: dict) -> Tuple[List[torch.Tensor], ViewParams]:
    """Draw + apply; consumes the torch and numpy global generators like the reference."""
    params = draw_params(spec.shape[0], spec.shape[-1], cfg)
    return apply(spec, params, cfg["specaug_params"]["mask_value"]), params
