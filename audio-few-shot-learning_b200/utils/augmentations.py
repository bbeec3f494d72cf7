"""SpecAugment on the GPU kernel (mirror of utils/augmentations.py:21-157).

``SpecAugment(experiment_config)`` keeps the reference's constructor, attribute names and the
methods ``frequency_mask``, ``time_mask``, ``time_warp`` and ``apply_augmentations``.  Random
parameters are drawn on the HOST from the same generators in the same order as the reference
(torch global generator for the warp control points, NumPy legacy global generator for the
masks), so with equal seeds the masked views are bit-identical and the warped view uses the
same control points.  ``draw_batch`` / ``apply_batch`` are the batched extension: parameters for
many 25-sample sets at once and one kernel launch for all of them.

The waveform classes of the reference file (WaveAugment etc.) belong to the ``input_type: "wav"``
path, which is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .. import ops


@dataclass
class SpecAugParams:
    """Host-drawn parameters for ``sets`` consecutive sets of ``set_size`` samples."""
    warp_p: torch.Tensor       # int64 [sets*set_size]
    warp_d: torch.Tensor       # int64 [sets*set_size]
    time_masks: torch.Tensor   # int64 [sets, num_mask, 2]  (start, length)
    freq_masks: torch.Tensor   # int64 [sets, num_mask, 2]
    set_size: int
    set_ids: Optional[torch.Tensor] = None   # int64 [samples]: ragged sets (then set_size is unused)
    src_x: Optional[torch.Tensor] = None     # fp32 [samples, T]: the warp spline evaluated on the host (reference-exact)

    def with_spline(self, spec_len: int) -> "SpecAugParams":
        """Attach the reference-exact source coordinates of the time warp (``warp_source_x``; a few vectorised torch-CPU
        ops on [samples, T]), so that the kernel does not evaluate the spline itself."""
        self.src_x = warp_source_x(self.warp_p, self.warp_d, spec_len)
        return self

    @staticmethod
    def cat(parts: "List[SpecAugParams]") -> "SpecAugParams":
        """Parameters of several draws back to back (sets keep their order; ragged set ids are renumbered)."""
        ids = None
        if parts[0].set_ids is not None:
            offs, chunks = 0, []
            for p in parts:
                chunks.append(p.set_ids + offs)
                offs += p.time_masks.shape[0]
            ids = torch.cat(chunks)
        return SpecAugParams(torch.cat([p.warp_p for p in parts]), torch.cat([p.warp_d for p in parts]),
                             torch.cat([p.time_masks for p in parts]), torch.cat([p.freq_masks for p in parts]),
                             parts[0].set_size, ids,
                             torch.cat([p.src_x for p in parts]) if parts[0].src_x is not None else None)


class SpecAugment():

    def __init__(self, experiment_config):
        sp = experiment_config['specaug_params']
        self.time_mask_param = sp['mask_param']
        self.W = sp['W']
        self.freq_mask_param = sp['mask_param']
        self.freq_num_mask = sp['num_mask']
        self.time_num_mask = sp['num_mask']
        self.mask_value = sp['mask_value']
        self.p = sp['p']

    # ------------------------------------------------------------------ host-side draws
    def _draw_time_masks(self, time: int) -> List[List[int]]:
        cap = int(self.p * time)
        out = []
        for _ in range(self.time_num_mask):
            t = np.random.randint(1, min(self.time_mask_param, cap) + 1)
            t0 = np.random.randint(0, time - t)
            out.append([int(t0), int(t)])
        return out

    def _draw_freq_masks(self) -> List[List[int]]:
        out = []
        for _ in range(self.freq_num_mask):
            f = np.random.randint(1, self.freq_mask_param + 1)
            f0 = np.random.randint(0, 128 - f)          # 128 mel bins are hard-coded in the reference
            out.append([int(f0), int(f)])
        return out

    @staticmethod
    def _draw_warp(batch_size: int, spec_len: int, W: int):
        warp_p = torch.randint(W, spec_len - W, (batch_size,))
        warp_d = torch.randint(-W, W, (batch_size,))
        return warp_p, warp_d

    def draw_batch(self, sets: int, set_size: int, time: int, replay_reference_rng: bool = True) -> SpecAugParams:
        """Parameters for ``sets`` calls of ``apply_augmentations`` on ``set_size`` samples each.

        ``replay_reference_rng=True`` consumes the global generators call by call exactly as ``sets``
        successive reference calls would (warp points, then time masks, then frequency masks per
        set).  ``False`` draws everything with a few vectorised calls: same distributions, different
        stream, ~1000x less host time for thousands of sets.
        """
        if replay_reference_rng:
            wp, wd, tm, fm = [], [], [], []
            for _ in range(sets):
                p, d = self._draw_warp(set_size, time, self.W)
                wp.append(p); wd.append(d)
                tm.append(self._draw_time_masks(time))
                fm.append(self._draw_freq_masks())
            return SpecAugParams(torch.cat(wp), torch.cat(wd), torch.tensor(tm, dtype=torch.int64).view(sets, -1, 2),
                                 torch.tensor(fm, dtype=torch.int64).view(sets, -1, 2), set_size)
        n = sets * set_size
        warp_p = torch.randint(self.W, time - self.W, (n,))
        warp_d = torch.randint(-self.W, self.W, (n,))
        k = self.time_num_mask
        t = np.random.randint(1, min(self.time_mask_param, int(self.p * time)) + 1, size=(sets, k))
        t0 = np.floor(np.random.random_sample((sets, k)) * (time - t)).astype(np.int64)       # U{0..time-t-1}
        f = np.random.randint(1, self.freq_mask_param + 1, size=(sets, k))
        f0 = np.floor(np.random.random_sample((sets, k)) * (128 - f)).astype(np.int64)
        tm = torch.from_numpy(np.stack([t0, t], axis=-1).astype(np.int64))
        fm = torch.from_numpy(np.stack([f0, f], axis=-1).astype(np.int64))
        return SpecAugParams(warp_p, warp_d, tm, fm, set_size)

    def draw_ragged(self, set_sizes, time: int, replay_reference_rng: bool = True) -> SpecAugParams:
        """Parameters for one ``apply_augmentations`` call per entry of ``set_sizes`` (sets of different sizes packed
        back to back, e.g. the query segments of multi-segment tasks, datasets/batch_creation.py:113-115)."""
        sizes = [int(v) for v in set_sizes]
        ids = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
        if replay_reference_rng:
            wp, wd, tm, fm = [], [], [], []
            for n in sizes:
                p, d = self._draw_warp(n, time, self.W)
                wp.append(p); wd.append(d)
                tm.append(self._draw_time_masks(time))
                fm.append(self._draw_freq_masks())
            return SpecAugParams(torch.cat(wp), torch.cat(wd), torch.tensor(tm, dtype=torch.int64).view(len(sizes), -1, 2),
                                 torch.tensor(fm, dtype=torch.int64).view(len(sizes), -1, 2), 1, ids)
        rows = int(sum(sizes))
        base = self.draw_batch(len(sizes), 1, time, replay_reference_rng=False)
        return SpecAugParams(torch.randint(self.W, time - self.W, (rows,)), torch.randint(-self.W, self.W, (rows,)),
                             base.time_masks, base.freq_masks, 1, ids)

    # ------------------------------------------------------------------ kernel launches
    def apply_batch(self, spec: torch.Tensor, params: SpecAugParams, views_mask: int = 0b1111,
                    exact_spline: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """spec [N,1,F,T] on the GPU -> [4,N,1,F,T] = (copy, time-warp, time-mask, freq-mask)."""
        src_x = params.src_x
        if src_x is None and exact_spline:
            src_x = warp_source_x(params.warp_p, params.warp_d, spec.shape[-1])
        return ops.specaug_views(spec, params.warp_p, params.warp_d, params.time_masks, params.freq_masks,
                                 float(self.mask_value), params.set_size, src_x=src_x, views_mask=views_mask, out=out,
                                 set_ids=params.set_ids)

    def _single(self, spec, view, warp=None, tmasks=None, fmasks=None):
        n, t = spec.shape[0], spec.shape[-1]
        zero = torch.zeros(n, dtype=torch.int64)
        wp, wd = warp if warp is not None else (zero + 1, zero)
        tm = torch.tensor(tmasks if tmasks is not None else [], dtype=torch.int64).view(1, -1, 2)
        fm = torch.tensor(fmasks if fmasks is not None else [], dtype=torch.int64).view(1, -1, 2)
        k = max(tm.shape[1], fm.shape[1])
        pad = lambda m: torch.cat([m, torch.zeros(1, k - m.shape[1], 2, dtype=torch.int64)], 1)   # zero-length masks
        params = SpecAugParams(wp, wd, pad(tm), pad(fm), n)
        return self.apply_batch(spec, params, views_mask=1 << view, exact_spline=True)[view]

    def frequency_mask(self, spec):
        """Batch-shared frequency masks (utils/augmentations.py:33-57)."""
        return self._single(spec, 3, fmasks=self._draw_freq_masks())

    def time_mask(self, spec):
        """Batch-shared time masks of length <= min(mask_param, int(p*T)) (utils/augmentations.py:59-89)."""
        return self._single(spec, 2, tmasks=self._draw_time_masks(spec.shape[-1]))

    def time_warp(self, specs, W=50):
        """Per-sample time warp through a 3-point Hermite spline (utils/augmentations.py:110-146)."""
        return self._single(specs, 1, warp=self._draw_warp(specs.shape[0], specs.shape[-1], W))

    def apply_augmentations(self, spectrogram):
        """[original, time-warped, time-masked, frequency-masked] (utils/augmentations.py:148-157).

        One kernel launch; RNG draw order as the reference: warp points, time masks, frequency masks.
        The spline is evaluated on the host with the reference's own torch ops (tiny: N x T), so the
        warped view differs from the reference only in the last bit of the bilinear blend."""
        params = self.draw_batch(1, spectrogram.shape[0], spectrogram.shape[-1], replay_reference_rng=True)
        views = self.apply_batch(spectrogram, params, exact_spline=True)
        return [views[0], views[1], views[2], views[3]]


def warp_source_x(warp_p: torch.Tensor, warp_d: torch.Tensor, spec_len: int) -> torch.Tensor:
    """Host evaluation of the warp spline with the reference's torch op sequence
    (utils/augmentations.py:91-108,129-141): normalised source x per (sample, output column)."""
    warp_p, warp_d = warp_p.cpu(), warp_d.cpu()
    n = warp_p.numel()
    ends = torch.tensor([0, spec_len - 1])
    cx = torch.stack([ends[0].expand(n), warp_p, ends[1].expand(n)], 1)
    cy = torch.stack([torch.tensor([-1.]).expand(n), (warp_p - warp_d) * 2 / (spec_len - 1) - 1,
                      torch.tensor([1]).expand(n)], 1)
    xs = torch.linspace(0, spec_len - 1, spec_len).unsqueeze(0).expand(n, -1).contiguous()
    m = (cy[..., 1:] - cy[..., :-1]) / (cx[..., 1:] - cx[..., :-1])
    m = torch.cat([m[..., [0]], (m[..., 1:] + m[..., :-1]) / 2, m[..., [-1]]], -1)
    idx = torch.searchsorted(cx[..., 1:].contiguous(), xs)
    lo = cx.take_along_dim(idx, dim=-1)
    dx = cx.take_along_dim(idx + 1, dim=-1) - lo
    u = (xs - lo) / dx
    basis = torch.tensor([[1, 0, -3, 2], [0, 1, -2, 1], [0, 0, 3, -2], [0, 0, -1, 1]], dtype=u.dtype) @ \
        (u.unsqueeze(-2) ** torch.arange(4).view(-1, 1))
    return (basis[..., 0, :] * cy.take_along_dim(idx, dim=-1) + basis[..., 1, :] * m.take_along_dim(idx, dim=-1) * dx
            + basis[..., 2, :] * cy.take_along_dim(idx + 1, dim=-1)
            + basis[..., 3, :] * m.take_along_dim(idx + 1, dim=-1) * dx)
