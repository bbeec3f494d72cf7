"""Episode-sharded data parallelism: one process per GPU, torch.distributed (NCCL on GPUs).

Episodes are independent, so rank r simply owns a contiguous block of the step's episodes
(weak scaling: per-GPU work is fixed as GPUs are added).  Training exchanges ONE flat fp32
gradient bucket per optimizer step (the whole model is ~0.24 M parameters, ~1 MB - latency
bound over NVSwitch, so a single all-reduce beats per-parameter buckets); evaluation exchanges
nothing until the end, when per-task accuracies are all-gathered so that every rank can compute
``np.mean`` / ``np.std`` over tasks in task order, bit-identical to a single-process run.
There is no collective inside the head kernels: nothing on this path has an exchange step.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` episodes owned by ``rank`` (sizes differ by at most 1)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class EpisodeDataParallel:
    """Keeps replicas in sync: broadcast parameters and buffers once, then one flat all-reduce of gradients per step.

    ``sync_gradients`` costs three launches whatever the number of parameters: one gather of all gradients into the flat
    bucket (``torch.cat(out=)``), one ``all_reduce`` (NCCL averages in the collective itself), one multi-tensor copy back
    into the ``.grad`` tensors (``torch._foreach_copy_``).  ``local_episodes`` / ``total_episodes`` weight a rank's
    gradient by its share of the step's episodes, so unequal shards (``shard_range``) still give the global episode mean.
    """

    def __init__(self, module: torch.nn.Module):
        self.module = module
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self._flat: Optional[torch.Tensor] = None
        self._views: List[torch.Tensor] = []
        self._nccl = dist.is_initialized() and dist.get_backend() == "nccl"
        if self.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0)

    def _bucket(self, like: torch.Tensor) -> None:
        total = sum(p.numel() for p in self.params)
        self._flat = torch.zeros(total, device=like.device, dtype=torch.float32)
        self._views, off = [], 0
        for p in self.params:
            self._views.append(self._flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def sync_gradients(self, local_episodes: Optional[int] = None, total_episodes: Optional[int] = None) -> None:
        """Average gradients over ranks with a single all-reduce (mean over all episodes of the step)."""
        if self.world == 1:
            return
        if self._flat is None:
            self._bucket(next(p for p in self.params))
        for p in self.params:                                   # parameters this step did not touch contribute zeros
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        torch.cat([g.reshape(-1) for g in grads], out=self._flat)
        weight = 1.0
        if local_episodes is not None and total_episodes:
            weight = local_episodes * self.world / float(total_episodes)     # 1.0 for equal shards
        if weight != 1.0:
            self._flat.mul_(weight)
        if self._nccl:
            dist.all_reduce(self._flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
            self._flat.div_(self.world)
        torch._foreach_copy_(grads, self._views)

    def sync_buffers(self) -> None:
        """Rank 0's buffers (BatchNorm running statistics) to every rank: call before evaluation / checkpointing, since
        each rank folds only its own episodes into the running statistics."""
        if self.world == 1:
            return
        for b in self.module.buffers():
            dist.broadcast(b.data, src=0)


def gather_accuracies(local_acc: np.ndarray, total: int) -> np.ndarray:
    """All ranks' per-task accuracies in task order (rank blocks are contiguous)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_acc
    world, rank = dist.get_world_size(), dist.get_rank()
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    sizes = [shard_range(total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(longest, dtype=torch.float64, device=device)
    buf[:local_acc.size] = torch.from_numpy(np.ascontiguousarray(local_acc, dtype=np.float64)).to(device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:hi - lo].cpu().numpy() for o, (lo, hi) in zip(out, sizes)])
