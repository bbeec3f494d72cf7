"""Batched episode steps: E independent episodes per launch.

The reference runs one episode per optimizer step (loops/loops.py:26-61) and one task per
evaluation call (:66-81, :250-277).  Episodes are independent (prototypes, losses, BatchNorm
batch statistics and votes are all per episode), so this module runs E of them at once:

* ``EpisodeBatch``      - E episodes of spectrograms + labels, on the host (pinned) or the device;
* ``EpisodeRunner``     - ``train_step`` (SpecAugment views -> encoder -> view fusion -> fused
                          prototype head + CPL/angular loss -> backward -> optimizer step) and
                          ``eval_step`` (single-segment accuracy or multi-segment vote);
  per-episode arithmetic equals the reference's; the optimizer step uses the mean loss over the E
  episodes (E = 1 reproduces the reference's one-step-per-episode schedule exactly).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .loops.loss import draw_keep_reference, draw_keep_vectorised
from .utils.augmentations import SpecAugment, SpecAugParams


@dataclass
class EpisodeBatch:
    support: torch.Tensor          # [E, Ns, 1, F, T] fp32
    support_labels: torch.Tensor   # [E, Ns] int64, values 0..W-1
    query: torch.Tensor            # [E, Nq, 1, F, T]
    query_labels: torch.Tensor     # [E, Nq]
    n_way: int

    @property
    def episodes(self) -> int:
        return self.support.shape[0]

    def pin(self) -> "EpisodeBatch":
        return EpisodeBatch(self.support.pin_memory(), self.support_labels.pin_memory(), self.query.pin_memory(),
                            self.query_labels.pin_memory(), self.n_way)

    def to(self, device, non_blocking: bool = True) -> "EpisodeBatch":
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return EpisodeBatch(mv(self.support), mv(self.support_labels), mv(self.query), mv(self.query_labels), self.n_way)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.support, self.support_labels, self.query, self.query_labels))


def synthetic_batch(episodes: int, n_way: int, k_shot: int, k_query: int, t_len: int, seed: int = 1234,
                    device: str = "cpu", mels: int = 128) -> EpisodeBatch:
    """N(0,1) spectrograms of the dataset's shape with labels arange(W).repeat_interleave(K)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    ns, nq = n_way * k_shot, n_way * k_query
    support = torch.randn(episodes, ns, 1, mels, t_len, generator=gen, device=device)
    query = torch.randn(episodes, nq, 1, mels, t_len, generator=gen, device=device)
    sl = torch.arange(n_way, device=device).repeat_interleave(k_shot).expand(episodes, -1).contiguous()
    ql = torch.arange(n_way, device=device).repeat_interleave(k_query).expand(episodes, -1).contiguous()
    return EpisodeBatch(support, sl, query, ql, n_way)


class EpisodeRunner:
    """Runs batches of episodes through a few-shot model according to ``experiment_config``.

    Config keys honoured (same meaning as src/train_test.py:47-80 and loops/loops.py:19-50):
    ``use_contrastive``, ``loss.l_param``, ``loss.cpl.{use,m_param,t_param}``,
    ``loss.angular.{use,angle,prototypes_as_anchors}``, ``project_prototypes`` (overrides
    ``normalize_prototypes``), ``train_query_augmentations``, ``specaug_params.*``.
    ``replay_reference_rng`` makes every host-side draw (SpecAugment parameters, view shuffle, CPL
    negatives) follow the reference's generators and order episode by episode.
    ``use_cuda_graph`` replays the device part of ``train_step`` from a CUDA graph (one capture per batch shape).
    """

    def __init__(self, model, experiment_config: dict, optimizer: Optional[torch.optim.Optimizer] = None,
                 replay_reference_rng: bool = False, use_cuda_graph: bool = False):
        self.model = model
        self.use_cuda_graph = use_cuda_graph      # capture / replay the device part of train_step per batch shape
        self._graphs: Dict[tuple, tuple] = {}
        self.launches_per_replay = 0
        self._stage: Optional[EpisodeBatch] = None     # device staging buffers of the prefetched next batch
        self._prefetched: Optional[EpisodeBatch] = None
        self._copy_stream = None
        self._pending_rnd = None                       # (batch, randomness) drawn ahead for the next train_step
        if use_cuda_graph:
            self._prefetch_done = torch.cuda.Event()
            self._inputs_consumed = torch.cuda.Event()
        self.cfg = experiment_config
        self.optimizer = optimizer
        self.replay = replay_reference_rng
        self.specaug = SpecAugment(experiment_config) if experiment_config["specaug_params"]["use"] else None
        self.concat_views = type(model).__name__ == "ContrastivePrototypicalNetworksWithoutAttention"
        self.grad_sync = None          # set by parallel.EpisodeDataParallel: all-reduce of the flat gradient

    # ------------------------------------------------------------------ views
    def _draw_views(self, e: int, n: int, t_len: int, augment: bool):
        """Host-side SpecAugment parameters of E sets (datasets/batch_creation.py:111-121), or None."""
        if self.specaug is None or not augment:
            return None
        # the warp spline is always evaluated on the host with the reference's op sequence (also in CUDA-graph mode,
        # where the result refreshes a static buffer): the warped view is the reference's up to the last bit of the blend
        return self.specaug.draw_batch(e, n, t_len, replay_reference_rng=self.replay).with_spline(t_len)

    def _views(self, spec: torch.Tensor, params, out: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """[E,N,1,F,T] -> list of V tensors [E,N,1,F,T]; ``params``: SpecAugParams (host or device tensors) or None.
        ``out`` [4, E*N, 1, F, T]: where the kernel writes the views (a slice of a buffer shared with the other set)."""
        if params is None:
            return [spec]
        e, n = spec.shape[:2]
        views = self.specaug.apply_batch(spec.reshape(e * n, *spec.shape[2:]), params, exact_spline=True, out=out)
        return [views[v].view(e, n, *spec.shape[2:]) for v in range(4)]

    def _both_views(self, batch: EpisodeBatch, rnd: Dict[str, object]):
        """Views of the support and the query set.  When both are augmented and have one shape, the two SpecAugment launches
        write into ONE buffer [8, E*N, 1, F, T] (support views, then query views): the encoder batch of ``_features`` is then
        that buffer as it stands, no concatenation copy."""
        sup, qry = batch.support, batch.query
        if rnd["sup"] is not None and rnd["qry"] is not None and sup.shape == qry.shape and sup.is_cuda:
            e, n = sup.shape[:2]
            buf = torch.empty(8, e * n, *sup.shape[2:], device=sup.device, dtype=torch.float32)
            return self._views(sup, rnd["sup"], out=buf[:4]), self._views(qry, rnd["qry"], out=buf[4:])
        return self._views(sup, rnd["sup"]), self._views(qry, rnd["qry"])

    def _features(self, s_views: List[torch.Tensor], q_views: List[torch.Tensor]):
        """Support and query features.  When both sets have the same shape and view count, ONE encoder call and one
        fusion call serve both: the encoder sees the 2V view tensors in the order support v0..v(V-1), query v0..v(V-1) -
        the group order (BatchNorm statistics per 25-sample group, running-statistics updates) of two separate calls -
        so nothing changes numerically except that every encoder parameter receives one gradient instead of two
        (no accumulation kernels) and every encoder kernel launches once with twice the work."""
        model = self.model
        if (hasattr(model, "attention_model") and len(s_views) == len(q_views) and s_views[0].dim() == 5
                and s_views[0].shape == q_views[0].shape):
            nv, e = len(s_views), s_views[0].shape[0]
            feats = model.backbone(list(s_views) + list(q_views))
            s_list, q_list = feats[:nv], feats[nv:]
            model.query_feature_list = q_list
            fused = model.attention_model(torch.cat([torch.stack(s_list, dim=-2), torch.stack(q_list, dim=-2)], dim=0))
            return fused[:e], fused[e:]
        return model.compute_features(s_views), model(q_views)

    # ------------------------------------------------------------------ training
    def _draw_step_randomness(self, batch: EpisodeBatch) -> Dict[str, object]:
        """Everything the step draws on the HOST, in the reference's order per step: SpecAugment parameters
        (support then query, batch_creation.py:113-115), the view shuffle of contrastive_forward
        (prototypical.py:66-70) and the CPL negatives (loss.py:149).  Returned as host tensors; the device part
        of the step (``_train_compute``) is a pure function of the batch and of these."""
        cfg = self.cfg
        e, ns = batch.support.shape[:2]
        nq, t_len = batch.query.shape[1], batch.support.shape[-1]
        aug_q = cfg["train_query_augmentations"]
        fused = cfg["use_contrastive"] and hasattr(self.model, "attention_model")
        use_cpl = cfg["use_contrastive"] and cfg["loss"]["cpl"]["use"]
        views = 4 if (self.specaug is not None and aug_q) else 1
        ql = batch.query_labels
        if self.concat_views:
            ql = ql.repeat(1, views)
        rnd: Dict[str, object] = {}
        if self.replay:
            # episode by episode, each generator consumed in the reference's order: SpecAugment support, SpecAugment
            # query (torch randint + NumPy), view shuffle (Python random), CPL negatives (torch randperm)
            sup, qry, perms, keeps = [], [], [], []
            m = int(cfg["loss"]["cpl"]["m_param"]) if use_cpl else 0
            ql_host = ql.cpu()
            for i in range(e):
                sup.append(self._draw_views(1, ns, t_len, True))
                qry.append(self._draw_views(1, nq, t_len, aug_q))
                if fused:
                    rest = list(range(1, views))
                    random.shuffle(rest)
                    perms.append([0] + rest)
                if use_cpl:
                    keeps.append(draw_keep_reference(ql_host[i], m))
            rnd["sup"] = SpecAugParams.cat(sup) if sup[0] is not None else None
            rnd["qry"] = SpecAugParams.cat(qry) if qry[0] is not None else None
            if fused:
                rnd["perm"] = torch.tensor(perms, dtype=torch.int64)
            if use_cpl:
                rnd["keep"] = ops.pack_keep(torch.stack(keeps))
            return rnd
        rnd["sup"] = self._draw_views(e, ns, t_len, True)
        rnd["qry"] = self._draw_views(e, nq, t_len, aug_q)
        if fused:
            perms = []
            for _ in range(e):
                rest = list(range(1, views))
                random.shuffle(rest)
                perms.append([0] + rest)
            rnd["perm"] = torch.tensor(perms, dtype=torch.int64)
        if use_cpl:
            m = int(cfg["loss"]["cpl"]["m_param"])
            if m < ql.shape[1] // batch.n_way:                 # balanced synthetic / sampled episodes
                rnd["keep"] = ops.pack_keep(draw_keep_vectorised(ql.cpu(), m, batch.n_way))
        return rnd

    @staticmethod
    def _rnd_to(rnd: Dict[str, object], device) -> Dict[str, object]:
        out = {}
        for k, v in rnd.items():
            if isinstance(v, torch.Tensor):
                out[k] = v.to(device, non_blocking=True)
            elif v is None:
                out[k] = None
            else:                                              # SpecAugParams
                mv = lambda t, dt: None if t is None else t.to(device=device, dtype=dt, non_blocking=True)
                out[k] = type(v)(mv(v.warp_p, torch.int32), mv(v.warp_d, torch.int32), mv(v.time_masks, torch.int32),
                                 mv(v.freq_masks, torch.int32), v.set_size, mv(v.set_ids, torch.int32),
                                 mv(v.src_x, torch.float32))
        return out

    def _train_compute(self, batch: EpisodeBatch, rnd: Dict[str, object]) -> Dict[str, torch.Tensor]:
        """Device part of one step: views -> encoder -> fusion -> fused head (+ CPL / angular) -> backward."""
        cfg, model = self.cfg, self.model
        s_views, q_views = self._both_views(batch, rnd)
        sl, ql = batch.support_labels, batch.query_labels
        if self.concat_views:                                   # loops/loops.py:33-37
            sl, ql = sl.repeat(1, len(s_views)), ql.repeat(1, len(q_views))
        # support + query features, then the fused head: prototypes + FSL loss in one kernel
        support_features, query_features = self._features(s_views, q_views)
        fsl, protos, _ = ops.proto_head(support_features, sl, query_features, ql, n_way=batch.n_way)
        model.prototypes, model.support_features, model.support_labels = protos, support_features, sl
        out = {"fsl_loss": fsl}
        total = fsl
        if cfg["use_contrastive"]:
            project = cfg["project_prototypes"]
            cfeats, cprotos = self._contrastive_forward_batched(project, rnd.get("perm"))
            if not project and cfg["normalize_prototypes"]:     # loops/loops.py:45-48
                cprotos = ops.l2_normalize(cprotos, eps=1e-12)
            extra = self._extra_loss(cprotos, cfeats, ql, rnd.get("keep"))
            total = fsl + cfg["loss"]["l_param"] * extra
            out["cpl_loss"] = extra
        out["loss"] = total
        total.mean().backward()
        return {k: v.detach() for k, v in out.items()}

    def train_step(self, batch: EpisodeBatch, next_batch: Optional[EpisodeBatch] = None) -> Dict[str, torch.Tensor]:
        """One optimizer step on E episodes.  Returns per-episode losses (device tensors).

        ``next_batch`` (optional, CUDA-graph mode): the batch of the following call; its host->device copy is started
        on a side stream while this step computes, so the next call finds its inputs already on the device.

        With ``use_cuda_graph`` the device part (views, encoder, head, losses, backward) is captured once per batch
        shape and replayed: the host only draws the step's randomness, refreshes the static input buffers and
        launches one graph, then the (eager) gradient all-reduce and optimizer step."""
        model = self.model
        model.train()
        device = next(model.parameters()).device
        if self._pending_rnd is not None and self._pending_rnd[0] is batch:
            rnd = self._pending_rnd[1]                      # drawn while the previous step was running on the GPU
        else:
            rnd = self._draw_step_randomness(batch)
        self._pending_rnd = None
        if self.use_cuda_graph:
            out = self._graph_step(batch, rnd, device)
            if next_batch is not None:
                self._prefetch(next_batch, device)
        else:
            if self.optimizer is not None:
                self.optimizer.zero_grad(set_to_none=True)
            out = self._train_compute(batch.to(device), self._rnd_to(rnd, device))
        if self.grad_sync is not None:
            self.grad_sync()
        if self.optimizer is not None:
            self.optimizer.step()
        if next_batch is not None:
            # the following step's host draws (same generators, same order) overlap this step's device work
            self._pending_rnd = (next_batch, self._draw_step_randomness(next_batch))
        return out

    # ------------------------------------------------------------------ CUDA-graph replay of the device part
    def _graph_step(self, batch: EpisodeBatch, rnd: Dict[str, object], device) -> Dict[str, torch.Tensor]:
        key = (tuple(batch.support.shape), tuple(batch.query.shape), batch.n_way,
               tuple(sorted(k for k, v in rnd.items() if v is not None)))
        state = self._graphs.get(key)
        if state is None:
            state = self._capture(batch, rnd, device)
            self._graphs[key] = state
        s_batch, s_rnd, graph, s_out = state
        src_batch = batch
        if self._prefetched is batch:                       # already on the device (staged by the previous call)
            torch.cuda.current_stream(device).wait_event(self._prefetch_done)
            src_batch = self._stage
        for dst, src in ((s_batch.support, src_batch.support), (s_batch.support_labels, src_batch.support_labels),
                         (s_batch.query, src_batch.query), (s_batch.query_labels, src_batch.query_labels)):
            dst.copy_(src, non_blocking=True)
        self._inputs_consumed.record(torch.cuda.current_stream(device))
        self._prefetched = None
        self._copy_rnd(s_rnd, rnd)
        graph.replay()
        return s_out

    def _prefetch(self, batch: EpisodeBatch, device) -> None:
        """Start the host->device copy of ``batch`` into the staging buffers on the copy stream; it overlaps the
        graph replay that was just launched and only waits for the previous staging buffers to have been consumed."""
        if self._stage is None or self._stage.support.shape != batch.support.shape or self._stage.query.shape != batch.query.shape:
            self._stage = EpisodeBatch(*(torch.empty_like(t, device=device) for t in (batch.support, batch.support_labels,
                                                                                      batch.query, batch.query_labels)), batch.n_way)
            self._copy_stream = torch.cuda.Stream(device=device)
        self._copy_stream.wait_event(self._inputs_consumed)
        with torch.cuda.stream(self._copy_stream):
            for dst, src in ((self._stage.support, batch.support), (self._stage.support_labels, batch.support_labels),
                             (self._stage.query, batch.query), (self._stage.query_labels, batch.query_labels)):
                dst.copy_(src, non_blocking=True)
            self._prefetch_done.record(self._copy_stream)
        self._prefetched = batch

    @staticmethod
    def _copy_rnd(dst: Dict[str, object], src: Dict[str, object]) -> None:
        for k, v in src.items():
            if isinstance(v, torch.Tensor):
                dst[k].copy_(v, non_blocking=True)
            elif v is not None:
                dst[k].warp_p.copy_(v.warp_p, non_blocking=True)
                dst[k].warp_d.copy_(v.warp_d, non_blocking=True)
                dst[k].time_masks.copy_(v.time_masks, non_blocking=True)
                dst[k].freq_masks.copy_(v.freq_masks, non_blocking=True)
                if v.src_x is not None:
                    dst[k].src_x.copy_(v.src_x, non_blocking=True)

    def _capture(self, batch: EpisodeBatch, rnd: Dict[str, object], device):
        """Static buffers + three eager warm-up steps on a side stream (cuDNN autotuning, lazy initialisation),
        then the capture.  Gradients are (re)created inside the capture so that every replay rewrites them."""
        s_batch = EpisodeBatch(*(torch.empty_like(t, device=device) for t in (batch.support, batch.support_labels, batch.query,
                                                                               batch.query_labels)), batch.n_way)
        for dst, src in ((s_batch.support, batch.support), (s_batch.support_labels, batch.support_labels),
                         (s_batch.query, batch.query), (s_batch.query_labels, batch.query_labels)):
            dst.copy_(src)
        s_rnd = self._rnd_to(rnd, device)
        # the warm-up passes must leave no trace: BatchNorm running statistics / num_batches_tracked are put back
        # afterwards, so the first batch of a new shape is folded into them once (by the first replay), not four times
        buffers = [(b, b.detach().clone()) for b in self.model.buffers()]
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                for prm in self.model.parameters():
                    prm.grad = None
                self._train_compute(s_batch, s_rnd)
            with torch.no_grad():
                for buf, saved in buffers:
                    buf.copy_(saved)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        for prm in self.model.parameters():
            prm.grad = None
        graph = torch.cuda.CUDAGraph()
        launches = ops.launch_count()
        with torch.cuda.graph(graph):
            s_out = self._train_compute(s_batch, s_rnd)
        self.launches_per_replay = ops.launch_count() - launches     # libafsl kernels inside one replay
        return s_batch, s_rnd, graph, s_out

    def _contrastive_forward_batched(self, project: bool, perm: Optional[torch.Tensor] = None):
        """contrastive_forward for E episodes with an independent view permutation per episode.

        The reference shuffles views 1..V-1 with ``random.shuffle`` once per episode
        (prototypical.py:66-70).  The E permutations are drawn up front by ``_draw_step_randomness`` (same Python
        ``random`` stream, one ``shuffle`` of a (V-1)-list per episode) and applied with one gather."""
        model = self.model
        if not hasattr(model, "attention_model"):
            return model.contrastive_forward(project)
        feats = torch.stack(model.query_feature_list, dim=-2)            # [E, N, V, D]
        e, n, v, d = feats.shape
        idx = perm.to(feats.device).view(e, 1, v, 1).expand(e, n, v, d)
        shuffled = model.attention_model(torch.gather(feats, 2, idx))
        projected = model.projection_head(shuffled)
        protos = model.projection_head(model.prototypes) if project else model.prototypes
        return projected, protos

    def _extra_loss(self, protos, feats, labels, keep=None):
        lc = self.cfg["loss"]
        if lc["cpl"]["use"]:
            return ops.cpl_loss(protos, feats, labels, float(lc["cpl"]["t_param"]), keep=keep)
        if lc["angular"]["use"]:
            return ops.angular_loss(protos, feats, labels, float(lc["angular"]["angle"]), 40.0,
                                    bool(lc["angular"]["prototypes_as_anchors"]), False)
        raise ValueError("use_contrastive is set but neither loss.cpl.use nor loss.angular.use")

    # ------------------------------------------------------------------ evaluation
    def _eval_compute(self, batch: EpisodeBatch, rnd: Dict[str, object]):
        """Device part of a single-segment evaluation step -> (#correct per task [E] int32, queries per task)."""
        model = self.model
        s_views, q_views = self._both_views(batch, rnd)
        sl, ql = batch.support_labels, batch.query_labels
        if self.concat_views:
            sl, ql = sl.repeat(1, len(s_views)), ql.repeat(1, len(q_views))
        support_features, feats = self._features(s_views, q_views)
        _, _, correct, _ = ops.proto_eval(support_features, sl, feats, ql, n_way=batch.n_way)
        return correct, ql.shape[1]

    def _eval_graph_step(self, batch: EpisodeBatch, rnd: Dict[str, object], device):
        """Single-segment evaluation replayed from a CUDA graph (one capture per batch shape), like _graph_step."""
        key = ("eval", tuple(batch.support.shape), tuple(batch.query.shape), batch.n_way,
               tuple(sorted(k for k, v in rnd.items() if v is not None)))
        state = self._graphs.get(key)
        if state is None:
            s_batch = EpisodeBatch(*(torch.empty_like(t, device=device) for t in (batch.support, batch.support_labels,
                                                                                   batch.query, batch.query_labels)), batch.n_way)
            for dst, src in ((s_batch.support, batch.support), (s_batch.support_labels, batch.support_labels),
                             (s_batch.query, batch.query), (s_batch.query_labels, batch.query_labels)):
                dst.copy_(src)
            s_rnd = self._rnd_to(rnd, device)
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._eval_compute(s_batch, s_rnd)
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                s_out = self._eval_compute(s_batch, s_rnd)
            state = (s_batch, s_rnd, graph, s_out)
            self._graphs[key] = state
        s_batch, s_rnd, graph, s_out = state
        for dst, src in ((s_batch.support, batch.support), (s_batch.support_labels, batch.support_labels),
                         (s_batch.query, batch.query), (s_batch.query_labels, batch.query_labels)):
            dst.copy_(src, non_blocking=True)
        self._copy_rnd(s_rnd, rnd)
        graph.replay()
        return s_out

    def _draw_eval_views(self, batch: EpisodeBatch, t_len: int, augment_query: bool, counts):
        """SpecAugment parameters of E evaluation tasks -> (support, query) or None where a set is not augmented.
        ``counts``: query rows per task for packed multi-segment tasks (ragged sets), None for [E,Nq,...] queries.
        With ``replay_reference_rng`` the generators are consumed task by task, support then query - the order of E
        successive ``sample_episode`` calls (datasets/batch_creation.py:111-115)."""
        if self.specaug is None:
            return None, None
        e, ns = batch.support.shape[:2]
        nq = batch.query.shape[1]
        sa = self.specaug
        if self.replay:
            sup, qry = [], []
            for i in range(e):
                sup.append(sa.draw_batch(1, ns, t_len, True))
                if augment_query:
                    qry.append(sa.draw_batch(1, nq, t_len, True) if counts is None else sa.draw_ragged([counts[i]], t_len, True))
            sup_p = SpecAugParams.cat(sup).with_spline(t_len)
            return sup_p, (SpecAugParams.cat(qry).with_spline(t_len) if augment_query else None)
        sup_p = sa.draw_batch(e, ns, t_len, False).with_spline(t_len)
        if not augment_query:
            return sup_p, None
        qry_p = sa.draw_batch(e, nq, t_len, False) if counts is None else sa.draw_ragged(counts, t_len, False)
        return sup_p, qry_p.with_spline(t_len)

    @torch.no_grad()
    def eval_step(self, batch: EpisodeBatch, augment_query: bool = False, clip_ids: Optional[torch.Tensor] = None,
                  seg_offsets: Optional[torch.Tensor] = None, tie_strategy: str = "", as_tensor: bool = False):
        """Per-task accuracies (float64 numpy, ``correct/total`` like loops/loops.py:114,277).  ``as_tensor`` returns them
        as a float64 device tensor instead, without synchronising: the host goes on to draw the next step's parameters
        while the GPU works (collect the tensors and convert once at the end).

        Single-segment: ``batch.query`` is [E,Nq,...].  Multi-segment: ``batch.query`` is
        [1,rows,...] packed over tasks, with ``seg_offsets`` [E+1] and ``clip_ids`` [rows]."""
        model = self.model
        model.eval()
        device = next(model.parameters()).device
        t_len = batch.support.shape[-1]
        if seg_offsets is None:
            sup_p, qry_p = self._draw_eval_views(batch, t_len, augment_query, None)
            rnd = {"sup": sup_p, "qry": qry_p}
            if self.use_cuda_graph:
                correct, per_task = self._eval_graph_step(batch, rnd, device)
            else:
                correct, per_task = self._eval_compute(batch.to(device), self._rnd_to(rnd, device))
            if as_tensor:
                return correct.double() / per_task
            return correct.cpu().numpy().astype(np.float64) / per_task
        batch = batch.to(device)
        # one SpecAugment draw per task, shared by all its query segments (batch_creation.py:113-115)
        counts = (seg_offsets[1:] - seg_offsets[:-1]).tolist()
        sup_p, q_params = self._draw_eval_views(batch, t_len, augment_query, counts)
        s_views = self._views(batch.support, sup_p)
        sl = batch.support_labels
        if self.concat_views:
            sl = sl.repeat(1, len(s_views))
        support_features = model.compute_features(s_views)
        q_views = self._views(batch.query, q_params)
        feats = model(q_views)[0]                                # packed rows
        ql = batch.query_labels[0]
        max_rows = int(max(counts))
        pred, post, _, _ = ops.proto_eval(support_features, sl, feats, ql, n_way=batch.n_way,
                                          q_offsets=seg_offsets.to(feats.device), max_rows=max_rows)
        correct, clips = ops.eval_vote(pred, clip_ids, ql, post, seg_offsets, tie_strategy)
        if as_tensor:
            return correct.double() / clips.double()
        return correct.cpu().numpy().astype(np.float64) / clips.cpu().numpy().astype(np.float64)
