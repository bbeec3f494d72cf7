"""Encoders, view fusion and projection head (host-side mirror of models/main_modules.py).

Class names, constructor arguments and parameter names follow the reference so that
its ``model.pt`` state-dicts load unchanged.  Differences, all additive:

* every module also accepts a leading episode dimension (``[E, N, ...]``): the encoder
  then keeps **per-(episode, view, set) BatchNorm statistics** - in the reference one
  encoder call sees exactly one 25-sample set (models/main_modules.py:18-23), so
  batching E episodes into one cuDNN call must not pool their statistics;
* the convolutions stay PyTorch/cuDNN (north star); BatchNorm -> ReLU -> MaxPool of each
  stage, the view fusion layer and the L2 normalisation closing the projection head go
  through libafsl kernels when the tensors live on the GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import os

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


def floor_power(num, divisor, power):
    """``power`` successive floor divisions (models/main_modules.py:26-40)."""
    for _ in range(power):
        num = np.floor(num / divisor)
    return num


class GroupedBatchNorm2d(nn.BatchNorm2d):
    """BatchNorm2d whose batch axis may hold several independent groups.

    ``group_size`` consecutive samples form one group (= one encoder call of the
    reference) with its own batch statistics in training mode; running statistics are
    updated group by group in order, exactly as that many separate calls would.
    With ``group_size=None`` (or a single group) this is nn.BatchNorm2d.
    """

    group_size: Optional[int] = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        gs = self.group_size
        if not self.training or gs is None or x.shape[0] == gs:
            return super().forward(x)
        n, c, h, w = x.shape
        if n % gs:
            raise ValueError(f"batch of {n} samples is not a whole number of groups of {gs}")
        g = n // gs
        xg = x.view(g, gs, c, h, w)
        var, mean = torch.var_mean(xg, dim=(1, 3, 4), unbiased=False, keepdim=True)      # [g,1,c,1,1]
        y = (xg - mean) * torch.rsqrt(var + self.eps)
        y = y.view(n, c, h, w) * self.weight.view(1, c, 1, 1) + self.bias.view(1, c, 1, 1)
        if self.track_running_stats:
            with torch.no_grad():
                _update_running(self, mean.view(g, c), var.view(g, c), gs * h * w)
        return y


class GroupedBatchNorm1d(nn.BatchNorm1d):
    """BatchNorm1d counterpart of :class:`GroupedBatchNorm2d` for ``[N, C]`` inputs."""

    group_size: Optional[int] = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        gs = self.group_size
        if not self.training or gs is None or x.shape[0] == gs:
            return super().forward(x)
        n, c = x.shape
        if n % gs:
            raise ValueError(f"batch of {n} samples is not a whole number of groups of {gs}")
        g = n // gs
        xg = x.view(g, gs, c)
        var, mean = torch.var_mean(xg, dim=1, unbiased=False, keepdim=True)
        y = ((xg - mean) * torch.rsqrt(var + self.eps)).view(n, c) * self.weight + self.bias
        if self.track_running_stats:
            with torch.no_grad():
                _update_running(self, mean.view(g, c), var.view(g, c), gs)
        return y


def _update_running(bn, mean: torch.Tensor, var_biased: torch.Tensor, count: int) -> None:
    """Apply ``g`` sequential momentum updates in one shot (groups in batch order)."""
    if mean.is_cuda:
        ops.bn_running_update(bn, mean, var_biased, float(count))
        return
    g = mean.shape[0]
    m = bn.momentum
    bn.num_batches_tracked += g
    if m is None:                                   # cumulative moving average
        raise NotImplementedError("momentum=None is not used by the reference encoders")
    unbiased = var_biased * (count / max(count - 1, 1))
    decay = (1.0 - m) ** torch.arange(g - 1, -1, -1, device=mean.device, dtype=mean.dtype)   # weight of group j
    keep = (1.0 - m) ** g
    bn.running_mean.mul_(keep).add_((decay.unsqueeze(1) * mean).sum(0), alpha=m)
    bn.running_var.mul_(keep).add_((decay.unsqueeze(1) * unbiased).sum(0), alpha=m)


class ConvBlock(nn.Sequential):
    """Conv3x3(pad 1) -> BatchNorm -> ReLU -> MaxPool(pool_dim) (models/main_modules.py:43-60).

    Same children (and state-dict keys ``0.*`` / ``1.*``) as the reference's nn.Sequential.  On the GPU
    with the reference's 3x3 pooling, BatchNorm -> ReLU -> MaxPool run as ONE libafsl kernel per pass
    (per-group statistics included); the convolution is cuDNN.  Other pool sizes, or CPU tensors in the
    unit tests, take the stock module chain - the encoder is the part of the model that stays PyTorch.
    """

    def forward(self, x):
        conv, bn, relu, pool = self[0], self[1], self[2], self[3]
        is3 = _all_equal(pool.kernel_size, 3) and _all_equal(pool.stride, 3) and _all_equal(pool.padding, 0)
        if x.is_cuda and is3 and FUSED_STAGE1 and FUSED_STAGES and ops.stage1_supported(conv, x):
            # 1-channel first block: convolution fused into the BatchNorm/ReLU/pool kernel (no cuDNN call)
            return ops.stage1_conv_bn_relu_pool(x, conv, bn, getattr(bn, "group_size", None))
        if x.is_cuda and is3 and FUSED_STAGES:
            # bias-free cuDNN convolution; the bias is folded into the BatchNorm statistics by the kernel.
            # A channels-last input (stage 1 emits one) gives a channels-last output without conversion kernels.
            weight = conv.weight
            if ops._is_nhwc(x) and NCHW_FP32_CONVS and not torch.backends.cudnn.allow_tf32 and x.dtype == torch.float32:
                # opt-in (see NCHW_FP32_CONVS): the stack leaves channels-last here, once
                x = ops.nhwc_to_nchw(x)
            if ops._is_nhwc(x):
                weight = weight.contiguous(memory_format=torch.channels_last)
            u = F.conv2d(x, weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
            if u.shape[-1] >= 3 and u.shape[-2] >= 3:
                return ops.gbn_relu_pool(u, bn, getattr(bn, "group_size", None), conv_bias=conv.bias)
            x = u if conv.bias is None else u + conv.bias.view(1, -1, 1, 1)
        else:
            x = conv(x)
        return pool(relu(bn(x)))


def _all_equal(v, k) -> bool:
    return v == k if isinstance(v, int) else all(int(e) == k for e in v)


FUSED_VIEW_FUSION = True
FUSED_STAGE1 = True
FUSED_STAGES = True     # switch for A/B measurements of the fused BatchNorm-ReLU-MaxPool kernels
# fp32 convolutions of blocks 2-4 on NCHW tensors (AFSL_NCHW_FP32=1): cuDNN's NCHW fp32 kernels are ~25 % faster on sm_100
# (75.8 against 88.6 ms per 32-episode step) but less exact: embeddings 1e-5 away from the channels-last kernels',
# projection-head gradients 1e-2 from the fp32 oracle (tools/nchw_ab_probe.py, profiles/).  OFF by default: the step then
# stays channels-last, the arithmetic the parity suite validates; bench.py reports the NCHW step as a labelled variant.
NCHW_FP32_CONVS = os.environ.get("AFSL_NCHW_FP32", "0") == "1"


def conv_block(in_channels, out_channels, pool_dim):
    return ConvBlock(
        nn.Conv2d(in_channels, out_channels, 3, padding=1),
        GroupedBatchNorm2d(out_channels),
        nn.ReLU(),
        nn.MaxPool2d(kernel_size=pool_dim, stride=pool_dim),
    )


def conv_encoder(in_channels, hidden_channels, pool_dim):
    """Four conv blocks (models/main_modules.py:63-81)."""
    blocks = [conv_block(in_channels, hidden_channels, pool_dim)]
    blocks += [conv_block(hidden_channels, hidden_channels, pool_dim) for _ in range(3)]
    return nn.Sequential(*blocks)


def _logits(features: int, out_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Dropout(p=0.3), GroupedBatchNorm1d(features, eps=1e-05, momentum=0.1, affine=True),
                         nn.Linear(in_features=features, out_features=out_dim))


def _count_params(module: nn.Module) -> int:
    return int(sum(np.prod(p.size()) for p in module.parameters() if p.requires_grad))


class StandardCNN(nn.Module):
    """Conv4 backbone (models/main_modules.py:84-114)."""

    def __init__(self, in_channels, trial_shape, hidden_channels, pool_dim, out_dim):
        super().__init__()
        self.conv_encoder = conv_encoder(in_channels, hidden_channels, pool_dim)
        num_logits = int(64 * floor_power(trial_shape[2], pool_dim[0], 4) * floor_power(trial_shape[3], pool_dim[1], 4))
        self.logits = _logits(num_logits, out_dim)
        self.params = _count_params(self)
        print(f'Trainable Params: {self.params}')

    def forward(self, x):
        x = self.conv_encoder(x)
        return self.logits(x.reshape(x.size(0), -1))      # logical NCHW order whatever the memory format


class StandardHybrid(nn.Module):
    """Conv stack + recurrent layer with skip connection (models/main_modules.py:117-198)."""

    def __init__(self, in_channels, seq_layers, seq_type, bidirectional, hidden_channels, pool_dim, out_dim):
        super().__init__()
        self.bidirectional = bidirectional
        self.seq_type = seq_type
        hidden = 64
        self.conv_encoder = conv_encoder(in_channels, hidden_channels, pool_dim)
        if seq_type not in ['LSTM', 'GRU', 'RNN']:
            raise ValueError('Seq type not recognised')
        self.seq_layers = getattr(nn, seq_type)(input_size=hidden, hidden_size=hidden, num_layers=seq_layers,
                                                bidirectional=bidirectional, batch_first=True)
        self.logits = _logits(hidden, out_dim)
        self.params = _count_params(self)
        print(f'Num Layers: {seq_layers} -> Trainable Params: {self.params}')

    def many_to_one(self, t, lengths):
        if isinstance(lengths, int):                   # the only call site passes skip.shape[-2]
            return t[:, lengths - 1]
        return t[torch.arange(t.size(0), device=t.device), lengths - 1]

    def forward(self, x):
        x = self.conv_encoder(x)
        x = x.transpose(1, -1)                        # (batch, time, freq, channel)
        batch, time = x.size()[:2]
        x = x.reshape(batch, time, -1)
        output = self.seq_layers(x)[0]
        hs = self.seq_layers.hidden_size
        skip = output[:, :, :hs] + x
        if self.bidirectional:
            skip = skip + output[:, :, hs:]
        return self.logits(self.many_to_one(skip, skip.shape[-2]))


def get_backbone_model(encoder_name, model_config):
    """Build the encoder named by ``experiment_config['encoder_name']`` (models/main_modules.py:258-285).

    The reference's 'CNN' branch omits StandardCNN's required ``trial_shape`` and raises
    TypeError; here an optional ``model_config['CNN']['trial_shape']`` key supplies it (same
    TypeError without it).
    """
    cfg = model_config[encoder_name]
    if encoder_name == 'CNN':
        kwargs = dict(in_channels=cfg['in_channels'], hidden_channels=cfg['hidden_channels'],
                      pool_dim=cfg['pool_dim'], out_dim=cfg['out_dim'])
        if 'trial_shape' in cfg:
            kwargs['trial_shape'] = cfg['trial_shape']
        return StandardCNN(**kwargs)
    if encoder_name == 'Hybrid':
        return StandardHybrid(in_channels=cfg['in_channels'], seq_layers=cfg['seq_layers'], seq_type=cfg['seq_type'],
                              bidirectional=cfg['bidirectional'], hidden_channels=cfg['hidden_channels'],
                              pool_dim=cfg['pool_dim'], out_dim=cfg['out_dim'])
    raise UnboundLocalError(f"unknown encoder_name {encoder_name!r}")     # the reference fails the same way


def set_group_size(module: nn.Module, group_size: Optional[int]) -> None:
    for m in module.modules():
        if isinstance(m, (GroupedBatchNorm2d, GroupedBatchNorm1d)):
            m.group_size = group_size


def _stack_views(spec_list: Sequence[torch.Tensor], e: int, n: int) -> torch.Tensor:
    """The V view tensors [E,N,1,F,T] as one encoder batch [V*E*N,1,F,T].  Views that already are consecutive slices of one
    buffer (what the SpecAugment kernel writes: [V, E*N, 1, F, T]) are re-viewed in place; anything else is concatenated."""
    first = spec_list[0]
    if not first.requires_grad and all(
            x.is_contiguous() and x.shape == first.shape and not x.requires_grad
            and x.untyped_storage().data_ptr() == first.untyped_storage().data_ptr()
            and x.storage_offset() == first.storage_offset() + i * first.numel() for i, x in enumerate(spec_list)):
        shape = (len(spec_list) * e * n, *first.shape[2:])
        strides, acc = [], 1
        for d in reversed(shape):
            strides.append(acc)
            acc *= d
        return torch.as_strided(first, shape, tuple(reversed(strides)), first.storage_offset())
    return torch.cat([x.reshape(e * n, *x.shape[2:]) for x in spec_list], dim=0)


MAX_SAMPLES_PER_CALL = 8192      # samples per encoder call on the batched path (see EncoderModule.forward)


class EncoderModule(nn.Module):
    """Applies the encoder to every view (models/main_modules.py:10-23).

    Views shaped ``[N,1,F,T]`` are encoded one call per view like the reference.  Views shaped
    ``[E,N,1,F,T]`` (E episodes) are encoded in ONE call for all views, with BatchNorm
    statistics per (episode, view) group of N samples, and returned as ``[E,N,D]`` each.
    """

    def __init__(self, experiment_config, model_config, encoder: Optional[nn.Module] = None):
        super().__init__()
        self.experiment_config = experiment_config
        self.encoder_str = experiment_config['encoder_name']
        self.encoder = encoder if encoder is not None else get_backbone_model(self.encoder_str, model_config)

    def forward(self, spec_list: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        first = spec_list[0]
        if first.dim() == 4:
            set_group_size(self.encoder, None)
            return [self.encoder(x) for x in spec_list]
        e, n = first.shape[:2]
        stacked = _stack_views(spec_list, e, n)
        set_group_size(self.encoder, n)
        try:
            # whole groups per cuDNN call, at most MAX_SAMPLES_PER_CALL samples: BatchNorm statistics are per group (or the
            # running ones), so splitting the batch changes nothing numerically, and the 64-channel stage-2 tensors stay
            # below 2^31 elements (beyond that cuDNN's fp32 convolutions fall back to kernels ~20x slower)
            per_call = max(n, (MAX_SAMPLES_PER_CALL // n) * n)
            if stacked.shape[0] <= per_call:
                feats = self.encoder(stacked)
            else:
                feats = torch.cat([self.encoder(stacked[lo:lo + per_call]) for lo in range(0, stacked.shape[0], per_call)])
        finally:
            set_group_size(self.encoder, None)
        return [f.view(e, n, -1) for f in feats.chunk(len(spec_list), dim=0)]


class SelfAttention(nn.Module):
    """One post-norm transformer encoder layer over the V views, output ``[N, V*D]``
    (models/main_modules.py:201-228)."""

    def __init__(self, model_config):
        super().__init__()
        att = model_config['Attention']
        self.embed_dim, self.num_heads = att['embed_dim'], att['num_heads']
        self.ffn_dim, self.dropout = att['ffn_dim'], att['dropout']
        self.encoder_layer = nn.TransformerEncoderLayer(d_model=self.embed_dim, nhead=self.num_heads,
                                                        dim_feedforward=self.ffn_dim, dropout=self.dropout,
                                                        batch_first=True)

    def forward(self, x):
        lead = x.shape[:-2]                            # [N] or [E, N]
        if FUSED_VIEW_FUSION:
            # the view fusion is a head kernel of this build: no eager / CPU path is taken silently
            if not x.is_cuda:
                raise RuntimeError("SelfAttention runs on the libafsl view-fusion kernel: CUDA tensors only (no CPU fallback)")
            if not ops.view_fusion_supported(self.encoder_layer, x.shape[-2]):
                raise NotImplementedError("view-fusion kernel supports embed_dim=64, num_heads=1, ffn_dim=256, post-norm ReLU "
                                          f"and 1/2/4/8 views; got {self.embed_dim}/{self.num_heads}/{self.ffn_dim}, "
                                          f"{x.shape[-2]} views")
            y = ops.view_fusion(x, self.encoder_layer)                 # libafsl kernel (fwd + bwd)
        else:                                          # A/B switch for measurements and tests only (never set by the package)
            y = self.encoder_layer(x.reshape(-1, *x.shape[-2:]))
        return y.reshape(*lead, -1)                    # views side by side == cat(y[:, i, :])


class ProjectionHead(nn.Module):
    """Linear -> ReLU -> Linear -> L2 normalise (models/main_modules.py:231-255).
    ``ln1`` / ``ln2`` exist in the state-dict but are never applied, as in the reference."""

    def __init__(self, model_config):
        super().__init__()
        pc = model_config['Projection']
        self.input_dim, self.hidden_dim, self.output_dim = pc['input_dim'], pc['hidden_dim'], pc['output_dim']
        self.fc1 = nn.Linear(self.input_dim, self.hidden_dim)
        self.ln1 = nn.LayerNorm(self.hidden_dim)
        self.fc2 = nn.Linear(self.hidden_dim, self.output_dim)
        self.ln2 = nn.LayerNorm(self.output_dim)

    def forward(self, x):
        # three libafsl launches (fc1 + ReLU, fc2, L2 normalise), forward and backward; CUDA only, no eager fallback
        h = ops.linear(x, self.fc1.weight, self.fc1.bias, relu=True)
        return ops.l2_normalize(ops.linear(h, self.fc2.weight, self.fc2.bias), eps=1e-12)
