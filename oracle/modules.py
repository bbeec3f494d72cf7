"""Oracle (TEST INFRASTRUCTURE): plain-torch restatement of the modules around the head.

Encoders, view-fusion layer, projection head and the three few-shot model
classes, parameter names identical to the reference so that a reference
``state_dict`` loads.  Runs on CPU; used for parity fixtures and as the
single-episode CPU baseline.  Pinned by tests/golden/modules_*.npz.
"""
from __future__ import annotations

import random
from typing import List

import torch
import torch.nn.functional as F
from torch import nn

from . import head


def _shrink(n: int, pool: int, times: int = 4) -> int:
    for _ in range(times):
        n = n // pool
    return n


def _stage(cin: int, cout: int, pool) -> nn.Sequential:
    # Conv3x3(pad 1) - BatchNorm - ReLU - MaxPool(pool, stride pool); models/main_modules.py:43-60
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(),
                         nn.MaxPool2d(kernel_size=pool, stride=pool))


def _conv_stack(cin: int, hidden: int, pool) -> nn.Sequential:
    # four stages; models/main_modules.py:63-81
    return nn.Sequential(_stage(cin, hidden, pool), _stage(hidden, hidden, pool),
                         _stage(hidden, hidden, pool), _stage(hidden, hidden, pool))


class Conv4(nn.Module):
    """StandardCNN, models/main_modules.py:84-114."""

    def __init__(self, in_channels, trial_shape, hidden_channels, pool_dim, out_dim):
        super().__init__()
        self.conv_encoder = _conv_stack(in_channels, hidden_channels, pool_dim)
        flat = 64 * _shrink(trial_shape[2], pool_dim[0]) * _shrink(trial_shape[3], pool_dim[1])
        self.logits = nn.Sequential(nn.Dropout(p=0.3), nn.BatchNorm1d(flat), nn.Linear(flat, out_dim))

    def forward(self, x):
        x = self.conv_encoder(x)
        return self.logits(x.view(x.size(0), -1))


class Hybrid(nn.Module):
    """StandardHybrid, models/main_modules.py:117-198 (conv stack, recurrent layer over the
    pooled time axis with a skip connection, last step, Dropout-BN-Linear)."""

    def __init__(self, in_channels, seq_layers, seq_type, bidirectional, hidden_channels, pool_dim, out_dim):
        super().__init__()
        if seq_type not in ("LSTM", "GRU", "RNN"):
            raise ValueError("Seq type not recognised")
        self.bidirectional, self.seq_type = bidirectional, seq_type
        self.conv_encoder = _conv_stack(in_channels, hidden_channels, pool_dim)
        self.seq_layers = getattr(nn, seq_type)(input_size=64, hidden_size=64, num_layers=seq_layers,
                                                bidirectional=bidirectional, batch_first=True)
        self.logits = nn.Sequential(nn.Dropout(p=0.3), nn.BatchNorm1d(64), nn.Linear(64, out_dim))

    def forward(self, x):
        x = self.conv_encoder(x).transpose(1, -1)
        b, steps = x.shape[:2]
        x = x.reshape(b, steps, -1)
        out = self.seq_layers(x)[0]
        h = self.seq_layers.hidden_size
        x = out[:, :, :h] + out[:, :, h:] + x if self.bidirectional else out[:, :, :h] + x
        return self.logits(x[:, -1])          # many_to_one with lengths == steps, :189-197


class ViewEncoder(nn.Module):
    """EncoderModule, models/main_modules.py:10-23: one encoder call per view."""

    def __init__(self, encoder: nn.Module):
        super().__init__()
        self.encoder = encoder

    def forward(self, views: List[torch.Tensor]):
        return [self.encoder(v) for v in views]


class ViewFusion(nn.Module):
    """SelfAttention, models/main_modules.py:201-228: one post-norm encoder layer over the
    V views, then the views are laid side by side -> [N, V*D]."""

    def __init__(self, embed_dim=64, num_heads=1, ffn_dim=256, dropout=0.1):
        super().__init__()
        self.encoder_layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=num_heads,
                                                        dim_feedforward=ffn_dim, dropout=dropout,
                                                        batch_first=True)

    def forward(self, x):
        y = self.encoder_layer(x)
        return torch.cat([y[:, i, :] for i in range(y.size(1))], dim=-1)


class Projection(nn.Module):
    """ProjectionHead, models/main_modules.py:231-255 (ln1/ln2 exist but are never applied)."""

    def __init__(self, input_dim=256, hidden_dim=128, output_dim=256):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, hidden_dim)
        self.ln1 = nn.LayerNorm(hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self.ln2 = nn.LayerNorm(output_dim)

    def forward(self, x):
        return F.normalize(self.fc2(F.relu(self.fc1(x))), p=2.0, dim=1, eps=1e-12)


class FusedViewsNet(nn.Module):
    """ContrastivePrototypicalNetworks, models/prototypical.py:46-93."""

    def __init__(self, backbone, attention_model, projection_head):
        super().__init__()
        self.backbone, self.attention_model, self.projection_head = backbone, attention_model, projection_head
        self.prototypes = torch.tensor(())

    def process_support_set(self, views, labels):
        feats = self.attention_model(torch.stack(self.backbone(views), dim=1))
        self.support_features, self.support_labels = feats, labels
        self.prototypes = head.prototypes(feats, labels)

    def forward(self, views, inference=False):
        self.query_feature_list = self.backbone(views)
        feats = self.attention_model(torch.stack(self.query_feature_list, dim=1))
        return head.l2_scores(feats, self.prototypes) if inference else feats

    def contrastive_forward(self, project_prototypes):
        rest = self.query_feature_list[1:]
        random.shuffle(rest)                                   # prototypical.py:66-70
        mixed = torch.stack([self.query_feature_list[0]] + rest, dim=1)
        projected = self.projection_head(self.attention_model(mixed))
        protos = self.projection_head(self.prototypes) if project_prototypes else self.prototypes
        return projected, protos


class ConcatViewsNet(nn.Module):
    """ContrastivePrototypicalNetworksWithoutAttention, models/prototypical.py:96-126."""

    def __init__(self, backbone, projection_head):
        super().__init__()
        self.backbone, self.projection_head = backbone, projection_head
        self.prototypes = torch.tensor(())

    def process_support_set(self, views, labels):
        feats = torch.concat(self.backbone(views), dim=0)
        self.support_features, self.support_labels = feats, labels
        self.prototypes = head.prototypes(feats, labels)

    def forward(self, views, inference=False):
        self.query_feature_list = torch.concat(self.backbone(views), dim=0)
        q = self.query_feature_list
        return head.l2_scores(q, self.prototypes) if inference else q

    def contrastive_forward(self, project_prototypes):
        projected = self.projection_head(self.query_feature_list)
        protos = self.projection_head(self.prototypes) if project_prototypes else self.prototypes
        return projected, protos


def build_encoder(name: str, t_len: int) -> nn.Module:
    """Encoders at the shapes of the README model_config (README.md:384-429)."""
    if name == "CNN":
        return Conv4(1, (1, 1, 128, t_len), 64, [3, 3], 64)
    if name == "Hybrid":
        return Hybrid(1, 1, "RNN", False, 64, [3, 3], 64)
    raise ValueError(name)
