"""The three prototypical-network models (mirror of models/prototypical.py:16-126).

Stateful API identical to the reference: ``process_support_set`` -> ``forward`` ->
``contrastive_forward``.  View lists may hold ``[N,1,F,T]`` tensors (one episode, reference
behaviour) or ``[E,N,1,F,T]`` tensors (E episodes at once).
"""
import random
from typing import Optional

import torch
from torch import Tensor

from .few_shot_classifier import FewShotClassifier


class PrototypicalNetworks(FewShotClassifier):
    """Plain ProtoNet on a tensor batch (prototypical.py:16-43)."""

    def forward(self, query_images: Tensor, inference: Optional[bool] = True) -> Tensor:
        query_features = self.compute_features(query_images)
        self._raise_error_if_features_are_multi_dimensional(query_features)
        return self.softmax_if_specified(self.l2_distance_to_prototypes(query_features))

    @staticmethod
    def is_transductive() -> bool:
        return False


class ContrastivePrototypicalNetworks(FewShotClassifier):
    """Multi-view model with self-attention view fusion (prototypical.py:46-93)."""

    def __init__(self, backbone, attention_model, projection_head, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.backbone = backbone
        self.attention_model = attention_model
        self.projection_head = projection_head

    def _fuse(self, feature_list):
        return self.attention_model(torch.stack(feature_list, dim=-2))      # [.., N, V, D] -> [.., N, V*D]

    def compute_features(self, images) -> Tensor:
        return self._fuse(self.backbone(images))

    def compute_query_features(self, images):
        self.query_feature_list = self.backbone(images)
        return self.query_feature_list

    def shuffle_augmentations(self, feature_list):
        """Keep view 0 first and shuffle the augmented views with Python's ``random`` (prototypical.py:66-70)."""
        augmentations = feature_list[1:]
        random.shuffle(augmentations)
        return torch.stack([feature_list[0]] + augmentations, dim=-2)

    def forward(self, query_images, inference=False):
        self.query_feature_list = self.compute_query_features(query_images)
        query_features = self._fuse(self.query_feature_list)
        self._raise_error_if_features_are_multi_dimensional(query_features)
        if inference == True:  # noqa: E712  (same truthiness rule as the reference)
            return self.l2_distance_to_prototypes(query_features)
        return query_features

    def contrastive_forward(self, project_prototypes):
        shuffled = self.attention_model(self.shuffle_augmentations(self.query_feature_list))
        projected_features = self.projection_head(shuffled)
        if project_prototypes == True:  # noqa: E712
            return projected_features, self.projection_head(self.prototypes)
        return projected_features, self.prototypes

    @staticmethod
    def is_transductive() -> bool:
        return False


class ContrastivePrototypicalNetworksWithoutAttention(FewShotClassifier):
    """Multi-view model that concatenates the views along the sample axis (prototypical.py:96-126);
    the caller repeats the labels V times (loops/loops.py:33-37)."""

    def __init__(self, backbone, projection_head, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.backbone = backbone
        self.projection_head = projection_head

    def compute_features(self, images) -> Tensor:
        return torch.concat(self.backbone(images), dim=-2)      # sample axis: dim 0, or dim 1 under [E, ...]

    def forward(self, query_images, inference=False):
        self.query_feature_list = self.compute_features(query_images)
        query_features = self.query_feature_list
        self._raise_error_if_features_are_multi_dimensional(query_features)
        if inference == True:  # noqa: E712
            return self.l2_distance_to_prototypes(query_features)
        return query_features

    def contrastive_forward(self, project_prototypes):
        projected_features = self.projection_head(self.query_feature_list)
        if project_prototypes == True:  # noqa: E712
            return projected_features, self.projection_head(self.prototypes)
        return projected_features, self.prototypes

    @staticmethod
    def is_transductive() -> bool:
        return False
