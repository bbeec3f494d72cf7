"""Base class of the few-shot models (mirror of models/few_shot_classifier.py:13-148).

Same constructor, attributes (``prototypes``, ``support_features``, ``support_labels``),
methods and errors as the reference; prototypes and distances run on libafsl kernels and every
method also accepts a leading episode dimension.
"""
from abc import abstractmethod
from typing import Optional

import torch
from torch import Tensor, nn

from .. import ops
from .util_functions import compute_prototypes


class FewShotClassifier(nn.Module):
    def __init__(self, backbone: Optional[nn.Module] = None, use_softmax: bool = False,
                 feature_centering: Optional[Tensor] = None, feature_normalization: Optional[float] = None):
        super().__init__()
        self.backbone = backbone if backbone is not None else nn.Identity()
        self.use_softmax = use_softmax
        self.prototypes = torch.tensor(())
        self.support_features = torch.tensor(())
        self.support_labels = torch.tensor(())
        self.feature_centering = feature_centering if feature_centering is not None else torch.tensor(0)
        self.feature_normalization = feature_normalization

    @abstractmethod
    def forward(self, query_images: Tensor) -> Tensor:
        raise NotImplementedError("All few-shot algorithms must implement a forward method.")

    def process_support_set(self, support_images: Tensor, support_labels: Tensor):
        self.compute_prototypes_and_store_support_set(support_images, support_labels)

    @staticmethod
    def is_transductive() -> bool:
        raise NotImplementedError("All few-shot algorithms must implement a is_transductive method.")

    def compute_features(self, images: Tensor) -> Tensor:
        features = self.backbone(images) - self.feature_centering
        if self.feature_normalization is not None:
            return nn.functional.normalize(features, p=self.feature_normalization, dim=-1)
        return features

    def softmax_if_specified(self, output: Tensor, temperature: float = 1.0) -> Tensor:
        return (temperature * output).softmax(-1) if self.use_softmax else output

    def l2_distance_to_prototypes(self, samples: Tensor) -> Tensor:
        """Negated Euclidean distance to the stored prototypes (few_shot_classifier.py:108-116)."""
        return ops.l2_scores(samples, self.prototypes)

    def cosine_distance_to_prototypes(self, samples) -> Tensor:
        """Cosine logits (few_shot_classifier.py:118-126); never called by the reference drivers."""
        return ops.l2_normalize(samples) @ ops.l2_normalize(self.prototypes).transpose(-1, -2)

    def compute_prototypes_and_store_support_set(self, support_images: Tensor, support_labels: Tensor):
        self.support_labels = support_labels
        self.support_features = self.compute_features(support_images)
        self._raise_error_if_features_are_multi_dimensional(self.support_features)
        self.prototypes = compute_prototypes(self.support_features, support_labels)

    @staticmethod
    def _raise_error_if_features_are_multi_dimensional(features: Tensor):
        # [N, D], or [E, N, D] when episodes are batched
        if len(features.shape) not in (2, 3):
            raise ValueError("Illegal backbone or feature shape. "
                             "Expected output for an image is a 1-dim tensor.")
