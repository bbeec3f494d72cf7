"""compute_prototypes on the GPU kernel (mirror of models/util_functions.py:6-19).

The reference file's other helpers (entropy, k_nearest_neighbours, power_transform,
strip_prefix) are never called on the hot path and are out of scope.
"""
from typing import Optional

from torch import Tensor

from .. import ops


def compute_prototypes(support_features: Tensor, support_labels: Tensor, n_way: Optional[int] = None) -> Tensor:
    """Mean feature vector per label.

    ``[Ns, D], [Ns] -> [W, D]`` exactly as the reference; with a leading episode dimension
    ``[E, Ns, D], [E, Ns] -> [E, W, D]``.  ``W = len(unique(labels))`` (labels must be 0..W-1) unless
    ``n_way`` is given, which avoids the device->host synchronisation of ``unique``.
    """
    return ops.prototypes(support_features, support_labels, n_way)
