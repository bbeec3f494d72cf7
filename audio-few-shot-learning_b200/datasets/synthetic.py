"""Synthetic episode dataset with the reference's dataset protocol (what ``sample_episode`` reads from a
``MetaAudioDataset``, datasets/datasets.py:19-60 and datasets/batch_creation.py:22-23,38,50,112): ``class_to_label``,
``data_df`` (``label`` / ``index_column``), ``multi_segm``, ``input_type``, ``specaug_use``, ``waveaug_use``,
``experiment_config`` and ``__getitem__ -> (spectrogram [S,1,128,T], label)``.

The reference's dataset classes read .npy files from /data/<name> (file IO: out of scope); this one makes
log-mel-shaped N(0,1) clips with a smooth rank-one pattern per class on top, so that few-shot tasks on it are learnable.
Used by ``train_test.py`` when ``dataset_name == "synthetic"`` and by the tests.
"""
from __future__ import annotations

import pandas as pd
import torch


def class_patterns(classes: int, t_len: int, seed: int, scale: float = 0.4, mels: int = 128) -> torch.Tensor:
    """[classes, 1, mels, t_len]: one smooth rank-one time-frequency pattern per class (deterministic in ``seed``)."""
    g = torch.Generator().manual_seed(10_000 + seed)
    f = torch.nn.functional.avg_pool1d(torch.randn(classes, 1, mels + 8, generator=g), 9, 1)[:, 0]
    t = torch.nn.functional.avg_pool1d(torch.randn(classes, 1, t_len + 8, generator=g), 9, 1)[:, 0]
    return (scale * 9.0 * f.unsqueeze(2) * t.unsqueeze(1)).unsqueeze(1)


class SyntheticEpisodeDataset:
    def __init__(self, experiment_config: dict, split: str = "train", classes: int = 12, per_class: int = 16, t_len: int = 157,
                 max_segments: int = 1, seed: int = 0, scale: float = 0.4):
        self.experiment_config = experiment_config
        self.split = split
        salt = {"train": 0, "valid": 1, "test": 2}.get(split, 3)
        g = torch.Generator().manual_seed(1000 * seed + salt)
        n = classes * per_class
        pat = class_patterns(classes, t_len, 1000 * seed + salt, scale)
        self.multi_segm = bool(experiment_config.get("multi_segm", max_segments > 1))
        segs = torch.randint(1, max_segments + 1, (n,), generator=g).tolist() if self.multi_segm else [1] * n
        self.clips = [torch.randn(s, 1, 128, t_len, generator=g) + pat[i // per_class] for i, s in enumerate(segs)]
        names = [f"{split}_class{i}" for i in range(classes)]
        self.class_to_label = {nm: i for i, nm in enumerate(names)}
        self.data_df = pd.DataFrame({"label": [names[i // per_class] for i in range(n)], "index_column": list(range(n))})
        self.input_type = "spec"
        self.specaug_use = bool(experiment_config.get("specaug_params", {}).get("use", False))
        self.waveaug_use = False

    def __len__(self):
        return len(self.clips)

    def __getitem__(self, item):
        return self.clips[item], self.data_df["label"].iloc[item]
