"""Episode sampling (host-side mirror of datasets/batch_creation.py:21-170) and its batched form.

``sample_episode`` keeps the reference's signature, return value and - draw for draw - its use of Python's
``random`` (class sample, per-class shuffle, segment pick) and of the SpecAugment generators, so that with the same
seeds it selects the same clips and draws the same masks as the reference.  The dataset protocol is the reference's:
``class_to_label``, ``data_df`` (columns ``label`` / ``index_column``), ``multi_segm``, ``input_type``,
``specaug_use``, ``experiment_config`` and ``__getitem__ -> (spectrogram [S,1,128,T], label)``.
SpecAugment runs on the GPU kernel, so the views are produced after the move to ``device``.

``sample_episode_batch`` draws E episodes (same per-episode draw order) and returns an ``EpisodeBatch`` for
``EpisodeRunner``: index selection is one pass over a per-class index table built once per dataset instead of
one pandas filter per class per episode.

``input_type == 'wav'`` follows the reference too: 5-second splits of the waveform (``variable_wav_splits``), then the log-mel
front end with the caller's ``feat_extractor`` configuration (torchaudio ``MelSpectrogram``): on the GPU ONE libafsl launch
(``afsl_logmel_f32``: STFT, power, the transform's own mel filters, ``10*log10(x + eps)``, the dataset's global normalisation).  Waveform augmentation (``waveaug_use``) needs torch_audiomentations, which is not
part of this build, and raises.
"""
from __future__ import annotations

import random
from typing import Dict, List, Tuple

import numpy as np
import torch

from .. import ops
from ..episodes import EpisodeBatch
from ..utils.augmentations import SpecAugment


def augment_spectrogram(item, experiment_config):
    """Four SpecAugment views of a set (datasets/batch_creation.py:10-13)."""
    return SpecAugment(experiment_config).apply_augmentations(item)


def _class_index_table(dataset) -> Dict[int, List[int]]:
    """label id -> dataset indices in ``data_df`` order, cached on the dataset (replaces the per-episode pandas
    filter of batch_creation.py:37-38; the lists are copied before they are shuffled)."""
    table = getattr(dataset, "_afsl_class_index_table", None)
    if table is None:
        name_to_label = dataset.class_to_label
        table = {label: [] for label in name_to_label.values()}
        for name, idx in zip(dataset.data_df["label"].tolist(), dataset.data_df["index_column"].tolist()):
            table[name_to_label[name]].append(idx)
        try:
            dataset._afsl_class_index_table = table
        except AttributeError:
            pass
    return table


def _episode_classes(dataset, n_classes: int, k_support: int, k_query: int):
    """The reference's draws of one episode, class by class: the sorted class sample up front, then - lazily, so that
    the caller's segment picks for class c happen before class c+1's shuffle exactly as in the reference
    (batch_creation.py:25,36-48) - one shuffle of each sampled class's index list.
    Yields (new label, support indices, query indices)."""
    class_labels = list(dataset.class_to_label.values())
    sampled = sorted(random.sample(class_labels, n_classes))
    table = _class_index_table(dataset)
    label_to_name = {v: k for k, v in dataset.class_to_label.items()}
    for new_label, label in enumerate(sampled):
        indices = list(table[label])
        random.shuffle(indices)
        if len(indices) < k_support + k_query:
            raise ValueError(f"Not enough samples for class {label_to_name[label]}. "
                             f"Available: {len(indices)}, required: {k_support + k_query}")
        yield new_label, indices[:k_support], indices[k_support:k_support + k_query]


def variable_wav_splits(sample):
    """5-second (80 000-sample) pieces of a 16 kHz waveform (datasets/batch_creation.py:172-213): shorter clips are
    tiled up to one piece; longer ones are cut, and - as in the reference, which tiles the WHOLE clip there - the
    remainder piece is the first 5 seconds of the clip repeated."""
    expected_size = 5 * 16000
    raw_splits = []
    if sample.shape[0] < expected_size:
        multiply_up = int(np.ceil(expected_size / sample.shape[0]))
        raw_splits.append(sample.repeat((multiply_up,))[:expected_size])
    else:
        start = 0
        while start < sample.shape[0]:
            to_end = sample.shape[0] - start
            if to_end >= expected_size:
                raw_splits.append(sample[start:start + expected_size])
                start += expected_size
            else:
                multiply_up = int(np.ceil(expected_size / to_end))
                raw_splits.append(sample.repeat((multiply_up,))[:expected_size])
                start = sample.shape[0]
    return raw_splits


def mel_spec_function_gpu(x, mel_transform):
    """Log-mel spectrogram in dB of a waveform batch (datasets/batch_creation.py:215-218).  On the GPU with the
    reference's MelSpectrogram configuration this is the one-launch libafsl front end (STFT, power, mel filters, dB);
    on the CPU, or with another transform, the torchaudio module itself is applied."""
    if x.is_cuda and ops.log_mel_supported(mel_transform):
        return ops.log_mel(x, mel_transform)[:, 0]
    mel_spec = mel_transform(x)
    return 20.0 / 2 * torch.log10(mel_spec + torch.finfo(mel_spec.dtype).eps)


def log_mel_normalized(x, mel_transform, mean, std):
    """``((mel_spec_function_gpu(x) - mean) / std).unsqueeze(1)`` (datasets/batch_creation.py:138-143): one libafsl launch
    on the GPU (normalisation fused into the front-end kernel), the eager chain elsewhere."""
    if x.is_cuda and ops.log_mel_supported(mel_transform):
        return ops.log_mel(x, mel_transform, float(mean), float(std))
    return ((mel_spec_function_gpu(x, mel_transform=mel_transform) - mean) / std).unsqueeze(1)


def _sample_wav_episode(dataset, n_classes, k_support, k_query, is_test, device, feat_extractor, augment_query):
    """``input_type == 'wav'`` branch of sample_episode (datasets/batch_creation.py:77-101,138-143)."""
    if dataset.waveaug_use == True:                                         # noqa: E712
        raise NotImplementedError("waveform augmentation needs torch_audiomentations, which is outside this build")
    multi_segm = dataset.multi_segm
    support_set, support_labels, query_set, query_labels, audio_ids = [], [], [], [], []
    query_counter = 0
    for new_label, s_idx, q_idx in _episode_classes(dataset, n_classes, k_support, k_query):
        for idx in s_idx:
            wav, _ = dataset[idx]
            if multi_segm == True:                                          # noqa: E712
                pieces = variable_wav_splits(wav)
                picked = torch.from_numpy(pieces[random.randint(0, len(pieces) - 1)]).reshape(1, -1)
            else:
                picked = torch.from_numpy(wav).reshape(1, -1)
            support_set.append(picked)
            support_labels.append(new_label)
        for idx in q_idx:
            wav, _ = dataset[idx]
            if multi_segm == True:                                          # noqa: E712
                pieces = variable_wav_splits(wav)
                if is_test == False:                                        # noqa: E712
                    picked = torch.from_numpy(pieces[random.randint(0, len(pieces) - 1)].reshape(1, -1))
                else:
                    picked = torch.cat([torch.from_numpy(piece.reshape(1, -1)) for piece in pieces], dim=0)
            else:
                picked = torch.from_numpy(wav.reshape(1, -1))
            query_set.append(picked)
            query_labels.extend([new_label] * picked.shape[0])
            audio_ids.extend([query_counter] * picked.shape[0])
            query_counter += 1
    mean, std = dataset.get_normalization_stats()
    stacked = torch.cat(support_set + query_set, dim=0).to(device)
    spectrograms = log_mel_normalized(stacked, feat_extractor, mean, std)
    n_support = n_classes * k_support
    return ([spectrograms[:n_support]], torch.tensor(support_labels), [spectrograms[n_support:]], torch.tensor(query_labels),
            torch.tensor(audio_ids))


def sample_episode(dataset, n_classes, k_support, k_query, is_test, device, feat_extractor, augment_query):
    """One episode: (support view list, support labels, query view list, query labels, audio ids)."""
    if dataset.input_type == "wav":
        return _sample_wav_episode(dataset, n_classes, k_support, k_query, is_test, device, feat_extractor, augment_query)
    support_set, support_labels, query_set, query_labels, audio_ids = [], [], [], [], []
    query_counter = 0
    for new_label, s_idx, q_idx in _episode_classes(dataset, n_classes, k_support, k_query):
        for idx in s_idx:
            spectrogram, _ = dataset[idx]
            if spectrogram.shape[0] != 1:                                  # multi-segment clip: one random segment
                spectrogram = spectrogram[random.randint(0, spectrogram.shape[0] - 1)].unsqueeze(0)
            support_set.append(spectrogram)
            support_labels.append(new_label)
        for idx in q_idx:
            spectrogram, _ = dataset[idx]
            if is_test == False and spectrogram.shape[0] != 1:             # noqa: E712 - test keeps every segment
                spectrogram = spectrogram[random.randint(0, spectrogram.shape[0] - 1)].unsqueeze(0)
            query_set.append(spectrogram)
            query_labels.extend([new_label] * spectrogram.shape[0])
            audio_ids.extend([query_counter] * spectrogram.shape[0])
            query_counter += 1
    support_set = torch.cat(support_set, dim=0).to(device)
    query_set = torch.cat(query_set, dim=0).to(device)
    if dataset.specaug_use == True:                                         # noqa: E712
        support_list = augment_spectrogram(support_set, dataset.experiment_config)
        query_list = augment_spectrogram(query_set, dataset.experiment_config) if augment_query == True else [query_set]  # noqa: E712
    else:
        support_list, query_list = [support_set], [query_set]
    return (support_list, torch.tensor(support_labels), query_list, torch.tensor(query_labels), torch.tensor(audio_ids))


def sample_episode_batch(dataset, episodes: int, n_classes: int, k_support: int, k_query: int,
                         pin_memory: bool = False) -> EpisodeBatch:
    """E single-segment training episodes as one host ``EpisodeBatch`` ([E,Ns,1,F,T] / [E,Nq,1,F,T], labels
    0..W-1 block-sorted as the reference builds them).  Draw order per episode = ``sample_episode``'s; the
    SpecAugment views are NOT taken here - ``EpisodeRunner`` draws and applies them on the device."""
    if dataset.input_type != "spec":
        raise NotImplementedError("input_type 'wav' (GPU MelSpectrogram front end) is outside this build's scope")
    def fetch(idxs):
        rows = []
        for idx in idxs:
            spectrogram, _ = dataset[idx]
            if spectrogram.shape[0] != 1:                                  # multi-segment clip: one random segment
                spectrogram = spectrogram[random.randint(0, spectrogram.shape[0] - 1)].unsqueeze(0)
            rows.append(spectrogram)
        return rows

    sup, qry = [], []
    for _ in range(episodes):
        s_rows, q_rows = [], []
        for _, s_idx, q_idx in _episode_classes(dataset, n_classes, k_support, k_query):
            s_rows += fetch(s_idx)                                         # support picks before query picks, per class
            q_rows += fetch(q_idx)
        sup.append(torch.cat(s_rows))
        qry.append(torch.cat(q_rows))
    support, query = torch.stack(sup), torch.stack(qry)
    sl = torch.arange(n_classes).repeat_interleave(k_support).expand(episodes, -1).contiguous()
    ql = torch.arange(n_classes).repeat_interleave(k_query).expand(episodes, -1).contiguous()
    batch = EpisodeBatch(support, sl, query, ql, n_classes)
    return batch.pin() if pin_memory else batch
