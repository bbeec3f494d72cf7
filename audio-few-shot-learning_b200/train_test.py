"""Config-driven training + test driver (mirror of src/train_test.py:20-181), single GPU or one process per GPU.

    python -m afsl_b200.train_test -e experiment_config.json -m model_config.json [--runs 5] [--episodes-per-step E]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m afsl_b200.train_test -e ... -m ... --episodes-per-step 32

Reads the reference's two JSON files with the reference's keys (README.md:73-168,384-429): encoder / attention /
projection configuration, n_way / n_shot / n_query per split, loss selection (cpl | angular), SpecAugment parameters,
learning rate, MultiStepLR milestones, patience, tie strategy, single- or multi-segment test.  It then does what the
reference driver does per run: build the model, Adam + MultiStepLR, ``contrastive_training_loop`` (early stopping on the
validation accuracy, ``experiments/<experiment_folder>/model.pt``), reload the best checkpoint, test, print the message.

Datasets: the reference's ``MetaAudioDataset`` reads preprocessed .npy files (file IO, out of scope here).  Pass
``--dataset-factory module:callable`` to supply any object with that protocol (``callable(experiment_config, split)``; the
reference's own class works: ``datasets.datasets:MetaAudioDataset`` with /root/reference on PYTHONPATH and its extra
dependencies installed), or set ``"dataset_name": "synthetic"`` for the built-in synthetic episodes.

Extensions: ``--episodes-per-step E`` steps E episodes at once through ``EpisodeRunner`` (CUDA graph); under torchrun every
rank trains on its own episodes and the gradients are averaged with one NCCL all-reduce per step, the validation accuracy
of rank 0 drives early stopping on all ranks, and the test tasks are sharded with a final gather.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import random

import numpy as np
import torch
from torch.optim.lr_scheduler import MultiStepLR

from . import parallel
from .episodes import EpisodeRunner
from .loops.loops import contrastive_training_loop, evaluate_multisegment_loop, evaluate_single_segment
from .loops.loss import AngularLossClass, CPL_Loss, FSL_Loss
from .models.main_modules import EncoderModule, ProjectionHead, SelfAttention
from .models.prototypical import ContrastivePrototypicalNetworks, ContrastivePrototypicalNetworksWithoutAttention


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-e", "--experiment_config", help="Path to Experiment configuration file.", required=True)
    parser.add_argument("-m", "--model_config", help="Path to model_params file", required=True)
    parser.add_argument("--runs", type=int, default=5, help="repeated runs (the reference hard-codes 5)")
    parser.add_argument("--episodes-per-step", type=int, default=1, help="episodes per optimizer step per GPU (reference: 1)")
    parser.add_argument("--dataset-factory", default=None, help="module:callable(experiment_config, split) -> dataset")
    parser.add_argument("--seed", type=int, default=None, help="seed Python / NumPy / torch generators (reference: unseeded)")
    return parser.parse_args(argv)


def build_added_loss(experiment_config, device):
    """src/train_test.py:69-80: cpl wins over angular; neither -> None."""
    loss = experiment_config["loss"]
    if loss["cpl"]["use"] == True:                                            # noqa: E712
        return CPL_Loss(T=loss["cpl"]["t_param"], M=loss["cpl"]["m_param"]).to(device)
    if loss["angular"]["use"] == True:                                        # noqa: E712
        return AngularLossClass(angle=loss["angular"]["angle"],
                                prototypes_as_anchors=loss["angular"]["prototypes_as_anchors"]).to(device)
    return None


def build_model(experiment_config, model_config, device):
    """src/train_test.py:106-116."""
    backbone = EncoderModule(experiment_config=experiment_config, model_config=model_config)
    projection = ProjectionHead(model_config=model_config)
    if experiment_config["use_attention"] == True:                            # noqa: E712
        model = ContrastivePrototypicalNetworks(backbone=backbone, attention_model=SelfAttention(model_config=model_config),
                                                projection_head=projection)
    else:
        model = ContrastivePrototypicalNetworksWithoutAttention(backbone=backbone, projection_head=projection)
    return model.to(device)


def make_datasets(experiment_config, factory=None):
    if factory:
        module, name = factory.split(":")
        fn = getattr(importlib.import_module(module), name)
        return tuple(fn(experiment_config, split) for split in ("train", "valid", "test"))
    if experiment_config.get("dataset_name") == "synthetic":
        from .datasets.synthetic import SyntheticEpisodeDataset
        kw = dict(experiment_config.get("synthetic", {}))
        return tuple(SyntheticEpisodeDataset(experiment_config, split, **kw) for split in ("train", "valid", "test"))
    raise SystemExit("dataset classes / file IO are outside this build: pass --dataset-factory module:callable "
                     "(e.g. the reference's datasets.datasets:MetaAudioDataset) or use dataset_name 'synthetic'")


def mel_extractor(device):
    """src/train_test.py:123-129 (only needed for input_type 'wav')."""
    import torchaudio
    return torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_mels=128, n_fft=1024, hop_length=512, power=2.0).to(device)


def run_once(experiment_config, model_config, datasets, device, episodes_per_step=1, rank=0, world=1):
    """One of the reference's repeated runs (src/train_test.py:103-181) -> the test message dictionary."""
    cfg = experiment_config
    train_set, val_set, test_set = datasets
    model = build_model(cfg, model_config, device)
    fsl_loss = FSL_Loss().to(device)
    added_loss = build_added_loss(cfg, device)
    optimizer = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    scheduler = MultiStepLR(optimizer, milestones=cfg["scheduler_milestones"], gamma=cfg["scheduler_gamma"])
    feat_extractor = mel_extractor(device) if cfg.get("input_type", "spec") == "wav" else None
    runner, accuracy_sync = None, None
    dp = parallel.EpisodeDataParallel(model)
    if episodes_per_step > 1 or world > 1:
        runner = EpisodeRunner(model, cfg, optimizer, use_cuda_graph=episodes_per_step > 1)
        if world > 1:
            runner.grad_sync = dp.sync_gradients
    if world > 1:
        def accuracy_sync(acc):
            t = torch.tensor([float(acc)], dtype=torch.float64, device=device)
            torch.distributed.broadcast(t, src=0)
            return float(t.item())
    folder = cfg["experiment_folder"] if rank == 0 else os.path.join(cfg["experiment_folder"], f"rank{rank}")
    print("Starting to train")
    trained = contrastive_training_loop(
        model=model, train_dataset=train_set, validation_dataset=val_set, optimizer=optimizer,
        num_train_tasks=cfg["n_training_tasks"], num_val_tasks=cfg["n_training_tasks"], device=device, fsl_loss_fn=fsl_loss,
        cpl_loss_fn=added_loss, l_param=cfg["loss"]["l_param"], epochs=cfg["num_epochs"], train_scheduler=scheduler,
        patience=cfg["patience"], results_path=folder, project_prototypes=cfg["project_prototypes"],
        normalize_prototypes=cfg["normalize_prototypes"], n_train_classes=cfg["n_way_train"],
        n_validation_classes=cfg["n_way_validation"], k_support_train=cfg["n_shot_train"],
        k_support_validation=cfg["n_shot_validation"], k_query_train=cfg["n_query_train"],
        k_query_validation=cfg["n_query_validation"], feat_extractor=feat_extractor, use_contrastive=cfg["use_contrastive"],
        train_query_augmentations=cfg["train_query_augmentations"],
        validation_query_augmentations=cfg["validation_query_augmentations"],
        episodes_per_step=max(1, episodes_per_step) if runner is not None else 1, runner=runner, accuracy_sync=accuracy_sync)
    dp.sync_buffers()
    print("Starting to test")
    lo, hi = parallel.shard_range(cfg["n_testing_tasks"], rank, world)
    if cfg["multi_segm"] == False:                                            # noqa: E712
        mean, std, accs = evaluate_single_segment(model=trained, dataset=test_set, num_val_tasks=hi - lo, device=device,
                                                  n_classes=cfg["n_way_test"], k_support=cfg["n_shot_test"],
                                                  k_query=cfg["n_query_test"], feat_extractor=feat_extractor,
                                                  eval_query_augmentation=cfg["test_query_augmentations"], return_accuracies=True)
    else:
        trained.eval()
        msg = evaluate_multisegment_loop(test_dataset=test_set, n_classes=cfg["n_way_test"], k_support=cfg["n_shot_test"],
                                         k_query=cfg["n_query_test"], num_test_tasks=hi - lo, trained_model=trained,
                                         device=device, tie_strategy=cfg["tie_strategy"], feat_extractor=feat_extractor,
                                         eval_query_augmentation=cfg["test_query_augmentations"], return_accuracies=True)
        accs = msg.pop("accuracies")
    accs = parallel.gather_accuracies(np.asarray(accs, dtype=np.float64), cfg["n_testing_tasks"])
    if cfg["multi_segm"] == False:                                            # noqa: E712
        return (np.mean(accs), np.std(accs))                                  # the reference prints the (mean, std) tuple
    return {"mean_accuracy": np.mean(accs), "accuracy_std": np.std(accs)}


def main(argv=None):
    args = parse_args(argv)
    with open(args.experiment_config, "r") as f:
        experiment_config = json.load(f)
    with open(args.model_config, "r") as f:
        model_config = json.load(f)
    rank, world, local = parallel.init_from_env()
    if world > 1:
        device = f"cuda:{local}"
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))        # torchrun pins OMP_NUM_THREADS=1
    elif experiment_config["device"] == "cuda":
        device = f"cuda:{experiment_config['gpu_index']}"                     # src/train_test.py:40-45
    else:
        raise SystemExit("device 'cpu': the libafsl kernels have no CPU path (use the reference itself for CPU runs)")
    if args.seed is not None:
        random.seed(args.seed + rank); np.random.seed(args.seed + rank); torch.manual_seed(args.seed + rank)
    print(f"Loading Dataset:::  {experiment_config['dataset_name']}, Device used:::  {device}")
    datasets = make_datasets(experiment_config, args.dataset_factory)
    messages = []
    for i in range(args.runs):
        print(f"NEW RUN !!! NUMBER OF RUN ::: {i}")
        msg = run_once(experiment_config, model_config, datasets, device, args.episodes_per_step, rank, world)
        if rank == 0:
            print(msg)
        messages.append(msg)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return messages


if __name__ == "__main__":
    main()
