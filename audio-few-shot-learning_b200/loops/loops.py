"""Training / evaluation loops (host-side mirror of loops/loops.py:19-283).

Same function names, arguments and returned dictionaries as the reference, running on the libafsl kernels:
one optimizer step per episode (the reference's schedule), or - extension - ``episodes_per_step`` episodes per
step through ``EpisodeRunner`` (per-episode arithmetic unchanged, mean loss over the batch).
"""
from __future__ import annotations

import os
from statistics import mean

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from ..callbacks.early_stopping import EarlyStopping
from ..datasets.batch_creation import sample_episode, sample_episode_batch

PROJECT_PATH = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _repeat_labels_for_concat(model, support_list, support_labels, query_list, query_labels):
    """loops/loops.py:33-37: the no-attention model stacks the views along the sample axis."""
    if model.__class__.__name__ == "ContrastivePrototypicalNetworksWithoutAttention":
        support_labels = support_labels.repeat(len(support_list))
        query_labels = query_labels.repeat(len(query_list))
    return support_labels, query_labels


def training_epoch(model, dataset, optimizer, num_train_tasks, device, fsl_loss_fn, cpl_loss_fn, l_param, project_prototypes,
                   normalize_prototypes, n_classes, k_support, k_query, feat_extractor, use_contrastive, train_query_augmentations,
                   episodes_per_step: int = 1, runner=None):
    """``num_train_tasks`` episodes; returns {"loss", "fsl_loss", "cpl_loss"} means (loops/loops.py:19-63).

    ``episodes_per_step > 1`` (extension) needs ``runner`` = an ``EpisodeRunner`` built on ``model`` / ``optimizer``:
    episodes are drawn E at a time and stepped together."""
    all_loss, fsl_loss_list, cpl_loss_list = [], [], []
    model.train()
    if episodes_per_step > 1:
        if runner is None:
            raise ValueError("episodes_per_step > 1 needs runner=EpisodeRunner(model, experiment_config, optimizer)")
        done = 0
        while done < num_train_tasks:
            e = min(episodes_per_step, num_train_tasks - done)
            out = runner.train_step(sample_episode_batch(dataset, e, n_classes, k_support, k_query))
            all_loss += out["loss"].tolist()
            fsl_loss_list += out["fsl_loss"].tolist()
            cpl_loss_list += out["cpl_loss"].tolist() if "cpl_loss" in out else [np.nan] * e
            done += e
        return {"loss": mean(all_loss), "fsl_loss": mean(fsl_loss_list), "cpl_loss": mean(cpl_loss_list)}
    for _ in range(num_train_tasks):
        support_list, support_labels, query_list, query_labels, _ = sample_episode(
            dataset=dataset, n_classes=n_classes, k_support=k_support, k_query=k_query, is_test=False, device=device,
            feat_extractor=feat_extractor, augment_query=train_query_augmentations)
        support_labels, query_labels = _repeat_labels_for_concat(model, support_list, support_labels, query_list, query_labels)
        optimizer.zero_grad()
        model.process_support_set(support_list, support_labels.to(device))
        query_features = model(query_list)
        fsl_loss = fsl_loss_fn(model.prototypes, query_features, query_labels.to(device))
        if use_contrastive == True:                                          # noqa: E712
            cpl_query_features, prototypes = model.contrastive_forward(project_prototypes)
            if project_prototypes == True:                                   # noqa: E712
                normalize_prototypes = False
            if normalize_prototypes == True:                                 # noqa: E712
                prototypes = ops.l2_normalize(prototypes, eps=1e-12)
            cpl_loss = cpl_loss_fn(prototypes, cpl_query_features, query_labels.to(device))
            final_loss = fsl_loss + l_param * cpl_loss
            cpl_loss_list.append(cpl_loss.item())
        else:
            final_loss = fsl_loss
            cpl_loss_list.append(np.nan)
        all_loss.append(final_loss.item())
        fsl_loss_list.append(fsl_loss.item())
        final_loss.backward()
        optimizer.step()
    return {"loss": mean(all_loss), "fsl_loss": mean(fsl_loss_list), "cpl_loss": mean(cpl_loss_list)}


def evaluate_on_one_task(model, support_images, support_labels, query_images, query_labels):
    """(#correct query predictions, #queries) of one task (loops/loops.py:66-81)."""
    model.process_support_set(support_images, support_labels)
    with torch.no_grad():
        predictions = model(query_images, inference=True)
    correct = (torch.max(predictions, 1)[1] == query_labels).sum().item()
    return correct, len(query_labels)


def evaluate_single_segment(model, dataset, num_val_tasks, device, n_classes, k_support, k_query, feat_extractor,
                            eval_query_augmentation, return_accuracies: bool = False):
    """(mean, std) of the per-task accuracies over ``num_val_tasks`` sampled tasks (loops/loops.py:84-121);
    ``return_accuracies`` (extension) appends the per-task list - what a rank contributes to a sharded evaluation."""
    accuracies = []
    model.eval()
    with torch.no_grad():
        for _ in range(num_val_tasks):
            support_list, support_labels, query_list, query_labels, _ = sample_episode(
                dataset=dataset, n_classes=n_classes, k_support=k_support, k_query=k_query, is_test=False, device=device,
                feat_extractor=feat_extractor, augment_query=eval_query_augmentation)
            support_labels, query_labels = _repeat_labels_for_concat(model, support_list, support_labels, query_list, query_labels)
            correct, total = evaluate_on_one_task(model, [t.to(device) for t in support_list], support_labels.to(device),
                                                  [t.to(device) for t in query_list], query_labels.to(device))
            accuracies.append(correct / total)
    if return_accuracies:
        return np.mean(accuracies), np.std(accuracies), accuracies
    return np.mean(accuracies), np.std(accuracies)


def contrastive_training_loop(model, train_dataset, validation_dataset, optimizer, num_train_tasks, num_val_tasks, device, fsl_loss_fn,
                              cpl_loss_fn, l_param, epochs, train_scheduler, patience, results_path, project_prototypes,
                              normalize_prototypes, n_train_classes, n_validation_classes, k_support_train, k_support_validation,
                              k_query_train, k_query_validation, feat_extractor, use_contrastive, train_query_augmentations,
                              validation_query_augmentations, episodes_per_step: int = 1, runner=None, accuracy_sync=None):
    """Epoch loop with early stopping on the validation accuracy; reloads and returns the best model
    (loops/loops.py:124-167).  The checkpoint lives at experiments/<results_path>/model.pt like the reference's.
    ``accuracy_sync`` (extension, multi-GPU): maps this rank's validation accuracy to the one every rank must act on
    (rank 0's), so that all ranks stop at the same epoch."""
    folder = os.path.join(PROJECT_PATH, "experiments", results_path)
    os.makedirs(folder, exist_ok=True)
    checkpoint = os.path.join(folder, "model.pt")
    stopper = EarlyStopping(path=checkpoint, patience=patience, verbose=True)
    for epoch in range(1, epochs + 1):
        print(f"Epoch: {epoch:03}/{epochs + 1:03}")
        loss_msg = training_epoch(model=model, dataset=train_dataset, optimizer=optimizer, num_train_tasks=num_train_tasks,
                                  device=device, fsl_loss_fn=fsl_loss_fn, cpl_loss_fn=cpl_loss_fn, l_param=l_param,
                                  project_prototypes=project_prototypes, normalize_prototypes=normalize_prototypes,
                                  n_classes=n_train_classes, k_support=k_support_train, k_query=k_query_train,
                                  feat_extractor=feat_extractor, use_contrastive=use_contrastive,
                                  train_query_augmentations=train_query_augmentations, episodes_per_step=episodes_per_step,
                                  runner=runner)
        print(loss_msg)
        accuracy, _ = evaluate_single_segment(model=model, dataset=validation_dataset, num_val_tasks=num_val_tasks, device=device,
                                              n_classes=n_validation_classes, k_support=k_support_validation,
                                              k_query=k_query_validation, feat_extractor=feat_extractor,
                                              eval_query_augmentation=validation_query_augmentations)
        if accuracy_sync is not None:
            accuracy = accuracy_sync(accuracy)
        stopper(val_accuracy=accuracy, model=model, epoch=epoch)
        if stopper.early_stop:
            print("Early Stopping.")
            break
        train_scheduler.step()
    model.load_state_dict(torch.load(checkpoint))
    return model


def calculate_majority_vote_accuracy(predicted_labels, spectrogram_ids, query_labels, posterior_values, tie_strategy="min_label"):
    """Clip-level accuracy of one multi-segment task by majority vote over each clip's segments
    (loops/loops.py:169-247): ties -> "min_label": smallest tied label; "max_posterior": label of the first segment
    with the strictly greatest posterior among the tied labels; anything else: first-encountered tied label."""
    as_dev = lambda a: torch.as_tensor(np.asarray(a.detach().cpu()) if torch.is_tensor(a) and not a.is_cuda else a)
    pred = predicted_labels if torch.is_tensor(predicted_labels) else torch.as_tensor(predicted_labels)
    device = pred.device if pred.is_cuda else torch.device("cuda", torch.cuda.current_device())
    to = lambda a: (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))).to(device)
    n = int(pred.numel())
    offsets = torch.tensor([0, n], dtype=torch.int64)
    correct, clips = ops.eval_vote(to(pred), to(spectrogram_ids), to(query_labels), to(posterior_values).float(), offsets,
                                   tie_strategy)
    return float(correct.item()) / float(clips.item())


def evaluate_multisegment_loop(test_dataset, n_classes, k_support, k_query, num_test_tasks, trained_model, device, tie_strategy,
                               feat_extractor, eval_query_augmentation, return_accuracies: bool = False):
    """{"mean_accuracy", "accuracy_std"} over multi-segment test tasks (loops/loops.py:250-283); ``return_accuracies``
    (extension) adds the per-task list under "accuracies"."""
    accuracies = []
    for _ in range(num_test_tasks):
        support_list, support_labels, query_list, query_labels, audio_ids = sample_episode(
            dataset=test_dataset, n_classes=n_classes, k_support=k_support, k_query=k_query, is_test=True, device=device,
            feat_extractor=feat_extractor, augment_query=eval_query_augmentation)
        support_labels, query_labels = _repeat_labels_for_concat(trained_model, support_list, support_labels, query_list,
                                                                 query_labels)
        support_set = [t.to(device) for t in support_list]
        query_set = [t.to(device) for t in query_list]
        support_labels, query_labels, audio_ids = support_labels.to(device), query_labels.to(device), audio_ids.to(device)
        trained_model.process_support_set(support_set, support_labels)
        with torch.no_grad():
            predictions = trained_model(query_set, inference=True)
            posterior_values, predicted_labels = torch.max(predictions, 1)
            accuracies.append(calculate_majority_vote_accuracy(predicted_labels=predicted_labels, spectrogram_ids=audio_ids,
                                                               query_labels=query_labels, tie_strategy=tie_strategy,
                                                               posterior_values=posterior_values))
    msg = {"mean_accuracy": np.mean(accuracies), "accuracy_std": np.std(accuracies)}
    if return_accuracies:
        msg["accuracies"] = accuracies
    return msg
