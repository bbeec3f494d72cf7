"""B200-native episodic prototypical-network head (drop-in for that path of
magcil/audio-few-shot-learning).

Import as ``afsl_b200`` (the directory name ``audio-few-shot-learning_b200`` is
not a Python identifier; ``afsl_b200/__init__.py`` loads this package under
that name).  Layout mirrors the reference's modules for the hot path:

    afsl_b200.models.{few_shot_classifier, prototypical, util_functions, main_modules}
    afsl_b200.loops.{loss, loops}
    afsl_b200.utils.augmentations
    afsl_b200.ops          torch.autograd bindings of the C ABI (include/afsl.h)
    afsl_b200.episodes     batched (E episodes at once) training / evaluation steps
    afsl_b200.parallel     episode sharding across GPUs (torch.distributed / NCCL)

All compute goes through libafsl.so (hand-written sm_100a CUDA, built in-tree by
``__graft_entry__.build()``); there is no CPU or eager fallback.
"""
__version__ = "0.1.0"
