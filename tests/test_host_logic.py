"""CPU tests of the host-side logic: RNG replay, mask packing, grouped BatchNorm, state-dict
compatibility, episode sharding and the world-size-2 gloo path."""
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_names, load_golden, t
from oracle import head as ohead
from oracle import modules as omod
from oracle import specaug as ospec

CFG = {"specaug_params": {"use": True, "mask_param": 16, "W": 22, "num_mask": 2, "mask_value": 0, "p": 0.282}}


def test_specaug_draws_replay_reference_order():
    from afsl_b200.utils.augmentations import SpecAugment
    torch.manual_seed(5); np.random.seed(5)
    want = [ospec.draw_params(25, 157, CFG) for _ in range(3)]
    torch.manual_seed(5); np.random.seed(5)
    got = SpecAugment(CFG).draw_batch(3, 25, 157, replay_reference_rng=True)
    assert torch.equal(got.warp_p, torch.cat([w.warp_p for w in want]))
    assert torch.equal(got.warp_d, torch.cat([w.warp_d for w in want]))
    assert got.time_masks.tolist() == [[list(m) for m in w.time_masks] for w in want]
    assert got.freq_masks.tolist() == [[list(m) for m in w.freq_masks] for w in want]


def test_specaug_vectorised_draws_in_range():
    from afsl_b200.utils.augmentations import SpecAugment
    p = SpecAugment(CFG).draw_batch(2000, 25, 157, replay_reference_rng=False)
    assert p.warp_p.min() >= 22 and p.warp_p.max() < 157 - 22 and p.warp_d.min() >= -22 and p.warp_d.max() < 22
    t0, tl = p.time_masks[..., 0], p.time_masks[..., 1]
    assert tl.min() >= 1 and tl.max() <= min(16, int(0.282 * 157)) and t0.min() >= 0 and bool((t0 + tl <= 156).all())
    f0, fl = p.freq_masks[..., 0], p.freq_masks[..., 1]
    assert fl.min() >= 1 and fl.max() <= 16 and f0.min() >= 0 and bool((f0 + fl <= 127).all())
    assert set(tl.unique().tolist()) == set(range(1, 17))          # every length is reachable


@pytest.mark.parametrize("name", golden_names("specaug_"))
def test_host_spline_matches_reference(name):
    from afsl_b200.utils.augmentations import warp_source_x
    g = load_golden(name)
    assert torch.equal(warp_source_x(t(g["warp_p"]), t(g["warp_d"]), g["x"].shape[-1]), t(g["src_x"]))


@pytest.mark.parametrize("name", golden_names("cpl_"))
def test_cpl_keep_replay_matches_reference(name):
    from afsl_b200.loops.loss import draw_keep_reference
    from afsl_b200.ops import pack_keep
    g = load_golden(name)
    torch.manual_seed(int(g["seed"]))
    keep = draw_keep_reference(t(g["labels"]), int(g["m"]))
    assert torch.equal(keep, t(g["keep"]))
    packed = pack_keep(keep)
    n = keep.shape[0]
    for i in (0, n // 2, n - 1):
        for j in range(n):
            assert bool((int(packed[i, j // 32]) >> (j % 32)) & 1) == bool(keep[i, j])


def test_cpl_keep_vectorised_distribution():
    from afsl_b200.loops.loss import draw_keep_vectorised
    labels = torch.arange(5).repeat_interleave(8)[torch.randperm(40)].expand(6, -1)
    keep = draw_keep_vectorised(labels, 3, 5)
    same = labels.unsqueeze(2) == labels.unsqueeze(1)
    eye = torch.eye(40, dtype=torch.bool)
    assert bool((keep & same == eye).all())                        # own class: only the query itself
    for c in range(5):                                             # exactly M from every other class
        cols = (labels == c).unsqueeze(1)
        cnt = (keep & cols).sum(2)
        rows_other = labels != c
        assert bool((cnt[rows_other] == 3).all())


def test_grouped_batchnorm_equals_sequential_calls():
    from afsl_b200.models.main_modules import GroupedBatchNorm1d, GroupedBatchNorm2d
    torch.manual_seed(0)
    for grouped_cls, ref_cls, shape in ((GroupedBatchNorm2d, torch.nn.BatchNorm2d, (6, 8, 5, 7)),
                                        (GroupedBatchNorm1d, torch.nn.BatchNorm1d, (6, 8))):
        g, r = grouped_cls(8), ref_cls(8)
        with torch.no_grad():
            g.weight.uniform_(0.5, 1.5); g.bias.uniform_(-1, 1)
            r.weight.copy_(g.weight); r.bias.copy_(g.bias)
        x = torch.randn(4 * shape[0], *shape[1:], requires_grad=True)
        xr = x.detach().clone().requires_grad_(True)
        g.group_size = shape[0]
        yg = g(x)
        yr = torch.cat([r(xr[i * shape[0]:(i + 1) * shape[0]]) for i in range(4)])
        torch.testing.assert_close(yg, yr, rtol=1e-5, atol=1e-6)
        w = torch.randn_like(yg)
        (yg * w).sum().backward(); (yr * w).sum().backward()
        torch.testing.assert_close(x.grad, xr.grad, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(g.weight.grad, r.weight.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(g.running_mean, r.running_mean, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(g.running_var, r.running_var, rtol=1e-5, atol=1e-7)
        assert int(g.num_batches_tracked) == int(r.num_batches_tracked) == 4


def test_batched_encoder_equals_per_episode_calls():
    """[E,N,1,F,T] views through EncoderModule == the reference's one call per (episode, view)."""
    from afsl_b200.models.main_modules import EncoderModule, StandardCNN
    torch.manual_seed(1)
    enc = EncoderModule({"encoder_name": "CNN"}, {}, encoder=StandardCNN(1, (1, 1, 32, 40), 64, [2, 2], 16))
    for m in enc.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    enc.train()
    views = [torch.randn(3, 4, 1, 32, 40) for _ in range(2)]
    got = enc(views)
    import copy
    ref = copy.deepcopy(enc)
    for v in range(2):
        for e in range(3):
            want = ref([views[v][e]])[0]
            torch.testing.assert_close(got[v][e], want, rtol=1e-4, atol=1e-5)


def test_stack_views_reuses_a_shared_buffer():
    """Views that are consecutive slices of one buffer (what the SpecAugment kernel writes) become the encoder batch
    without a copy; anything else (separate tensors, gaps, a different order) is concatenated - same values either way."""
    from afsl_b200.models.main_modules import _stack_views
    e, n, f, tl = 3, 4, 8, 10
    buf = torch.randn(6, e * n, 1, f, tl)
    views = [buf[v].view(e, n, 1, f, tl) for v in range(6)]
    want = torch.cat([v.reshape(e * n, 1, f, tl) for v in views], 0)
    got = _stack_views(views, e, n)
    assert got.data_ptr() == buf.data_ptr() and torch.equal(got, want)                    # in place
    tail = _stack_views(views[2:5], e, n)                                                 # a run that starts inside the buffer
    assert tail.data_ptr() == buf[2].data_ptr() and torch.equal(tail, want[2 * e * n:5 * e * n])
    for other in ([views[1], views[0]], [views[0], views[2]], [v.clone() for v in views[:2]]):
        out = _stack_views(other, e, n)
        assert out.data_ptr() != other[0].data_ptr()                                      # copied
        assert torch.equal(out, torch.cat([v.reshape(e * n, 1, f, tl) for v in other], 0))


def test_state_dict_keys_match_reference():
    from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, SelfAttention
    from afsl_b200.models.prototypical import (ContrastivePrototypicalNetworks,
                                               ContrastivePrototypicalNetworksWithoutAttention)
    import bench
    g = load_golden("modules_model_concat")
    att, proj = SelfAttention(bench.MODEL_CONFIG), ProjectionHead({"Projection": {"input_dim": 256, "hidden_dim": 128, "output_dim": 256}})
    ident = torch.nn.Identity()
    fused = ContrastivePrototypicalNetworks(ident, att, proj)
    concat = ContrastivePrototypicalNetworksWithoutAttention(ident, proj)
    assert list(fused.state_dict().keys()) == [str(k) for k in g["keys_fused"]]
    assert list(concat.state_dict().keys()) == [str(k) for k in g["keys_concat"]]
    for name in golden_names("modules_encoder_"):
        ge = load_golden(name)
        kind = name.split("_")[2]
        mc = dict(bench.MODEL_CONFIG)
        mc["CNN"] = {"in_channels": 1, "hidden_channels": 64, "pool_dim": [3, 3], "out_dim": 64,
                     "trial_shape": (1, 1, 128, int(name.split("_t")[-1]))}
        enc = EncoderModule({"encoder_name": kind}, mc)
        assert list(enc.encoder.state_dict().keys()) == [str(k) for k in ge["keys"]]
        enc.encoder.load_state_dict({k[2:]: t(v) for k, v in ge.items() if k.startswith("w_")})
        enc.eval()
        with torch.no_grad():                      # cuDNN-free CPU check of the mirrored architecture
            torch.testing.assert_close(enc([t(ge["x"])])[0], t(ge["y_eval"]), rtol=1e-5, atol=1e-6)
    with pytest.raises(TypeError):                 # the reference's 'CNN' branch fails the same way
        from afsl_b200.models.main_modules import get_backbone_model
        get_backbone_model("CNN", {"CNN": {"in_channels": 1, "hidden_channels": 64, "pool_dim": [3, 3], "out_dim": 64}})


def test_shard_range_partitions():
    from afsl_b200.parallel import shard_range
    for total in (0, 1, 7, 2000, 65536):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from afsl_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    torch.manual_seed(100 + rank)                      # replicas start different; broadcast must fix that
    net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 2))
    dp = parallel.EpisodeDataParallel(net)
    # every rank owns a contiguous block of the step's episodes
    total = 7
    lo, hi = parallel.shard_range(total, r, w)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(total, 5, 6, generator=gen)
    loss = torch.stack([net(x[i]).pow(2).mean() for i in range(lo, hi)]).sum() / total * w   # mean over ALL episodes, times world
    loss.backward()
    dp.sync_gradients()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    acc = parallel.gather_accuracies(np.arange(lo, hi, dtype=np.float64) / 10.0, total)
    if rank == 0:
        torch.save({"grad": flat, "acc": acc, "w0": net[0].weight.detach().clone()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_ranks_match_single_process(tmp_path):
    port = 29500 + random.randint(0, 2000)
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    # single process reference: rank 0's initial weights, all 7 episodes
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 2))
    assert torch.equal(net[0].weight, got["w0"])
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(7, 5, 6, generator=gen)
    torch.stack([net(x[i]).pow(2).mean() for i in range(7)]).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    torch.testing.assert_close(got["grad"], want, rtol=1e-5, atol=1e-7)
    assert np.array_equal(got["acc"], np.arange(7, dtype=np.float64) / 10.0)


# ------------------------------------------------------------------ episode sampler vs the reference's sample_episode
class _FakeMultiSegDataset:
    """Same construction as tests/golden/make_golden.py::FakeMultiSegDataset (clips of 1..4 segments)."""

    def __init__(self, seed, classes=7, per_class=9, t_len=32):
        import pandas as pd
        g = torch.Generator().manual_seed(seed)
        n = classes * per_class
        self.segments = torch.randint(1, 5, (n,), generator=g).tolist()
        self.clips = [torch.randn(sg, 1, 128, t_len, generator=g) for sg in self.segments]
        names = [f"c{i}" for i in range(classes)]
        self.class_to_label = {nm: i for i, nm in enumerate(names)}
        order = torch.randperm(n, generator=g).tolist()
        self.data_df = pd.DataFrame({"label": [names[i % classes] for i in order], "index_column": order})
        self.multi_segm, self.input_type, self.specaug_use, self.waveaug_use = True, "spec", False, False
        self.experiment_config = {"specaug_params": {"use": False}}

    def __getitem__(self, i):
        return self.clips[i], 0


def test_sample_episode_matches_reference_draws():
    """sample_episode picks the clips / segments, labels and audio ids of the reference's sample_episode
    (datasets/batch_creation.py:21-170) draw for draw, and leaves Python's RNG in the same state."""
    import random
    from conftest import load_golden
    from afsl_b200.datasets.batch_creation import sample_episode, sample_episode_batch
    g = load_golden("sampler_multiseg")
    ds = _FakeMultiSegDataset(int(g["dataset_seed"]))
    fp = lambda x: x[:, 0, :2, :4].reshape(x.shape[0], -1)
    for name, is_test in (("train", False), ("test", True), ("train2", False)):
        random.seed(int(g[f"{name}_seed"]))
        s_list, s_lab, q_list, q_lab, ids = sample_episode(ds, 5, 3, 4, is_test, "cpu", None, False)
        assert len(s_list) == 1 and len(q_list) == 1
        assert torch.equal(fp(s_list[0]), torch.from_numpy(g[f"{name}_support"]))
        assert torch.equal(fp(q_list[0]), torch.from_numpy(g[f"{name}_query"]))
        assert torch.equal(s_lab, torch.from_numpy(g[f"{name}_support_labels"]))
        assert torch.equal(q_lab, torch.from_numpy(g[f"{name}_query_labels"]))
        assert torch.equal(ids, torch.from_numpy(g[f"{name}_audio_ids"]))
        assert random.random() == float(g[f"{name}_state_after"])
    # batched sampler: episode e of the batch == the e-th sample_episode call of the same stream
    random.seed(int(g["train_seed"]))
    batch = sample_episode_batch(ds, 1, 5, 3, 4)
    assert torch.equal(fp(batch.support[0]), torch.from_numpy(g["train_support"]))
    assert torch.equal(fp(batch.query[0]), torch.from_numpy(g["train_query"]))
    assert torch.equal(batch.support_labels[0], torch.from_numpy(g["train_support_labels"]))
    with pytest.raises(ValueError, match="Not enough samples"):
        sample_episode(ds, 5, 8, 4, False, "cpu", None, False)


def test_early_stopping_protocol(tmp_path):
    from afsl_b200.callbacks.early_stopping import EarlyStopping
    model = torch.nn.Linear(2, 2)
    msgs = []
    stop = EarlyStopping(patience=3, verbose=True, path=str(tmp_path / "model.pt"), trace_func=msgs.append)
    for epoch, acc in enumerate([0.5, 0.6, 0.55, 0.58, 0.59], 1):
        stop(acc, model, epoch)
    assert stop.early_stop and stop.counter == 3 and stop.best_score == 0.6
    assert (tmp_path / "model.pt").exists() and any("EarlyStopping counter: 2 out of 3" in m for m in msgs)


class _FakeWavDataset:
    """Same construction as tests/golden/make_golden.py::FakeWavDataset (16 kHz waveforms of 1.5 .. 12 s)."""

    def __init__(self, seed, classes=6, per_class=8):
        import pandas as pd
        rng = np.random.RandomState(seed)
        n = classes * per_class
        lengths = rng.randint(24000, 192000, size=n)
        self.wavs = [(rng.randn(int(length)) * 0.1).astype(np.float32) for length in lengths]
        names = [f"c{i}" for i in range(classes)]
        self.class_to_label = {nm: i for i, nm in enumerate(names)}
        self.data_df = pd.DataFrame({"label": [names[i // per_class] for i in range(n)], "index_column": list(range(n))})
        self.multi_segm, self.input_type, self.specaug_use, self.waveaug_use = True, "wav", False, False
        self.experiment_config = {"specaug_params": {"use": False}}

    def __getitem__(self, i):
        return self.wavs[i], 0

    def get_normalization_stats(self):
        return -21.5, 13.25


def test_sample_episode_waveform_input_matches_reference():
    """input_type 'wav': 5-second splits, the caller's MelSpectrogram, 10*log10, global normalisation - the episode
    tensors, labels and audio ids of the reference's sample_episode on the same fake dataset and seeds."""
    import random
    torchaudio = pytest.importorskip("torchaudio")
    from conftest import load_golden
    from afsl_b200.datasets.batch_creation import sample_episode, variable_wav_splits
    g = load_golden("sampler_wav")
    ds = _FakeWavDataset(int(g["dataset_seed"]))
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_mels=128, n_fft=1024, hop_length=512, power=2.0)
    fp = lambda x: x[:, 0, ::16, ::20].reshape(x.shape[0], -1)
    for name, is_test in (("train", False), ("test", True)):
        random.seed(int(g[f"{name}_seed"]))
        s_list, s_lab, q_list, q_lab, ids = sample_episode(ds, 4, 2, 3, is_test, "cpu", mel, False)
        assert list(q_list[0].shape) == g[f"{name}_shape"].tolist()
        torch.testing.assert_close(fp(s_list[0]), torch.from_numpy(g[f"{name}_support"]), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(fp(q_list[0]), torch.from_numpy(g[f"{name}_query"]), rtol=1e-5, atol=1e-5)
        assert torch.equal(s_lab, torch.from_numpy(g[f"{name}_support_labels"]))
        assert torch.equal(q_lab, torch.from_numpy(g[f"{name}_query_labels"]))
        assert torch.equal(ids, torch.from_numpy(g[f"{name}_audio_ids"]))
    # splits: short clips are tiled to 5 s, long clips are cut and the remainder piece is tiled
    assert [p.shape[0] for p in variable_wav_splits(np.ones(30000, np.float32))] == [80000]
    assert [p.shape[0] for p in variable_wav_splits(np.ones(170000, np.float32))] == [80000, 80000, 80000]
    ds.waveaug_use = True
    with pytest.raises(NotImplementedError):
        sample_episode(ds, 4, 2, 3, False, "cpu", mel, False)


# ------------------------------------------------------------------ epoch loop / early stopping vs the reference's control flow
@pytest.mark.parametrize("name", ["improves_then_stalls", "never_stops", "stops_at_patience_one"])
def test_contrastive_training_loop_control_flow_matches_reference(name, tmp_path):
    """loops.contrastive_training_loop + callbacks.EarlyStopping driven by scripted validation accuracies do exactly what the
    REFERENCE's loop did with the same script (tests/golden/training_loop_control_flow.json, generated by running
    /root/reference/loops/loops.py:124-167 itself): same printed lines (epoch headers, loss dictionaries, checkpoint and
    early-stopping messages), same number of epochs, the best checkpoint reloaded into the returned model, same scheduler
    position and learning rate."""
    import json
    from conftest import GOLDEN, golden_module
    import afsl_b200.loops.loops as loops
    with open(os.path.join(GOLDEN, "training_loop_control_flow.json")) as fh:
        want = json.load(fh)[name]
    mg = golden_module()
    script, patience, epochs = mg.LOOP_SCRIPTS[name]
    got = mg.scripted_loop(loops.contrastive_training_loop, loops, script, patience, epochs, str(tmp_path))
    assert got["lines"] == want["lines"]
    for key in ("weight", "train_calls", "val_calls", "scheduler_last_epoch", "lr"):
        assert got[key] == want[key], key
