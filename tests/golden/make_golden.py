"""Generate the golden fixtures by executing the REAL reference on CPU.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference is imported unmodified with empty stubs for third-party modules
that are not installed (pytorch_metric_learning, matplotlib, torch_audiomentations,
audiomentations, librosa - none is on the paths exercised here).  Every fixture
records inputs, the reference's outputs and (where autograd applies) the
reference's gradients.  The GPU box has no /root/reference; tests read only the
committed .npz files.
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AFSL_REFERENCE", "/root/reference")


def _stub(name, classes=()):
    mod = types.ModuleType(name)
    for c in classes:
        setattr(mod, c, type(c, (), {"__init__": lambda self, *a, **k: None}))
    sys.modules[name] = mod
    return mod


def import_reference():
    pml = _stub("pytorch_metric_learning")
    pml.losses = _stub("pytorch_metric_learning.losses", ["AngularLoss"])
    pml.miners = _stub("pytorch_metric_learning.miners", ["AngularMiner"])
    _stub("matplotlib").pyplot = _stub("matplotlib.pyplot")
    _stub("torch_audiomentations",
          "Compose Gain PolarityInversion AddColoredNoise BandPassFilter BandStopFilter HighPassFilter "
          "LowPassFilter PitchShift Shift SpliceOut TimeInversion PeakNormalization AddBackgroundNoise".split())
    _stub("audiomentations")
    _stub("librosa")
    sys.path.insert(0, REF)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: tuple(v.shape) for k, v in out.items() if not k.startswith(("w_", "init_", "after_"))})


# ------------------------------------------------------------------ head
HEAD_CASES = [
    # name, ways, shots, queries-per-class, dim, views (labels repeated when > 1)
    ("5w5s5q_d64", 5, 5, 5, 64, 1),
    ("5w5s5q_d256", 5, 5, 5, 256, 1),
    ("5w1s5q_d64", 5, 1, 5, 64, 1),
    ("20w5s5q_d256", 20, 5, 5, 256, 1),
    ("20w1s5q_d64", 20, 1, 5, 64, 1),
    ("5w5s5q_d64_v4", 5, 5, 5, 64, 4),
]


def gen_head():
    from loops.loss import FSL_Loss
    from models.util_functions import compute_prototypes
    for idx, (name, w, k, q, d, v) in enumerate(HEAD_CASES):
        torch.manual_seed(100 + idx)
        sl = torch.arange(w).repeat_interleave(k).repeat(v)
        ql = torch.arange(w).repeat_interleave(q).repeat(v)
        s = torch.randn(sl.numel(), d, requires_grad=True)
        qf = torch.randn(ql.numel(), d, requires_grad=True)
        protos = compute_prototypes(s, sl)
        scores = -torch.cdist(qf, protos)
        loss = FSL_Loss()(protos, qf, ql)
        loss.backward()
        post, pred = torch.max(scores, 1)
        save("head_" + name, support=s, support_labels=sl, query=qf, query_labels=ql, prototypes=protos,
             scores=scores, loss=loss, d_support=s.grad, d_query=qf.grad, pred=pred, posterior=post,
             correct=(pred == ql).sum())
    # ragged (multi-segment shaped) queries with shuffled support order and an exact-hit query
    torch.manual_seed(177)
    w, k, d = 5, 5, 64
    sl = torch.arange(w).repeat_interleave(k)[torch.randperm(w * k)]
    seg_counts = [1, 3, 2, 5, 1, 4, 2, 2, 3, 1, 6, 1, 2, 3, 1, 1, 2, 4, 1, 2, 3, 2, 1, 5, 2]
    ql = torch.cat([torch.full((c,), i // 5) for i, c in enumerate(seg_counts)])
    ids = torch.cat([torch.full((c,), i) for i, c in enumerate(seg_counts)])
    s = torch.randn(w * k, d, requires_grad=True)
    qf = torch.randn(ql.numel(), d)
    with torch.no_grad():
        qf[7] = compute_prototypes(s, sl)[int(ql[7])]            # zero distance -> zero gradient
    qf.requires_grad_(True)
    protos = compute_prototypes(s, sl)
    scores = -torch.cdist(qf, protos)
    loss = FSL_Loss()(protos, qf, ql)
    loss.backward()
    post, pred = torch.max(scores, 1)
    save("head_ragged_d64", support=s, support_labels=sl, query=qf, query_labels=ql, clip_ids=ids,
         prototypes=protos, scores=scores, loss=loss, d_support=s.grad, d_query=qf.grad, pred=pred,
         posterior=post, correct=(pred == ql).sum())


# ------------------------------------------------------------------ CPL
CPL_CASES = [
    # name, ways, per-class queries, dim, M, T, normalize prototypes first
    ("5w5q_d256_m5", 5, 5, 256, 5, 9.2361, False),
    ("5w5q_d256_m2", 5, 5, 256, 2, 9.2361, False),
    ("5w20q_d64_m5", 5, 20, 64, 5, 6.0488, True),
    ("20w5q_d256_m3", 20, 5, 256, 3, 4.081, False),
]


def replay_keep(labels, m):
    """Replay the reference's randperm draws (loops/loss.py:134-165) -> keep[i, j]."""
    uniq = labels.unique()
    groups = {int(c): torch.where(labels == c)[0] for c in uniq}
    n = labels.numel()
    keep = torch.zeros(n, n, dtype=torch.bool)
    for i in range(n):
        for c, members in groups.items():
            if c != int(labels[i]):
                keep[i, members[torch.randperm(len(members))[:m]]] = True
        keep[i, i] = True
    return keep


def gen_cpl():
    from loops.loss import CPL_Loss
    for idx, (name, w, q, d, m, t, norm) in enumerate(CPL_CASES):
        torch.manual_seed(200 + idx)
        labels = torch.arange(w).repeat_interleave(q)
        protos = torch.randn(w, d)
        if norm:
            protos = torch.nn.functional.normalize(protos, p=2.0, dim=1, eps=1e-12)
        protos.requires_grad_(True)
        queries = torch.randn(w * q, d, requires_grad=True)
        seed = 9000 + idx
        torch.manual_seed(seed)
        loss = CPL_Loss(T=t, M=m)(protos, queries, labels)
        loss.backward()
        torch.manual_seed(seed)
        keep = replay_keep(labels, m)
        save("cpl_" + name, prototypes=protos, queries=queries, labels=labels, keep=keep, loss=loss,
             d_prototypes=protos.grad, d_queries=queries.grad, seed=seed, m=m, temperature=t)


# ------------------------------------------------------------------ SpecAugment
SPEC_CASES = [
    ("fsd_t157", 2, 157, dict(mask_param=16, W=22, num_mask=1, mask_value=0, p=0.282)),
    ("nsynth_t126", 2, 126, dict(mask_param=9, W=36, num_mask=1, mask_value=0, p=0.42157)),
    ("esc_t157_2masks", 3, 157, dict(mask_param=7, W=20, num_mask=2, mask_value=-1.5, p=0.3127)),
]


def gen_specaug():
    from utils.augmentations import SpecAugment
    for idx, (name, n, t_len, sp) in enumerate(SPEC_CASES):
        seed = 300 + idx
        torch.manual_seed(seed)
        x = torch.randn(n, 1, 128, t_len)
        torch.manual_seed(seed + 50)
        np.random.seed(seed + 50)
        aug = SpecAugment({"specaug_params": sp})
        views = aug.apply_augmentations(x)
        # replay the draws to record the parameters the reference used
        torch.manual_seed(seed + 50)
        np.random.seed(seed + 50)
        w = sp["W"]
        warp_p = torch.randint(w, t_len - w, (n,))
        warp_d = torch.randint(-w, w, (n,))
        tm, fm = [], []
        for _ in range(sp["num_mask"]):
            t = np.random.randint(1, min(sp["mask_param"], int(sp["p"] * t_len)) + 1)
            tm.append((np.random.randint(0, t_len - t), t))
        for _ in range(sp["num_mask"]):
            f = np.random.randint(1, sp["mask_param"] + 1)
            fm.append((np.random.randint(0, 128 - f), f))
        # the reference's own spline evaluation for these control points
        xs = torch.linspace(0, t_len - 1, t_len).unsqueeze(0).expand(n, -1)
        cx = torch.stack([torch.tensor([0]).expand(n), warp_p, torch.tensor([t_len - 1]).expand(n)], 1)
        cy = torch.stack([torch.tensor([-1.]).expand(n), (warp_p - warp_d) * 2 / (t_len - 1) - 1,
                          torch.tensor([1]).expand(n)], 1)
        src_x = aug.hspline_interpolate_1D(cx, cy, xs)
        save("specaug_" + name, x=x, warped=views[1], time_masked=views[2], freq_masked=views[3],
             original=views[0], warp_p=warp_p, warp_d=warp_d, time_masks=np.array(tm), freq_masks=np.array(fm),
             src_x=src_x, seed=seed + 50, **{"cfg_" + k: v for k, v in sp.items()})


# ------------------------------------------------------------------ majority vote
def gen_vote():
    from loops.loops import calculate_majority_vote_accuracy
    rng = np.random.RandomState(4242)
    preds, ids, labels, posts, offsets, accs = [], [], [], [], [0], []
    for case in range(120):
        clips = rng.randint(1, 26)
        ways = 5 if case % 2 == 0 else 20
        seg = rng.randint(1, 9, size=clips) if case % 3 else np.minimum(36, 1 + rng.geometric(0.15, size=clips))
        cid = np.repeat(np.arange(clips), seg)
        if case % 5 == 0:                                   # interleave the segments of different clips
            cid = cid[rng.permutation(cid.size)]
        true = rng.randint(0, ways, size=clips)[cid]
        few = rng.randint(0, min(ways, 3), size=cid.size)   # few distinct votes -> many ties
        pred = np.where(rng.rand(cid.size) < 0.5, true, few)
        post = -rng.rand(cid.size).astype(np.float32) * 10
        if case % 7 == 0:
            post = np.round(post)                           # equal posteriors among tied segments
        row = []
        for strat in ("", "min_label", "max_posterior"):
            row.append(calculate_majority_vote_accuracy(torch.from_numpy(pred), torch.from_numpy(cid),
                                                        torch.from_numpy(true), torch.from_numpy(post),
                                                        tie_strategy=strat))
        preds.append(pred); ids.append(cid); labels.append(true); posts.append(post)
        offsets.append(offsets[-1] + cid.size); accs.append(row)
    save("vote_cases", pred=np.concatenate(preds), clip_ids=np.concatenate(ids), labels=np.concatenate(labels),
         posterior=np.concatenate(posts), offsets=np.array(offsets), accuracy=np.array(accs, dtype=np.float64))


# ------------------------------------------------------------------ modules
def gen_modules():
    from models.main_modules import SelfAttention, ProjectionHead, StandardCNN, StandardHybrid
    from models.prototypical import ContrastivePrototypicalNetworks, ContrastivePrototypicalNetworksWithoutAttention
    mc = {"Attention": {"embed_dim": 64, "num_heads": 1, "ffn_dim": 256, "dropout": 0.0},
          "Projection": {"input_dim": 256, "hidden_dim": 128, "output_dim": 256}}
    torch.manual_seed(400)
    att = SelfAttention(mc)
    att.train()                                            # dropout 0.0 -> deterministic slow path
    x = torch.randn(7, 4, 64, requires_grad=True)
    y = att(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    grads = {"g_" + k: p.grad for k, p in att.named_parameters()}
    att.eval()
    with torch.no_grad():
        y_eval = att(x)                                    # fused fast path
    save("modules_fusion", x=x, y=y, gy=gy, dx=x.grad, y_eval=y_eval,
         **{"w_" + k: v for k, v in att.state_dict().items()}, **grads)

    torch.manual_seed(401)
    proj = ProjectionHead(mc)
    x = torch.randn(9, 256, requires_grad=True)
    y = proj(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    save("modules_projection", x=x, y=y, gy=gy, dx=x.grad,
         **{"w_" + k: v for k, v in proj.state_dict().items()},
         **{"g_" + k: p.grad for k, p in proj.named_parameters() if p.grad is not None})

    for enc_name, t_len in (("CNN", 157), ("Hybrid", 157), ("Hybrid", 126)):
        torch.manual_seed(402)
        if enc_name == "CNN":
            enc = StandardCNN(1, (1, 1, 128, t_len), 64, [3, 3], 64)
        else:
            enc = StandardHybrid(1, 1, "RNN", False, 64, [3, 3], 64)
        x = torch.randn(4, 1, 128, t_len)
        enc.train()
        torch.manual_seed(7)
        for m in enc.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        y_train = enc(x)
        enc.eval()
        with torch.no_grad():
            y_eval = enc(x)
        save(f"modules_encoder_{enc_name}_t{t_len}", x=x, y_train=y_train, y_eval=y_eval,
             keys=np.array(list(enc.state_dict().keys())),
             **{"w_" + k: v for k, v in enc.state_dict().items()})

    # model-level: fused-views model in eval mode, 5w2s3q episode on given per-view features
    class Passthrough(torch.nn.Module):
        def forward(self, views):
            return list(views)
    torch.manual_seed(403)
    att = SelfAttention(mc).eval()
    proj = ProjectionHead(mc).eval()
    net = ContrastivePrototypicalNetworks(Passthrough(), att, proj).eval()
    sl = torch.arange(5).repeat_interleave(2)
    ql = torch.arange(5).repeat_interleave(3)
    s_views = [torch.randn(10, 64) for _ in range(4)]
    q_views = [torch.randn(15, 64) for _ in range(4)]
    with torch.no_grad():
        net.process_support_set(s_views, sl)
        scores = net(q_views, inference=True)
        random.seed(11)
        cf, cp = net.contrastive_forward(True)
    save("modules_model_fused", support_views=torch.stack(s_views), query_views=torch.stack(q_views),
         support_labels=sl, query_labels=ql, prototypes=net.prototypes, scores=scores, contrastive_features=cf,
         contrastive_prototypes=cp, shuffle_seed=11,
         **{"w_att_" + k: v for k, v in att.state_dict().items()},
         **{"w_proj_" + k: v for k, v in proj.state_dict().items()})
    net2 = ContrastivePrototypicalNetworksWithoutAttention(Passthrough(), proj).eval()
    with torch.no_grad():
        net2.process_support_set([v for v in s_views], sl.repeat(4))
        scores2 = net2([v for v in q_views], inference=True)
    save("modules_model_concat", prototypes=net2.prototypes, scores=scores2,
         keys_fused=np.array(list(net.state_dict().keys())), keys_concat=np.array(list(net2.state_dict().keys())))


# ------------------------------------------------------------------ full reference training epoch
class FakeDataset:
    """Satisfies what datasets/batch_creation.py::sample_episode reads (:22-23,38,50,112)."""

    def __init__(self, cfg, classes=6, per_class=12, t_len=157, seed=0):
        import pandas as pd
        g = torch.Generator().manual_seed(seed)
        self.items = torch.randn(classes * per_class, 1, 1, 128, t_len, generator=g)   # __getitem__ -> [S=1,1,128,T]
        names = [f"c{i}" for i in range(classes)]
        self.class_to_label = {n: i for i, n in enumerate(names)}
        self.data_df = pd.DataFrame({"label": [names[i // per_class] for i in range(classes * per_class)],
                                     "index_column": list(range(classes * per_class))})
        self.multi_segm = False
        self.input_type = "spec"
        self.specaug_use = cfg["specaug_params"]["use"]
        self.waveaug_use = False
        self.experiment_config = cfg

    def __getitem__(self, i):
        return self.items[i], 0


def gen_epoch():
    """Two reference training episodes + validation on a fake dataset, config-1 shaped (CNN, no views)."""
    from loops.loops import training_epoch, evaluate_single_segment
    from loops.loss import FSL_Loss
    from models.main_modules import StandardCNN, ProjectionHead
    from models.prototypical import ContrastivePrototypicalNetworksWithoutAttention
    cfg = {"specaug_params": {"use": False, "mask_param": 16, "W": 22, "num_mask": 1, "mask_value": 0, "p": 0.282}}
    mc = {"Projection": {"input_dim": 64, "hidden_dim": 32, "output_dim": 64}}
    ds = FakeDataset(cfg, t_len=157, seed=5)

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = StandardCNN(1, (1, 1, 128, 157), 64, [3, 3], 64)
        def forward(self, views):
            return [self.encoder(v) for v in views]
    torch.manual_seed(500)
    net = ContrastivePrototypicalNetworksWithoutAttention(Enc(), ProjectionHead(mc))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    init = {k: v.clone() for k, v in net.state_dict().items()}
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    random.seed(1234); np.random.seed(1234); torch.manual_seed(1234)
    msg = training_epoch(net, ds, opt, 2, "cpu", FSL_Loss(), None, 0.0, False, False, 5, 5, 5, None, False, False)
    random.seed(99)
    acc = evaluate_single_segment(net, ds, 3, "cpu", 5, 5, 5, None, False)
    after = {k: v.clone() for k, v in net.state_dict().items()}
    save("epoch_cnn_plain", items_seed=5, items_checksum=ds.items.double().sum(), loss=msg["loss"], fsl_loss=msg["fsl_loss"],
         val_mean=acc[0], val_std=acc[1],
         **{"init_" + k: v for k, v in init.items()}, **{"after_" + k: v for k, v in after.items()})


def gen_epoch_first():
    """First episode of gen_epoch alone (same seeds, same initial weights): its loss does not go through Adam,
    so it pins the forward of the whole loop (sampler -> encoder -> prototypes -> FSL loss) tightly."""
    from loops.loops import training_epoch
    from loops.loss import FSL_Loss
    from models.main_modules import StandardCNN, ProjectionHead
    from models.prototypical import ContrastivePrototypicalNetworksWithoutAttention
    cfg = {"specaug_params": {"use": False, "mask_param": 16, "W": 22, "num_mask": 1, "mask_value": 0, "p": 0.282}}
    mc = {"Projection": {"input_dim": 64, "hidden_dim": 32, "output_dim": 64}}
    ds = FakeDataset(cfg, t_len=157, seed=5)

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = StandardCNN(1, (1, 1, 128, 157), 64, [3, 3], 64)
        def forward(self, views):
            return [self.encoder(v) for v in views]
    torch.manual_seed(500)
    net = ContrastivePrototypicalNetworksWithoutAttention(Enc(), ProjectionHead(mc))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    random.seed(1234); np.random.seed(1234); torch.manual_seed(1234)
    msg = training_epoch(net, ds, opt, 1, "cpu", FSL_Loss(), None, 0.0, False, False, 5, 5, 5, None, False, False)
    save("epoch_cnn_first", loss=msg["loss"], fsl_loss=msg["fsl_loss"])


class FakeMultiSegDataset(FakeDataset):
    """Clips with 1..4 segments each (multi_segm datasets return [S,1,128,T] per item)."""

    def __init__(self, cfg, classes=7, per_class=9, t_len=32, seed=0):
        import pandas as pd
        g = torch.Generator().manual_seed(seed)
        n = classes * per_class
        self.segments = torch.randint(1, 5, (n,), generator=g).tolist()
        self.clips = [torch.randn(sg, 1, 128, t_len, generator=g) for sg in self.segments]
        names = [f"c{i}" for i in range(classes)]
        self.class_to_label = {nm: i for i, nm in enumerate(names)}
        order = torch.randperm(n, generator=g).tolist()            # rows of data_df are not grouped by class
        self.data_df = pd.DataFrame({"label": [names[i % classes] for i in order], "index_column": order})
        self.multi_segm = True
        self.input_type = "spec"
        self.specaug_use = False
        self.waveaug_use = False
        self.experiment_config = cfg

    def __getitem__(self, i):
        return self.clips[i], 0


def gen_sampler():
    """The reference's sample_episode on a multi-segment fake dataset: which clips / segments it selects, in which
    order, with which labels and audio ids, for train (is_test=False) and test (is_test=True) episodes."""
    from datasets.batch_creation import sample_episode
    cfg = {"specaug_params": {"use": False}}
    ds = FakeMultiSegDataset(cfg, seed=11)
    out = {}
    for name, is_test, seed in (("train", False, 21), ("test", True, 22), ("train2", False, 23)):
        random.seed(seed)
        s_list, s_lab, q_list, q_lab, ids = sample_episode(ds, 5, 3, 4, is_test, "cpu", None, False)
        assert len(s_list) == 1 and len(q_list) == 1
        fp = lambda x: x[:, 0, :2, :4].reshape(x.shape[0], -1).clone()     # 8 values identify a segment
        out.update({f"{name}_support": fp(s_list[0]), f"{name}_support_labels": s_lab, f"{name}_query": fp(q_list[0]),
                    f"{name}_query_labels": q_lab, f"{name}_audio_ids": ids, f"{name}_seed": seed,
                    f"{name}_state_after": random.random()})
    save("sampler_multiseg", dataset_seed=11, **out)


class FakeWavDataset:
    """16 kHz waveforms of 1.5 .. 12 s as numpy arrays (input_type 'wav', multi-segment)."""

    def __init__(self, cfg, classes=6, per_class=8, seed=0):
        import pandas as pd
        rng = np.random.RandomState(seed)
        n = classes * per_class
        lengths = rng.randint(24000, 192000, size=n)
        self.wavs = [(rng.randn(int(length)) * 0.1).astype(np.float32) for length in lengths]
        names = [f"c{i}" for i in range(classes)]
        self.class_to_label = {nm: i for i, nm in enumerate(names)}
        self.data_df = pd.DataFrame({"label": [names[i // per_class] for i in range(n)], "index_column": list(range(n))})
        self.multi_segm, self.input_type, self.specaug_use, self.waveaug_use = True, "wav", False, False
        self.experiment_config = cfg

    def __getitem__(self, i):
        return self.wavs[i], 0

    def get_normalization_stats(self):
        return -21.5, 13.25


def gen_sampler_wav():
    """The reference's sample_episode on waveform input: 5-second splits, MelSpectrogram, dB, global normalisation."""
    import torchaudio
    from datasets.batch_creation import sample_episode
    ds = FakeWavDataset({"specaug_params": {"use": False}}, seed=3)
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_mels=128, n_fft=1024, hop_length=512, power=2.0)
    out = {}
    for name, is_test, seed in (("train", False, 31), ("test", True, 32)):
        random.seed(seed)
        s_list, s_lab, q_list, q_lab, ids = sample_episode(ds, 4, 2, 3, is_test, "cpu", mel, False)
        fp = lambda x: x[:, 0, ::16, ::20].reshape(x.shape[0], -1).clone()
        out.update({f"{name}_support": fp(s_list[0]), f"{name}_support_labels": s_lab, f"{name}_query": fp(q_list[0]),
                    f"{name}_query_labels": q_lab, f"{name}_audio_ids": ids, f"{name}_seed": seed,
                    f"{name}_shape": np.array(q_list[0].shape)})
    save("sampler_wav", dataset_seed=3, **out)


# ------------------------------------------------------------------ multi-segment evaluation loop (config 4)
def class_patterns(classes, t_len, seed, scale=0.4):
    """[classes, 1, 128, t_len]: one smooth rank-one pattern per class (deterministic in ``seed``)."""
    g = torch.Generator().manual_seed(10_000 + seed)
    f = torch.nn.functional.avg_pool1d(torch.randn(classes, 1, 128 + 8, generator=g), 9, 1)[:, 0]
    t = torch.nn.functional.avg_pool1d(torch.randn(classes, 1, t_len + 8, generator=g), 9, 1)[:, 0]
    return (scale * 9.0 * f.unsqueeze(2) * t.unsqueeze(1)).unsqueeze(1)


class FakeMultiSegSpecDataset(FakeMultiSegDataset):
    """FakeMultiSegDataset with full-length spectrograms (the conv stack needs T >= 81) and an optional SpecAugment config."""

    def __init__(self, cfg, classes=7, per_class=9, t_len=157, seed=0):
        super().__init__(cfg, classes=classes, per_class=per_class, t_len=t_len, seed=seed)
        self.specaug_use = bool(cfg["specaug_params"]["use"])
        # class structure (a rank-one time-frequency pattern per class on top of the noise), so that the tasks are not
        # pure chance and the argmax over prototypes is not decided by the last bits of near-identical distances
        patterns = class_patterns(classes, t_len, seed)
        for label_name, idx in zip(self.data_df["label"].tolist(), self.data_df["index_column"].tolist()):
            self.clips[idx] = self.clips[idx] + patterns[self.class_to_label[label_name]]


def _multiseg_models(kind):
    """(model, experiment_config): 'concat' = Conv4, no views (ContrastivePrototypicalNetworksWithoutAttention);
    'fused' = Hybrid + SpecAugment support/query views + self-attention fusion (ContrastivePrototypicalNetworks)."""
    from models.main_modules import EncoderModule, ProjectionHead, SelfAttention, StandardCNN
    from models.prototypical import ContrastivePrototypicalNetworks, ContrastivePrototypicalNetworksWithoutAttention
    if kind == "concat":
        cfg = {"encoder_name": "CNN", "specaug_params": {"use": False}}
        mc = {"Projection": {"input_dim": 64, "hidden_dim": 16, "output_dim": 64}}

        class Enc(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.encoder = StandardCNN(1, (1, 1, 128, 157), 64, [3, 3], 64)

            def forward(self, views):
                return [self.encoder(v) for v in views]
        return ContrastivePrototypicalNetworksWithoutAttention(Enc(), ProjectionHead(mc)), cfg
    cfg = {"encoder_name": "Hybrid",
           "specaug_params": {"use": True, "mask_param": 16, "W": 22, "num_mask": 1, "mask_value": 0, "p": 0.282}}
    mc = {"Hybrid": {"in_channels": 1, "seq_layers": 1, "seq_type": "RNN", "bidirectional": False, "hidden_channels": 64,
                     "pool_dim": [3, 3], "out_dim": 64},
          "Attention": {"embed_dim": 64, "num_heads": 1, "ffn_dim": 256, "dropout": 0.1},
          "Projection": {"input_dim": 256, "hidden_dim": 16, "output_dim": 64}}
    return ContrastivePrototypicalNetworks(EncoderModule(cfg, mc), SelfAttention(mc), ProjectionHead(mc)), cfg


def _perturb_bn(model, seed):
    """Non-trivial running statistics / affine parameters, as a trained checkpoint would have."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def gen_multiseg_eval():
    """The reference's own evaluate_multisegment_loop (loops/loops.py:250-283) on a multi-segment fake dataset, for the three
    tie strategies: per-task accuracies (recorded by wrapping the reference's calculate_majority_vote_accuracy, which still
    runs unmodified), the returned mean / std, the per-segment predictions and top-2 score margins.  The model is in eval
    mode, as after contrastive_training_loop (whose last call is the validation).  The sampling seed is the first one from 41
    whose smallest top-2 margin is >= 1e-4 (scores are ~0.1 in magnitude, fp32 convolution noise ~1e-7): the comparison
    of per-task accuracies across devices must not hinge on the last bits of near-tied distances."""
    import loops.loops as L
    for kind in ("concat", "fused"):
        torch.manual_seed(600 if kind == "concat" else 601)
        model, cfg = _multiseg_models(kind)
        _perturb_bn(model, 77)
        model.eval()
        ds = FakeMultiSegSpecDataset(cfg, seed=12)
        margins, preds = [], []

        def hook(_m, _inp, out):
            top2 = torch.topk(out, 2, dim=1).values
            margins.append(top2[:, 0] - top2[:, 1])
            preds.append(torch.max(out, 1)[1])
        handle = model.register_forward_hook(hook)
        per_task, out = [], {}
        original = L.calculate_majority_vote_accuracy

        def recording(*a, **k):
            acc = original(*a, **k)
            per_task.append(acc)
            return acc
        L.calculate_majority_vote_accuracy = recording
        n_tasks, ways, shots, queries = 6, 5, 3, 4
        augment = cfg["specaug_params"]["use"]

        def run(seed, strat):
            per_task.clear(); margins.clear(); preds.clear()
            random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
            return L.evaluate_multisegment_loop(ds, ways, shots, queries, n_tasks, model, "cpu", strat, None, augment)
        try:
            seed = 41
            while True:
                run(seed, "")
                if float(torch.cat(margins).min()) >= 1e-4:
                    break
                seed += 1
            for strat in ("", "min_label", "max_posterior"):
                msg = run(seed, strat)
                key = strat or "first"
                out[f"acc_{key}"] = np.array(per_task, dtype=np.float64)
                out[f"mean_{key}"] = msg["mean_accuracy"]
                out[f"std_{key}"] = msg["accuracy_std"]
        finally:
            L.calculate_majority_vote_accuracy = original
            handle.remove()
        save(f"multiseg_eval_{kind}", dataset_seed=12, rng_seed=seed, n_tasks=n_tasks, ways=ways, shots=shots, queries=queries,
             margins=torch.cat(margins), pred=torch.cat(preds), rows_per_task=np.array([m.numel() for m in margins]),
             **out, **{"w_" + k: v for k, v in model.state_dict().items()})


# ------------------------------------------------------------------ 2000-task accuracy identity (north-star target)
class FakeManyClassDataset(FakeDataset):
    """24 classes x 12 single-segment clips: enough for 20-way 5-shot 5-query tasks."""

    def __init__(self, cfg, seed=0, t_len=157):
        super().__init__(cfg, classes=24, per_class=12, t_len=t_len, seed=seed)
        patterns = class_patterns(24, t_len, seed)
        self.items = self.items + patterns.repeat_interleave(12, dim=0).unsqueeze(1)


def gen_tasks2000():
    """The reference's evaluate_single_segment (loops/loops.py:84-121) over 2000 sampled tasks for each of
    (W, K) in {5, 20} x {1, 5}, Q = 5, Conv4 ProtoNet in eval mode with fixed weights (BASELINE config 5 / config 1 model).
    Per-task accuracies are recovered from a forward hook on the model (scores -> argmax == labels is what
    evaluate_on_one_task computes; asserted against the returned mean / std).  Also stored: the eval-mode embedding of every
    dataset item (so the head can be checked on the reference's own embeddings) and the smallest top-2 margin per config."""
    import loops.loops as L
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(700)
    model, cfg = _multiseg_models("concat")
    _perturb_bn(model, 78)
    model.eval()
    ds = FakeManyClassDataset(cfg, seed=13)
    with torch.no_grad():
        emb = torch.cat([model.backbone([ds.items[i:i + 32, 0]])[0] for i in range(0, ds.items.shape[0], 32)])
    out = {}
    n_tasks = int(os.environ.get("AFSL_N_TASKS", "2000"))
    for ways, shots in ((5, 1), (5, 5), (20, 1), (20, 5)):
        accs, margins = [], []
        labels = torch.arange(ways).repeat_interleave(5)

        def hook(_m, _inp, scores):
            top2 = torch.topk(scores, 2, dim=1).values
            margins.append(float((top2[:, 0] - top2[:, 1]).min()))
            accs.append(int((torch.max(scores, 1)[1] == labels).sum()) / labels.numel())
        handle = model.register_forward_hook(hook)
        random.seed(1000 + ways * 10 + shots)
        mean, std = L.evaluate_single_segment(model, ds, n_tasks, "cpu", ways, shots, 5, None, False)
        handle.remove()
        accs = np.array(accs, dtype=np.float64)
        assert accs.size == n_tasks and np.mean(accs) == mean and np.std(accs) == std
        key = f"{ways}w{shots}s"
        out.update({f"acc_{key}": accs, f"mean_{key}": mean, f"std_{key}": std, f"margin_{key}": np.array(margins)})
        print(key, mean, std, min(margins), flush=True)
    save("tasks2000_cnn", dataset_seed=13, n_tasks=n_tasks, embeddings=emb, **out,
         **{"w_" + k: v for k, v in model.state_dict().items()})


# ------------------------------------------------------------------ epoch loop / early stopping control flow
LOOP_SCRIPTS = {
    # name: (validation accuracies per epoch, patience, epochs)
    "improves_then_stalls": ([0.30, 0.42, 0.41, 0.42, 0.40, 0.39, 0.38, 0.37], 3, 8),
    "never_stops": ([0.2, 0.3, 0.25, 0.35, 0.3], 4, 5),
    "stops_at_patience_one": ([0.5, 0.4, 0.6], 1, 3),
}


def scripted_loop(loop_fn, module, script, patience, epochs, tmpdir):
    """Run an epoch loop (the reference's or the mirror's contrastive_training_loop) with its training epoch and
    validation replaced by scripts: every "training epoch" adds 1 to the single weight of a toy model, validation returns
    the scripted accuracy.  Returns what the loop did: printed lines, weight of the returned model, scheduler steps."""
    import contextlib
    import io
    model = torch.nn.Linear(1, 1, bias=False)
    with torch.no_grad():
        model.weight.fill_(0.0)
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2, 4], gamma=0.5)
    calls = {"train": 0, "val": 0}

    def fake_training_epoch(model, **kw):
        calls["train"] += 1
        with torch.no_grad():
            model.weight.add_(1.0)
        return {"loss": 1.0 / calls["train"], "fsl_loss": 0.5, "cpl_loss": float("nan")}

    def fake_validation(model, **kw):
        calls["val"] += 1
        return script[calls["val"] - 1], 0.0
    saved = (module.training_epoch, module.evaluate_single_segment)
    module.training_epoch, module.evaluate_single_segment = fake_training_epoch, fake_validation
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            trained = loop_fn(model, None, None, opt, 1, 1, "cpu", None, None, 0.0, epochs, sched, patience, tmpdir, False, False,
                              5, 5, 5, 5, 5, 5, None, False, False, False)
    finally:
        module.training_epoch, module.evaluate_single_segment = saved
    lines = [ln.replace(tmpdir, "<dir>") for ln in out.getvalue().splitlines()]
    return {"lines": lines, "weight": float(trained.weight.item()), "train_calls": calls["train"], "val_calls": calls["val"],
            "scheduler_last_epoch": int(sched.last_epoch), "lr": float(opt.param_groups[0]["lr"])}


def gen_training_loop():
    """The REFERENCE's contrastive_training_loop + EarlyStopping (loops/loops.py:124-167, callbacks/early_stopping.py:15-70)
    driven by scripted validation accuracies: epochs run, early-stopping messages, checkpoint reload, scheduler steps."""
    import json
    import tempfile
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                                 # the reference predates numpy 2.0 (early_stopping.py:38)
    import loops.loops as L
    out = {}
    for name, (script, patience, epochs) in LOOP_SCRIPTS.items():
        with tempfile.TemporaryDirectory() as tmp:
            out[name] = scripted_loop(L.contrastive_training_loop, L, script, patience, epochs, tmp)
        print(name, out[name]["train_calls"], out[name]["weight"])
    with open(os.path.join(HERE, "training_loop_control_flow.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    import_reference()
    torch.set_num_threads(1)            # fixed reduction order for the fixtures
    which = sys.argv[1:] or ["head", "cpl", "specaug", "vote", "modules", "epoch", "epoch_first", "sampler", "sampler_wav", "multiseg_eval"]
    for name in which:
        globals()["gen_" + name]()
