"""Pin the CPU oracle (oracle/) to outputs of the real reference (tests/golden/*.npz).

CPU only.  The fixtures were produced by tests/golden/make_golden.py, which
imports and executes the unmodified reference.
"""
import random

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, t
from oracle import angular, head, modules, specaug, vote


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", golden_names("head_"))
def test_head_matches_reference(name):
    g = load_golden(name)
    s = t(g["support"]).requires_grad_(True)
    q = t(g["query"]).requires_grad_(True)
    sl, ql = t(g["support_labels"]), t(g["query_labels"])
    protos = head.prototypes(s, sl)
    scores = head.l2_scores(q, protos)
    loss = head.fsl_loss(protos, q, ql)
    loss.backward()
    assert torch.equal(protos.detach(), t(g["prototypes"]))
    assert torch.equal(scores.detach(), t(g["scores"]))
    assert loss.item() == float(g["loss"])
    torch.testing.assert_close(s.grad, t(g["d_support"]), rtol=0, atol=0)
    torch.testing.assert_close(q.grad, t(g["d_query"]), rtol=0, atol=0)
    correct, total = head.evaluate_task(scores, ql)
    assert correct == int(g["correct"]) and total == ql.numel()


def test_survey_known_answers():
    # SURVEY.md section 4 known-answer values, captured from the reference
    torch.manual_seed(0)
    p, q, y = torch.rand(5, 256), torch.rand(25, 256), torch.arange(5).repeat_interleave(5)
    assert head.fsl_loss(p, q, y).item() == pytest.approx(1.620690941810608, rel=1e-7)
    torch.manual_seed(0)
    assert head.cpl_loss_loop(p, q, y, 9.2361, 5).item() == pytest.approx(0.12177345901727676, rel=1e-6)


@pytest.mark.parametrize("name", golden_names("cpl_"))
def test_cpl_matches_reference(name):
    g = load_golden(name)
    labels = t(g["labels"])
    temperature, m, seed = float(g["temperature"]), int(g["m"]), int(g["seed"])
    # loop form: same RNG stream, same op order -> same bits
    p = t(g["prototypes"]).requires_grad_(True)
    q = t(g["queries"]).requires_grad_(True)
    torch.manual_seed(seed)
    loss = head.cpl_loss_loop(p, q, labels, temperature, m)
    loss.backward()
    assert loss.item() == float(g["loss"])
    torch.testing.assert_close(p.grad, t(g["d_prototypes"]), rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(q.grad, t(g["d_queries"]), rtol=1e-6, atol=1e-9)
    # RNG replay gives the recorded keep mask; closed form agrees to rounding (fp64 to 1e-6 of fp32 ref)
    torch.manual_seed(seed)
    keep = head.cpl_draw_keep(labels, m)
    assert torch.equal(keep, t(g["keep"]))
    p2 = t(g["prototypes"]).double().requires_grad_(True)
    q2 = t(g["queries"]).double().requires_grad_(True)
    closed = head.cpl_loss_closed(p2, q2, labels, keep, temperature)
    closed.backward()
    assert closed.item() == pytest.approx(float(g["loss"]), rel=2e-6)
    torch.testing.assert_close(p2.grad.float(), t(g["d_prototypes"]), rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(q2.grad.float(), t(g["d_queries"]), rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("name", golden_names("specaug_"))
def test_specaug_matches_reference(name):
    g = load_golden(name)
    cfg = {"specaug_params": {k[4:]: g[k].item() for k in g if k.startswith("cfg_")}}
    x = t(g["x"])
    torch.manual_seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    views, params = specaug.apply_augmentations(x, cfg)
    assert torch.equal(params.warp_p, t(g["warp_p"])) and torch.equal(params.warp_d, t(g["warp_d"]))
    assert params.time_masks == [tuple(r) for r in g["time_masks"].tolist()]
    assert params.freq_masks == [tuple(r) for r in g["freq_masks"].tolist()]
    src = specaug.warp_source_x(params.warp_p, params.warp_d, x.shape[-1])
    assert torch.equal(src, t(g["src_x"]))
    for got, key in zip(views, ("original", "warped", "time_masked", "freq_masked")):
        assert torch.equal(got, t(g[key])), key


def test_vote_matches_reference():
    g = load_golden("vote_cases")
    off = g["offsets"]
    for c in range(len(off) - 1):
        sl = slice(off[c], off[c + 1])
        for k, strat in enumerate(("", "min_label", "max_posterior")):
            acc = vote.majority_vote_accuracy(g["pred"][sl], g["clip_ids"][sl], g["labels"][sl],
                                              g["posterior"][sl], strat)
            assert acc == g["accuracy"][c, k], (c, strat)


def _load_weights(module, g, prefix):
    sd = {k[len(prefix):]: t(v) for k, v in g.items() if k.startswith(prefix)}
    module.load_state_dict(sd, strict=True)


def test_fusion_and_projection_match_reference():
    g = load_golden("modules_fusion")
    fuse = modules.ViewFusion(64, 1, 256, 0.0)
    _load_weights(fuse, g, "w_")
    fuse.train()
    x = t(g["x"]).requires_grad_(True)
    y = fuse(x)
    y.backward(t(g["gy"]))
    assert torch.equal(y.detach(), t(g["y"]))
    assert torch.equal(x.grad, t(g["dx"]))
    for k, p in fuse.named_parameters():
        assert torch.equal(p.grad, t(g["g_" + k])), k
    fuse.eval()
    with torch.no_grad():
        assert torch.equal(fuse(x), t(g["y_eval"]))
    g = load_golden("modules_projection")
    proj = modules.Projection(256, 128, 256)
    _load_weights(proj, g, "w_")
    x = t(g["x"]).requires_grad_(True)
    y = proj(x)
    y.backward(t(g["gy"]))
    assert torch.equal(y.detach(), t(g["y"])) and torch.equal(x.grad, t(g["dx"]))


@pytest.mark.parametrize("name", golden_names("modules_encoder_"))
def test_encoders_match_reference(name):
    g = load_golden(name)
    kind, t_len = name.split("_")[2], int(name.split("_t")[-1])
    enc = modules.build_encoder(kind, t_len)
    assert list(enc.state_dict().keys()) == [str(k) for k in g["keys"]]
    _load_weights(enc, g, "w_")
    for m in enc.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    x = t(g["x"])
    enc.train()
    torch.testing.assert_close(enc(x).detach(), t(g["y_train"]), rtol=1e-5, atol=1e-6)
    enc.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith("w_")})
    enc.eval()
    with torch.no_grad():
        torch.testing.assert_close(enc(x), t(g["y_eval"]), rtol=1e-5, atol=1e-6)


def test_models_match_reference():
    g = load_golden("modules_model_fused")

    class Passthrough(torch.nn.Module):
        def forward(self, views):
            return list(views)
    fuse, proj = modules.ViewFusion(64, 1, 256, 0.0).eval(), modules.Projection(256, 128, 256).eval()
    _load_weights(fuse, g, "w_att_")
    _load_weights(proj, g, "w_proj_")
    net = modules.FusedViewsNet(Passthrough(), fuse, proj).eval()
    sv, qv = list(t(g["support_views"])), list(t(g["query_views"]))
    with torch.no_grad():
        net.process_support_set(sv, t(g["support_labels"]))
        scores = net(qv, inference=True)
        random.seed(int(g["shuffle_seed"]))
        cf, cp = net.contrastive_forward(True)
    assert torch.equal(net.prototypes, t(g["prototypes"])) and torch.equal(scores, t(g["scores"]))
    assert torch.equal(cf, t(g["contrastive_features"])) and torch.equal(cp, t(g["contrastive_prototypes"]))
    g2 = load_golden("modules_model_concat")
    net2 = modules.ConcatViewsNet(Passthrough(), proj).eval()
    with torch.no_grad():
        net2.process_support_set(sv, t(g["support_labels"]).repeat(4))
        scores2 = net2(qv, inference=True)
    assert torch.equal(net2.prototypes, t(g2["prototypes"])) and torch.equal(scores2, t(g2["scores"]))
    assert [k for k in net.state_dict().keys()] == [str(k) for k in g2["keys_fused"]]
    assert [k for k in net2.state_dict().keys()] == [str(k) for k in g2["keys_concat"]]


# ---- angular: parity UNPINNED (no pytorch_metric_learning); internal consistency only -------------
@pytest.mark.parametrize("angle", [0.0, 15.0, 30.0])
def test_angular_closed_form_consistent(angle):
    torch.manual_seed(3)
    w, q, d = 5, 5, 64
    protos = torch.nn.functional.normalize(torch.randn(w, d), dim=1).double()
    queries = torch.nn.functional.normalize(torch.randn(w * q, d), dim=1).double()
    labels = torch.arange(w).repeat_interleave(q)
    full = angular.angular_loss_class(protos, queries, labels, angle, True)
    a, p, n = angular.mine(protos, torch.arange(w), queries, labels, angle, same_set=False)
    m_a = torch.bincount(a, minlength=w)
    w_q = torch.bincount(p, minlength=w * q) + torch.bincount(n, minlength=w * q)
    closed = angular.anchors_closed_form(protos, queries, labels, m_a, w_q)
    assert closed.item() == pytest.approx(full.item(), rel=1e-10)
    if angle == 0.0:
        assert m_a.tolist() == [100] * 5 and w_q.tolist() == [40] * 25
