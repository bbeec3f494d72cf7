"""GPU parity: libafsl kernels (through the C ABI) vs the golden fixtures of the real reference
and vs the CPU oracle on seeded inputs.

Tolerances: the north star asks for losses, distances and gradients within 1e-5 relative in fp32
and bit-exact masks / argmax labels / per-task accuracies.  rtol below is 1e-5 with an absolute
floor scaled to the tensor magnitude (cancellation-free 1e-5 of max|ref|).  Every `close()` records the largest
observed error as a fraction of its bound; the run prints the table and writes gpurun_out/parity_observed.json
(a copy per round lives under profiles/), so the slack of each tolerance is visible next to it.
"""
import numpy as np
import pytest
import torch

from conftest import golden_module, golden_names, load_golden, record_observed, t
from oracle import head as ohead
from oracle import specaug as ospec
from oracle import vote as ovote

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def close(got, ref, rtol=RTOL, scale=None):
    """rtol > 0: relative check with an absolute floor of rtol * max|ref| (or rtol * scale);
    rtol == 0: pure absolute check with atol = scale."""
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    if rtol == 0:
        atol = float(scale)
    else:
        atol = rtol * float(ref.abs().max() if scale is None else scale)
    atol = max(atol, 1e-12)
    if got.shape == ref.shape and ref.numel():
        # observed error as a fraction of the elementwise bound atol + rtol*|ref| (reported at the end of the run)
        finite = torch.isfinite(ref) & torch.isfinite(got)
        if finite.any():
            err = (got - ref).abs()[finite]
            bound = (atol + rtol * ref.abs())[finite]
            k = int(torch.argmax(err / bound))
            record_observed(float(err[k]), float(bound[k]))
    torch.testing.assert_close(got, ref, rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def ops():
    import afsl_b200.ops as ops
    return ops


def dev(a, dtype=None):
    return t(a, dtype).cuda()


# ------------------------------------------------------------------ head vs golden (reference outputs)
@pytest.mark.parametrize("name", golden_names("head_"))
def test_head_vs_reference(ops, name):
    g = load_golden(name)
    ways = int(g["prototypes"].shape[0])
    s = dev(g["support"]).requires_grad_(True)
    q = dev(g["query"]).requires_grad_(True)
    sl, ql = dev(g["support_labels"]), dev(g["query_labels"])
    # separate ops, as the reference API composes them
    protos = ops.prototypes(s, sl)
    assert protos.shape == (ways, s.shape[1])
    scores = ops.l2_scores(q, protos)
    loss = ops.proto_loss(protos, q, ql)
    loss.backward()
    close(protos, t(g["prototypes"]))
    close(scores, t(g["scores"]))
    close(loss, t(g["loss"]))
    close(s.grad, t(g["d_support"]))
    close(q.grad, t(g["d_query"]))
    # fused head
    s2 = dev(g["support"]).requires_grad_(True)
    q2 = dev(g["query"]).requires_grad_(True)
    loss2, protos2, correct = ops.proto_head(s2, sl, q2, ql, n_way=ways)
    loss2.backward()
    close(loss2, t(g["loss"]))
    close(protos2, t(g["prototypes"]))
    close(s2.grad, t(g["d_support"]))
    close(q2.grad, t(g["d_query"]))
    assert int(correct) == int(g["correct"])
    # evaluation head: argmax labels, posterior, #correct
    pred, post, corr, sc = ops.proto_eval(s2.detach(), sl, q2.detach(), ql, n_way=ways, want_scores=True)
    assert torch.equal(pred.cpu().long(), t(g["pred"]))
    assert int(corr) == int(g["correct"])
    close(post, t(g["posterior"]))
    close(sc, t(g["scores"]))


def test_head_scores_gradient_path(ops):
    """Gradient through the scores output (-cdist) alone, vs torch autograd on CPU."""
    g = load_golden("head_5w5s5q_d64")
    p = dev(g["prototypes"]).requires_grad_(True)
    q = dev(g["query"]).requires_grad_(True)
    w = torch.randn(25, 5, generator=torch.Generator().manual_seed(1))
    (ops.l2_scores(q, p) * w.cuda()).sum().backward()
    pc, qc = t(g["prototypes"]).requires_grad_(True), t(g["query"]).requires_grad_(True)
    (ohead.l2_scores(qc, pc) * w).sum().backward()
    close(p.grad, pc.grad)
    close(q.grad, qc.grad)


@pytest.mark.parametrize("path", ["warp", "cta"])
@pytest.mark.parametrize("ways,shots,nq,dim", [(5, 5, 5, 64), (5, 5, 5, 256), (20, 5, 5, 256), (5, 1, 15, 128), (3, 2, 4, 32),
                                               (20, 5, 5, 64), (10, 3, 7, 128), (2, 5, 40, 256)])
def test_head_batched_vs_oracle(ops, monkeypatch, ways, shots, nq, dim, path):
    """E episodes at once == the oracle applied episode by episode (incl. extra prototype gradient), through
    both kernel families: one warp per episode (registers; small W*D) and one CTA per episode (any shape)."""
    monkeypatch.setenv("AFSL_HEAD_WARP", "1" if path == "warp" else "0")
    monkeypatch.setenv("AFSL_HEAD_WIDE", "1" if path == "warp" else "0")     # many-way forward: register-batch kernel / lane groups
    monkeypatch.setenv("AFSL_HEAD_MMA", "1" if path == "warp" else "0")      # many-way forward: tcgen05 kernel / fp32-pipe kernels
    e = 37
    gen = torch.Generator().manual_seed(ways * 1000 + dim)
    s = torch.randn(e, ways * shots, dim, generator=gen)
    q = torch.randn(e, ways * nq, dim, generator=gen)
    sl = torch.stack([torch.arange(ways).repeat_interleave(shots)[torch.randperm(ways * shots, generator=gen)] for _ in range(e)])
    ql = torch.stack([torch.arange(ways).repeat_interleave(nq)[torch.randperm(ways * nq, generator=gen)] for _ in range(e)])
    wl = torch.rand(e, generator=gen) + 0.5
    wp = torch.randn(e, ways, dim, generator=gen) * 0.01
    sg, qg = s.cuda().requires_grad_(True), q.cuda().requires_grad_(True)
    loss, protos, correct = ops.proto_head(sg, sl.cuda(), qg, ql.cuda(), n_way=ways)
    ((loss * wl.cuda()).sum() + (protos * wp.cuda()).sum()).backward()
    for i in range(e):
        sc, qc = s[i].clone().requires_grad_(True), q[i].clone().requires_grad_(True)
        pr = ohead.prototypes(sc, sl[i])
        lo = ohead.fsl_loss(pr, qc, ql[i])
        (lo * wl[i] + (pr * wp[i]).sum()).backward()
        close(loss[i], lo)
        close(protos[i], pr)
        close(sg.grad[i], sc.grad)
        close(qg.grad[i], qc.grad)
        assert int(correct[i]) == ohead.evaluate_task(ohead.l2_scores(qc, pr), ql[i])[0]


@pytest.mark.parametrize("ways,shots,nq,dim", [(20, 5, 5, 256), (20, 1, 5, 256), (20, 5, 5, 64), (20, 1, 5, 64), (20, 5, 5, 128),
                                               (10, 3, 7, 128), (24, 2, 5, 256), (8, 4, 16, 64)])
def test_head_many_way_tensor_core_kernel(ops, monkeypatch, ways, shots, nq, dim):
    """proto_head_tma.cu / proto_head_mma.cu (tcgen05, split-precision TF32) on 333 tasks per launch - several tasks per
    persistent CTA, so every barrier phase, both accumulators and all ring stages are reused - against the oracle task by task
    (scores / posterior / loss 1e-5, prototypes exact to rounding, argmax and #correct equal) and against the fp32-pipe
    kernel on the same inputs; also the given-prototypes entry (afsl_proto_scores_fwd_f32)."""
    e = 333
    gen = torch.Generator().manual_seed(ways * 100 + dim + shots)
    s = torch.randn(e, ways * shots, dim, generator=gen)
    q = torch.randn(e, ways * nq, dim, generator=gen)
    sl = torch.stack([torch.arange(ways).repeat_interleave(shots)[torch.randperm(ways * shots, generator=gen)] for _ in range(e)])
    ql = torch.randint(0, ways, (e, ways * nq), generator=gen)
    out = {}
    # TMA-fed tcgen05 kernel with one / two k-blocks per ring stage (support blocks of <= 32 rows in ONE stage; "1o": one
    # stage per k-block group for them too), LDG-fed tcgen05 kernel, fp32-pipe kernel
    for mma in ("1", "1p", "1o", "2", "0"):
        monkeypatch.setenv("AFSL_HEAD_MMA", mma[0])
        monkeypatch.setenv("AFSL_HEAD_PAIR", "1" if mma == "1" else "2")
        monkeypatch.setenv("AFSL_HEAD_ONESUP", "0" if mma == "1o" else "1")       # "1": forced on for every D
        pred, post, correct, scores = ops.proto_eval(s.cuda(), sl.cuda(), q.cuda(), ql.cuda(), n_way=ways, want_scores=True)
        loss, protos, corr2 = ops.proto_head(s.cuda(), sl.cuda(), q.cuda(), ql.cuda(), n_way=ways)
        sc2 = ops.l2_scores(q.cuda(), protos)
        torch.cuda.synchronize()
        out[mma] = [x.cpu() for x in (pred, post, correct, scores, loss, protos, corr2, sc2)]
    oracle = []
    for i in range(e):
        pr = ohead.prototypes(s[i], sl[i])
        sc = ohead.l2_scores(q[i], pr)
        oracle.append((pr, sc, ohead.fsl_loss(pr, q[i], ql[i])))
    for mma in ("1", "1p", "1o", "2"):
        pred, post, correct, scores, loss, protos, corr2, sc2 = out[mma]
        assert torch.equal(correct, corr2)
        flips = 0
        for i in range(e):
            pr, sc, lo = oracle[i]
            po, pd = torch.max(sc, 1)
            close(protos[i], pr)
            close(scores.view(e, ways * nq, ways)[i], sc)
            close(sc2[i], sc)
            close(post.view(e, -1)[i], po)
            close(loss[i], lo)
            same = pred.view(e, -1)[i].long() == pd
            if not bool(same.all()):                      # only a tie closer than fp32 rounding of the distances may differ
                top2 = torch.topk(sc, 2, dim=1).values
                assert float((top2[:, 0] - top2[:, 1])[~same].max()) < 1e-5 * float(sc.abs().max())
                flips += int((~same).sum())
            else:
                assert int(correct[i]) == int((pd == ql[i]).sum())
        agree = float((pred == out["0"][0]).float().mean())
        print(f"[{ways}w{shots}s q{nq} D={dim}, AFSL_HEAD_MMA={mma}] argmax flips vs oracle on near-ties: {flips} of "
              f"{e * ways * nq} rows; agreement with the fp32-pipe kernel {agree:.6f}")
        assert flips <= 2 and agree > 0.9999
        close(scores, out["0"][3])


@pytest.mark.parametrize("n,a,b", [(7, 2184, 64), (7, 64, 2184), (3, 37, 5), (2, 4, 4), (5, 238, 64), (1, 130, 68)])
def test_layout_transpose(ops, n, a, b):
    """afsl_transpose_f32 (the channels-last -> NCHW switch in front of the fp32 convolutions) is a pure data movement:
    bit-equal to torch, vector and scalar paths, ragged tiles; nhwc_to_nchw and its backward against Tensor.contiguous()."""
    x = torch.randn(n, a, b, device="cuda")
    y = ops._Transpose.apply(x)
    assert y.shape == (n, b, a) and y.is_contiguous() and torch.equal(y, x.transpose(1, 2).contiguous())
    if b % 4 == 0:
        h = 2 if a % 2 == 0 else 1
        img = torch.randn(n, b, h, a // h, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        out = ops.nhwc_to_nchw(img)
        assert out.is_contiguous() and torch.equal(out, img.detach().contiguous())
        g = torch.randn_like(out)
        out.backward(g)
        assert torch.equal(img.grad, g)
        assert img.grad.is_contiguous(memory_format=torch.channels_last) or img.grad.is_contiguous()


@pytest.mark.parametrize("ways,ns,nq,dim,e", [(24, 128, 128, 256, 301), (9, 40, 26, 64, 5), (20, 77, 100, 128, 150),
                                              (16, 16, 128, 256, 149), (8, 128, 27, 64, 297)])
def test_head_many_way_tensor_core_unbalanced_and_limits(ops, monkeypatch, ways, ns, nq, dim, e):
    """The TMA + tcgen05 head at the edges of what it takes: the largest blocks (128 support rows, 128 queries, 24 ways),
    the smallest (26 queries), UNBALANCED classes (1 .. many rows per class, rows of a class scattered over the block:
    the bucket warp's lists), fewer tasks than SMs and a ragged last wave, every ring depth the launcher may pick, with and
    without the L2 prefetch - against the fp32-pipe kernel on the same inputs and the oracle on a sample of tasks."""
    gen = torch.Generator().manual_seed(ways * 1000 + ns + nq + dim)
    s = torch.randn(e, ns, dim, generator=gen)
    q = torch.randn(e, nq, dim, generator=gen)
    sl = torch.randint(0, ways, (e, ns), generator=gen)
    for i in range(e):                                         # every class has at least one row, somewhere in the block
        sl[i, torch.randperm(ns, generator=gen)[:ways]] = torch.arange(ways)
    ql = torch.randint(0, ways, (e, nq), generator=gen)
    out = {}
    for tag, env in (("ref", {"AFSL_HEAD_MMA": "0"}), ("tc", {"AFSL_HEAD_MMA": "1"}),
                     ("tc_pf", {"AFSL_HEAD_MMA": "1", "AFSL_HEAD_L2PF": "1"}), ("tc_nopf", {"AFSL_HEAD_MMA": "1", "AFSL_HEAD_L2PF": "0"}),
                     ("tc_ring2", {"AFSL_HEAD_MMA": "1", "AFSL_HEAD_RING": "2"}),
                     ("tc_single", {"AFSL_HEAD_MMA": "1", "AFSL_HEAD_PAIR": "1", "AFSL_HEAD_RING": "3"})):
        for k in ("AFSL_HEAD_MMA", "AFSL_HEAD_L2PF", "AFSL_HEAD_RING", "AFSL_HEAD_PAIR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        pred, post, correct, scores = ops.proto_eval(s.cuda(), sl.cuda(), q.cuda(), ql.cuda(), n_way=ways, want_scores=True)
        pred2, post2, correct2, _ = ops.proto_eval(s.cuda(), sl.cuda(), q.cuda(), ql.cuda(), n_way=ways)     # evaluation-only epilogue
        loss, protos, _ = ops.proto_head(s.cuda(), sl.cuda(), q.cuda(), ql.cuda(), n_way=ways)
        torch.cuda.synchronize()
        out[tag] = [x.cpu() for x in (pred, post, correct, scores, loss, protos, pred2, post2, correct2)]
    ref = out["ref"]
    for tag in ("tc", "tc_pf", "tc_nopf", "tc_ring2", "tc_single"):
        pred, post, correct, scores, loss, protos, pred2, post2, correct2 = out[tag]
        close(scores, ref[3])
        close(post, ref[1])
        close(post2, ref[1])
        close(loss, ref[4])
        close(protos, ref[5])
        assert torch.equal(pred, pred2) and torch.equal(correct, correct2)
        assert float((pred == ref[0]).float().mean()) > 0.9999
        if tag != "tc_single":                                 # ring depth and prefetch do not touch the arithmetic
            assert torch.equal(scores, out["tc"][3]) and torch.equal(pred, out["tc"][0])
    for i in range(0, e, max(1, e // 7)):
        pr = ohead.prototypes(s[i], sl[i])
        sc = ohead.l2_scores(q[i], pr)
        close(out["tc"][5][i], pr)
        close(out["tc"][3].view(e, nq, ways)[i], sc)
        close(out["tc"][4][i], ohead.fsl_loss(pr, q[i], ql[i]))


@pytest.mark.parametrize("ways,shots,dim,wide", [(5, 5, 64, 1), (20, 5, 256, 1), (20, 5, 256, 0), (20, 1, 64, 1), (20, 1, 64, 0),
                                                 (13, 3, 128, 1)])
def test_head_ragged_tasks(ops, monkeypatch, ways, shots, dim, wide):
    """Packed multi-segment tasks with CSR offsets: labels/posteriors/#correct per task (5-way: warp kernel; many-way:
    the register-batch kernel of proto_head_wide.cu and the lane-group kernel of proto_head.cu)."""
    monkeypatch.setenv("AFSL_HEAD_WIDE", str(wide))
    gen = torch.Generator().manual_seed(5)
    tasks = 11
    counts = torch.randint(1, 60, (tasks,), generator=gen)
    off = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
    s = torch.randn(tasks, ways * shots, dim, generator=gen)
    sl = torch.arange(ways).repeat_interleave(shots).expand(tasks, -1)
    q = torch.randn(int(off[-1]), dim, generator=gen)
    ql = torch.randint(0, ways, (int(off[-1]),), generator=gen)
    pred, post, correct, scores = ops.proto_eval(s.cuda(), sl.cuda(), q.cuda(), ql.cuda(), n_way=ways,
                                                 q_offsets=off.cuda(), want_scores=True)
    for i in range(tasks):
        a, b = int(off[i]), int(off[i + 1])
        sc = ohead.l2_scores(q[a:b], ohead.prototypes(s[i], sl[i]))
        po, pr = torch.max(sc, 1)
        assert torch.equal(pred[a:b].cpu().long(), pr)
        close(post[a:b], po)
        close(scores[a:b], sc)
        assert int(correct[i]) == int((pr == ql[a:b]).sum())


def test_l2_normalize(ops):
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(40, 256, generator=gen)
    x[3] = 0                                   # clamped row
    gy = torch.randn(40, 256, generator=gen)
    xg = x.cuda().requires_grad_(True)
    y = ops.l2_normalize(xg)
    y.backward(gy.cuda())
    xc = x.clone().requires_grad_(True)
    yc = torch.nn.functional.normalize(xc, p=2.0, dim=1, eps=1e-12)
    yc.backward(gy)
    close(y, yc)
    mask = torch.ones(40, dtype=torch.bool)
    mask[3] = False
    close(xg.grad[mask.cuda()], xc.grad[mask])


# ------------------------------------------------------------------ CPL
@pytest.mark.parametrize("path", ["warp", "warp_recompute", "cta"])
@pytest.mark.parametrize("name", golden_names("cpl_"))
def test_cpl_vs_reference(ops, monkeypatch, name, path):
    # warp: the forward hands its similarity matrix to the backward; warp_recompute: the backward recomputes it
    monkeypatch.setenv("AFSL_CPL_WARP", "0" if path == "cta" else "1")
    monkeypatch.setenv("AFSL_CPL_SAVE", "0" if path == "warp_recompute" else "1")
    g = load_golden(name)
    p = dev(g["prototypes"]).requires_grad_(True)
    q = dev(g["queries"]).requires_grad_(True)
    labels = dev(g["labels"])
    keep = t(g["keep"])
    loss = ops.cpl_loss(p, q, labels, float(g["temperature"]), keep=keep)
    loss.backward()
    close(loss, t(g["loss"]))
    close(p.grad, t(g["d_prototypes"]))
    close(q.grad, t(g["d_queries"]))
    per_class = int((t(g["labels"]) == 0).sum())
    if int(g["m"]) >= per_class:               # deterministic case needs no mask
        p2 = dev(g["prototypes"]).requires_grad_(True)
        q2 = dev(g["queries"]).requires_grad_(True)
        loss2 = ops.cpl_loss(p2, q2, labels, float(g["temperature"]))
        loss2.backward()
        close(loss2, t(g["loss"]))
        close(q2.grad, t(g["d_queries"]))


@pytest.mark.parametrize("path", ["warp", "warp_recompute", "cta"])
@pytest.mark.parametrize("ways,per,dim,m", [(5, 6, 256, 3), (5, 5, 256, 5), (5, 20, 64, 5), (5, 9, 128, 2), (4, 7, 64, 3)])
def test_cpl_batched_vs_oracle(ops, monkeypatch, ways, per, dim, m, path):
    """E episodes at once == the closed-form oracle episode by episode, sampled-negative masks included,
    through both kernel families (one warp per episode for 5-way, with the forward's similarities saved for the backward
    or recomputed by it; one CTA per episode for any shape)."""
    monkeypatch.setenv("AFSL_CPL_WARP", "0" if path == "cta" else "1")
    monkeypatch.setenv("AFSL_CPL_SAVE", "0" if path == "warp_recompute" else "1")
    e, temp = 19, 2.6981
    gen = torch.Generator().manual_seed(77 + ways * per)
    p = torch.randn(e, ways, dim, generator=gen)
    q = torch.randn(e, ways * per, dim, generator=gen)
    labels = torch.stack([torch.arange(ways).repeat_interleave(per)[torch.randperm(ways * per, generator=gen)] for _ in range(e)])
    torch.manual_seed(31)
    keep = torch.stack([ohead.cpl_draw_keep(labels[i], m) for i in range(e)])
    wl = torch.rand(e, generator=gen) + 0.5
    pg, qg = p.cuda().requires_grad_(True), q.cuda().requires_grad_(True)
    loss = ops.cpl_loss(pg, qg, labels.cuda(), temp, keep=None if m >= per else keep)
    (loss * wl.cuda()).sum().backward()
    for i in range(e):
        pc, qc = p[i].clone().requires_grad_(True), q[i].clone().requires_grad_(True)
        lo = ohead.cpl_loss_closed(pc, qc, labels[i], keep[i], temp)
        (lo * wl[i]).backward()
        close(loss[i], lo)
        close(pg.grad[i], pc.grad)
        close(qg.grad[i], qc.grad)


# ------------------------------------------------------------------ angular (restated-oracle parity: PML unpinned)
def _pml():
    """pytorch_metric_learning when the box has it (runtime probe; it is in neither the reference's
    requirements.txt nor this image, so normally None -> the angular results are 'restated-oracle parity')."""
    try:
        import pytorch_metric_learning
        from pytorch_metric_learning import losses, miners
        return pytorch_metric_learning, losses, miners
    except ImportError:
        return None


def test_angular_oracle_provenance(ops):
    """Says, in the test log, which oracle pins the angular loss on this box, and - when the real package is importable -
    checks the restatement (oracle/angular.py) and the kernel against pytorch_metric_learning itself, driven exactly like
    the reference wrapper does (loops/loss.py:63-97)."""
    from oracle import angular as oang
    pml = _pml()
    if pml is None:
        print("pytorch_metric_learning NOT importable here: angular parity is restated-oracle parity (oracle/angular.py)")
        pytest.skip("pytorch_metric_learning not installed: restated-oracle parity only")
    mod, losses, miners = pml
    print("pytorch_metric_learning", mod.__version__, "importable: angular parity pinned to the real package")
    gen = torch.Generator().manual_seed(123)
    ways, per, dim = 5, 5, 64
    for angle in (0.0, 15.0):
        for anchors in (True, False):
            protos = torch.nn.functional.normalize(torch.randn(ways, dim, generator=gen), dim=-1)
            queries = torch.nn.functional.normalize(torch.randn(ways * per, dim, generator=gen) + 0.5 * protos.repeat_interleave(per, 0), dim=-1)
            labels = torch.arange(ways).repeat_interleave(per)
            loss_fn, miner = losses.AngularLoss(), miners.AngularMiner(angle=angle)
            pc, qc = protos.clone().requires_grad_(True), queries.clone().requires_grad_(True)
            proto_labels = torch.arange(ways)
            if anchors:                                                            # loops/loss.py:68-83
                a, pp, nn_ = miner(pc, proto_labels, qc, labels)
                emb = pc[a]
                ref = torch.cat([qc[pp], qc[nn_]])
                ref_labels = torch.cat([labels[pp], labels[nn_]])
                want = loss_fn(emb, proto_labels[a], ref_emb=ref, ref_labels=ref_labels)
            else:                                                                  # loops/loss.py:84-96
                x = torch.cat([pc, qc])
                y = torch.cat([proto_labels, labels])
                want = loss_fn(x, y, miner(x, y))
            want.backward()
            got_oracle = oang.angular_loss_class(protos, queries, labels, angle, anchors)
            close(got_oracle, want.detach(), rtol=1e-6)
            pg, qg = protos.cuda().requires_grad_(True), queries.cuda().requires_grad_(True)
            got = ops.angular_loss(pg, qg, labels.cuda(), angle, 40.0, anchors, False)
            got.backward()
            close(got, want.detach(), rtol=1e-5)
            close(pg.grad, pc.grad, rtol=2e-5)
            close(qg.grad, qc.grad, rtol=2e-5)


@pytest.mark.parametrize("anchors", [True, False])
@pytest.mark.parametrize("angle", [0.0, 15.0, 30.0])
@pytest.mark.parametrize("unit_protos", [True, False])
@pytest.mark.parametrize("path", ["tc", "warp", "cta"])
def test_angular_vs_restated_oracle(ops, monkeypatch, anchors, angle, unit_protos, path):
    from oracle import angular as oang
    if path == "tc" and not anchors:
        pytest.skip("the tensor-core kernel covers the prototypes-as-anchors branch")
    monkeypatch.setenv("AFSL_ANGULAR_TC", "1" if path == "tc" else "0")
    monkeypatch.setenv("AFSL_ANGULAR_WARP", "1" if path == "warp" else "0")
    e, ways, per, dim = (1203 if path == "tc" else 6), 5, 5, 64
    gen = torch.Generator().manual_seed(int(angle) * 10 + int(anchors) + 100 * int(unit_protos))
    protos = torch.randn(e, ways, dim, generator=gen)
    if unit_protos:
        protos = torch.nn.functional.normalize(protos, dim=-1)
    queries = torch.nn.functional.normalize(torch.randn(e, ways * per, dim, generator=gen) + 0.5 * protos.repeat_interleave(per, 1), dim=-1)
    labels = torch.arange(ways).repeat_interleave(per).expand(e, -1).contiguous()
    wl = torch.rand(e, generator=gen) + 0.5
    pg, qg = protos.cuda().requires_grad_(True), queries.cuda().requires_grad_(True)
    loss = ops.angular_loss(pg, qg, labels.cuda(), angle, 40.0, anchors, False)
    (loss * wl.cuda()).sum().backward()
    # the tensor-core kernel runs 1203 episodes (several tiles per CTA, a partial last tile); the oracle checks a sample
    for i in (range(e) if e <= 6 else [0, 1, 2, 3, 4, 5, 591, 592, 593, 1199, 1200, 1201, 1202]):
        pc, qc = protos[i].clone().requires_grad_(True), queries[i].clone().requires_grad_(True)
        lo = oang.angular_loss_class(pc, qc, labels[i], angle, anchors)
        (lo * wl[i]).backward()
        close(loss[i], lo, rtol=1e-5)
        zero = lambda g, like: torch.zeros_like(like) if g is None else g      # nothing mined -> constant zero loss
        close(pg.grad[i], zero(pc.grad, pc), rtol=2e-5)
        close(qg.grad[i], zero(qc.grad, qc), rtol=2e-5)


@pytest.mark.parametrize("ways,nq,angle,normalize_ref", [(5, 25, 0.0, False), (5, 25, 20.0, False), (3, 12, 0.0, True),
                                                         (8, 23, 10.0, False), (4, 27, 0.0, False), (5, 25, 35.0, True)])
def test_angular_tensor_core_vs_fp32_kernels(ops, monkeypatch, ways, nq, angle, normalize_ref):
    """The tcgen05 kernel against the warp-per-episode fp32 kernel (itself checked against the oracle above) on every
    episode of a multi-tile launch: unbalanced random labels, non-unit prototypes AND queries (the reference rows enter
    the loss un-normalised unless normalize_ref), several (W, Nq); and against the oracle on a sample."""
    from oracle import angular as oang
    e, dim = 777, 64
    gen = torch.Generator().manual_seed(ways * 1000 + nq * 10 + int(angle))
    protos = torch.randn(e, ways, dim, generator=gen) * (0.5 + torch.rand(e, ways, 1, generator=gen))
    labels = torch.randint(0, ways, (e, nq), generator=gen)
    labels[:, :ways] = torch.arange(ways)                      # every class occurs (loops/loss.py:65 asserts it)
    queries = torch.randn(e, nq, dim, generator=gen) + 0.7 * torch.gather(protos, 1, labels.unsqueeze(-1).expand(-1, -1, dim))
    queries = queries * (0.8 + 0.4 * torch.rand(e, nq, 1, generator=gen))
    wl = torch.rand(e, generator=gen) + 0.5
    got = {}
    for path in ("tc", "warp"):
        monkeypatch.setenv("AFSL_ANGULAR_TC", "1" if path == "tc" else "0")
        monkeypatch.setenv("AFSL_ANGULAR_WARP", "1")
        pg, qg = protos.cuda().requires_grad_(True), queries.cuda().requires_grad_(True)
        loss = ops.angular_loss(pg, qg, labels.cuda(), angle, 40.0, True, normalize_ref)
        (loss * wl.cuda()).sum().backward()
        got[path] = (loss.detach().cpu(), pg.grad.cpu(), qg.grad.cpu())
    close(got["tc"][0], got["warp"][0], rtol=1e-5)
    close(got["tc"][1], got["warp"][1], rtol=2e-5)
    close(got["tc"][2], got["warp"][2], rtol=2e-5)
    for i in (0, 333, 776):
        pc, qc = protos[i].clone().requires_grad_(True), queries[i].clone().requires_grad_(True)
        lo = oang.angular_loss_class(pc, qc, labels[i], angle, True, normalize_ref=normalize_ref)
        (lo * wl[i]).backward()
        close(got["tc"][0][i], lo, rtol=1e-5)
        zero = lambda g, like: torch.zeros_like(like) if g is None else g
        close(got["tc"][1][i], zero(pc.grad, pc), rtol=2e-5)
        close(got["tc"][2][i], zero(qc.grad, qc), rtol=2e-5)


def test_angular_tensor_core_full_size_properties(ops, monkeypatch):
    """16387 episodes (every CTA runs ~28 tiles, the last tile is partial) through the tcgen05 kernel: repeated launches are
    bit-identical (no race between the producer / epilogue / issuer roles), and permuting the episodes permutes losses and
    gradients bit for bit (an episode's result does not depend on its tile, its slot in the tile or its neighbours)."""
    monkeypatch.setenv("AFSL_ANGULAR_TC", "1")
    e, ways, per, dim = 16387, 5, 5, 64
    gen = torch.Generator().manual_seed(2026)
    protos = torch.nn.functional.normalize(torch.randn(e, ways, dim, generator=gen), dim=-1).cuda()
    queries = torch.nn.functional.normalize(torch.randn(e, ways * per, dim, generator=gen), dim=-1).cuda()
    labels = torch.arange(ways).repeat_interleave(per).expand(e, -1).contiguous().cuda()
    wl = (torch.rand(e, generator=gen) + 0.5).cuda()

    def run(p, q, w):
        pg, qg = p.clone().requires_grad_(True), q.clone().requires_grad_(True)
        loss = ops.angular_loss(pg, qg, labels, 0.0, 40.0, True, False)
        (loss * w).sum().backward()
        return loss.detach(), pg.grad, qg.grad

    first = run(protos, queries, wl)
    for _ in range(2):
        again = run(protos, queries, wl)
        for a, b in zip(first, again):
            assert torch.equal(a, b)
    perm = torch.randperm(e, generator=gen).cuda()
    shuffled = run(protos[perm].contiguous(), queries[perm].contiguous(), wl[perm].contiguous())
    for a, b in zip(first, shuffled):
        assert torch.equal(a[perm], b)
    assert torch.isfinite(first[0]).all() and float(first[0].min()) > 0


# ------------------------------------------------------------------ SpecAugment
def warp_atol(x):
    """Bound for the in-kernel spline.  torch evaluates u**2, u**3 with a 1-ulp vectorised powf that
    cannot be reproduced bit for bit, so the normalised source coordinate may differ by <= 2 ulp(1) =
    2.4e-7, i.e. (T-1)/2 * 2.4e-7 pixels; bilinear output then moves by at most that times the largest
    neighbour difference (<= 2 max|x|).  close() multiplies `scale` by rtol=0 -> use it as atol via rtol arg."""
    t_len = x.shape[-1]
    return 2.4e-7 * (t_len - 1) / 2 * 2 * float(x.abs().max())


@pytest.mark.parametrize("kernel", ["tile", "rows"])
@pytest.mark.parametrize("name", golden_names("specaug_"))
def test_specaug_vs_reference(ops, monkeypatch, name, kernel):
    monkeypatch.setenv("AFSL_SPECAUG_TILE", "1" if kernel == "tile" else "0")
    g = load_golden(name)
    x = dev(g["x"])
    n = x.shape[0]
    tm = t(g["time_masks"]).view(1, -1, 2)
    fm = t(g["freq_masks"]).view(1, -1, 2)
    value = float(g["cfg_mask_value"])
    # (a) spline evaluated by the reference on the host -> only the bilinear blend is ours
    v = ops.specaug_views(x, t(g["warp_p"]), t(g["warp_d"]), tm, fm, value, set_size=n, src_x=t(g["src_x"]))
    assert torch.equal(v[0].cpu(), t(g["original"]))
    assert torch.equal(v[2].cpu(), t(g["time_masked"]))          # masks: bit-exact
    assert torch.equal(v[3].cpu(), t(g["freq_masked"]))
    close(v[1], t(g["warped"]), rtol=1e-6)
    # (b) spline evaluated in the kernel from the host-drawn control points
    v2 = ops.specaug_views(x, t(g["warp_p"]), t(g["warp_d"]), tm, fm, value, set_size=n)
    close(v2[1], t(g["warped"]), rtol=0, scale=warp_atol(x))
    assert torch.equal(v2[2], v[2]) and torch.equal(v2[3], v[3])


@pytest.mark.parametrize("kernel", ["tile", "rows"])
@pytest.mark.parametrize("t_len", [157, 126, 64, 101])
def test_specaug_batched_sets_vs_oracle(ops, monkeypatch, t_len, kernel):
    """Several 25-sample sets in one launch, each with its own masks, vs the oracle set by set; both kernels
    (128-bit tile kernel, warp-per-row kernel) and time lengths with every alignment of T modulo 4."""
    monkeypatch.setenv("AFSL_SPECAUG_TILE", "1" if kernel == "tile" else "0")
    cfg = {"specaug_params": {"mask_param": 16, "W": 22, "num_mask": 2, "mask_value": 0.25, "p": 0.282}}
    sets, n = 5, 25
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(sets * n, 1, 128, t_len, generator=gen)
    torch.manual_seed(123)
    np.random.seed(123)
    params = [ospec.draw_params(n, t_len, cfg) for _ in range(sets)]
    wp = torch.cat([p.warp_p for p in params])
    wd = torch.cat([p.warp_d for p in params])
    tm = torch.tensor([p.time_masks for p in params])
    fm = torch.tensor([p.freq_masks for p in params])
    v = ops.specaug_views(x.cuda(), wp, wd, tm, fm, 0.25, set_size=n)
    for i, p in enumerate(params):
        ref = ospec.apply(x[i * n:(i + 1) * n], p, 0.25)
        sl = slice(i * n, (i + 1) * n)
        assert torch.equal(v[0, sl].cpu(), ref[0])
        assert torch.equal(v[2, sl].cpu(), ref[2])
        assert torch.equal(v[3, sl].cpu(), ref[3])
        close(v[1, sl], ref[1], rtol=0, scale=warp_atol(x))


@pytest.mark.parametrize("kernel", ["tile", "rows"])
def test_specaug_ragged_sets_vs_oracle(ops, monkeypatch, kernel):
    """Sets of different sizes packed back to back (the query segments of multi-segment tasks: one
    apply_augmentations call per task), named by set_ids, vs the oracle set by set."""
    from afsl_b200.utils.augmentations import SpecAugment
    monkeypatch.setenv("AFSL_SPECAUG_TILE", "1" if kernel == "tile" else "0")
    cfg = {"specaug_params": {"use": True, "mask_param": 16, "W": 22, "num_mask": 1, "mask_value": 0.0, "p": 0.282}}
    sizes, t_len = [3, 7, 1, 12, 5], 157
    gen = torch.Generator().manual_seed(17)
    x = torch.randn(sum(sizes), 1, 128, t_len, generator=gen)
    torch.manual_seed(5); np.random.seed(5)
    params = SpecAugment(cfg).draw_ragged(sizes, t_len, replay_reference_rng=True)
    torch.manual_seed(5); np.random.seed(5)
    ref_params = [ospec.draw_params(n, t_len, cfg) for n in sizes]          # the same draws, one oracle call per set
    v = SpecAugment(cfg).apply_batch(x.cuda(), params)
    a = 0
    for n, rp in zip(sizes, ref_params):
        ref = ospec.apply(x[a:a + n], rp, 0.0)
        sl = slice(a, a + n)
        assert torch.equal(v[0, sl].cpu(), ref[0])
        assert torch.equal(v[2, sl].cpu(), ref[2])
        assert torch.equal(v[3, sl].cpu(), ref[3])
        close(v[1, sl], ref[1], rtol=0, scale=warp_atol(x))
        a += n


def test_specaug_whole_sample_ctas_match_row_kernel(ops, monkeypatch):
    """Large launches switch the tile kernel to one CTA per sample (column tables set up once per sample);
    its four views must equal, bit for bit, those of the warp-per-row kernel that the oracle tests pin."""
    sets, n, t_len = 240, 25, 44
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(sets * n, 1, 128, t_len, generator=gen).cuda()
    wp = torch.randint(8, t_len - 8, (sets * n,), generator=gen)
    wd = torch.randint(-8, 8, (sets * n,), generator=gen)
    tm = torch.stack([torch.randint(0, t_len - 12, (sets, 2), generator=gen), torch.randint(1, 12, (sets, 2), generator=gen)], -1)
    fm = torch.stack([torch.randint(0, 110, (sets, 2), generator=gen), torch.randint(1, 16, (sets, 2), generator=gen)], -1)
    monkeypatch.setenv("AFSL_SPECAUG_TILE", "1")
    a = ops.specaug_views(x, wp, wd, tm, fm, -1.5, set_size=n)
    monkeypatch.setenv("AFSL_SPECAUG_TILE", "0")
    b = ops.specaug_views(x, wp, wd, tm, fm, -1.5, set_size=n)
    assert torch.equal(a, b)


# ------------------------------------------------------------------ majority vote
def test_vote_vs_reference(ops):
    g = load_golden("vote_cases")
    off = torch.from_numpy(g["offsets"])
    for k, strat in enumerate(("", "min_label", "max_posterior")):
        correct, clips = ops.eval_vote(dev(g["pred"]), dev(g["clip_ids"]), dev(g["labels"]), dev(g["posterior"]), off, strat)
        acc = correct.cpu().double() / clips.cpu().double()
        assert np.array_equal(acc.numpy(), g["accuracy"][:, k]), strat


def test_vote_random_vs_oracle(ops):
    rng = np.random.RandomState(1)
    preds, ids, labs, posts, offs = [], [], [], [], [0]
    for _ in range(300):
        clips = rng.randint(1, 30)
        seg = rng.randint(1, 12, size=clips)
        cid = np.repeat(rng.permutation(clips) * 3, seg)         # ids need not be 0..C-1 or sorted
        if rng.rand() < 0.3:
            cid = cid[rng.permutation(cid.size)]
        true = rng.randint(0, 5, size=clips * 3)[cid]
        pred = rng.randint(0, 4, size=cid.size)
        post = np.round(-rng.rand(cid.size) * 4, 1).astype(np.float32)
        preds.append(pred); ids.append(cid); labs.append(true); posts.append(post); offs.append(offs[-1] + cid.size)
    P, I, L, Q = map(np.concatenate, (preds, ids, labs, posts))
    for strat in ("", "min_label", "max_posterior"):
        correct, clips = ops.eval_vote(dev(P), dev(I), dev(L), dev(Q), torch.tensor(offs), strat)
        got = (correct.cpu().double() / clips.cpu().double()).numpy()
        want = np.array([ovote.majority_vote_accuracy(preds[i], ids[i], labs[i], posts[i], strat) for i in range(300)])
        assert np.array_equal(got, want), strat


# ------------------------------------------------------------------ full-size properties
def test_full_size_properties(ops):
    """BASELINE-scale batch (E = 16384 episodes, 5w5s5q, D = 256): size-independent properties."""
    e, ways, shots, nq, dim = 16384, 5, 5, 5, 256
    gen = torch.Generator(device="cuda").manual_seed(0)
    s = torch.randn(e, ways * shots, dim, device="cuda", generator=gen)
    q = torch.randn(e, ways * nq, dim, device="cuda", generator=gen)
    sl = torch.arange(ways, device="cuda").repeat_interleave(shots).expand(e, -1).contiguous()
    ql = torch.arange(ways, device="cuda").repeat_interleave(nq).expand(e, -1).contiguous()
    loss, protos, correct = ops.proto_head(s, sl, q, ql, n_way=ways)
    # (1) prototypes are the class means: linear in the support set
    close(protos, s.view(e, ways, shots, dim).mean(2), rtol=1e-5)
    # (2) permuting episodes permutes the outputs (no cross-episode leakage), bit-exactly
    perm = torch.randperm(e, device="cuda")
    loss_p, _, correct_p = ops.proto_head(s[perm].contiguous(), sl, q[perm].contiguous(), ql, n_way=ways)
    assert torch.equal(loss_p, loss[perm]) and torch.equal(correct_p, correct[perm])
    # (3) a query placed exactly on its own prototype is classified correctly with posterior 0
    q2 = protos[:, ql[0]].contiguous()
    pred, post, corr, _ = ops.proto_eval(s, sl, q2, ql, n_way=ways)
    assert torch.equal(pred.view(e, -1).long(), ql) and bool((post == 0).all()) and bool((corr == ways * nq).all())
    # (4) translation invariance of distances: shifting support and queries together leaves the loss unchanged
    shift = torch.randn(e, 1, dim, device="cuda", generator=gen)
    loss_s, _, _ = ops.proto_head(s + shift, sl, q + shift, ql, n_way=ways)
    close(loss_s, loss, rtol=1e-4)
    # (5) idempotence / determinism
    loss_again, _, _ = ops.proto_head(s, sl, q, ql, n_way=ways)
    assert torch.equal(loss_again, loss)


# ------------------------------------------------------------------ grouped BatchNorm + ReLU + MaxPool
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("shape,group", [((50, 8, 20, 25), 25), ((12, 64, 42, 52), 4), ((6, 16, 14, 17), 3),
                                          ((10, 64, 4, 5), 5), ((4, 4, 128, 157), 2)])
def test_gbn_relu_pool_vs_torch(ops, shape, group, layout):
    """Fused kernels vs the eager fp32 chain (per-group nn.BatchNorm2d -> ReLU -> MaxPool2d(3)) on the GPU, for
    NCHW activations and for channels-last ones (the layout the encoder runs in)."""
    n, c, h, w = shape
    gen = torch.Generator().manual_seed(n * c + h)
    x = (torch.randn(shape, generator=gen) * 2 + 0.7).cuda()
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(c).cuda()
    ref = torch.nn.BatchNorm2d(c).cuda()
    with torch.no_grad():
        bn.weight.uniform_(-1.5, 1.5); bn.bias.uniform_(-0.5, 0.5)        # negative gammas included
        ref.weight.copy_(bn.weight); ref.bias.copy_(bn.bias)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    cb = (torch.randn(c, generator=gen) * 0.3).cuda().requires_grad_(True)       # convolution bias, folded by the kernel
    cbr = cb.detach().clone().requires_grad_(True)
    y = ops.gbn_relu_pool(xa, bn, group, conv_bias=cb)
    if layout == "nhwc" and y.shape[2] * y.shape[3] > 1:
        assert ops._is_nhwc(y)                       # channels-last in, channels-last out
    yr = torch.cat([torch.nn.functional.max_pool2d(torch.relu(ref(xb[i:i + group] + cbr.view(1, -1, 1, 1))), 3, 3)
                    for i in range(0, n, group)])
    close(y, yr, rtol=1e-5)
    gy = torch.randn(y.shape, generator=gen).cuda()
    y.backward(gy); yr.backward(gy)
    close(xa.grad, xb.grad, rtol=1e-4)
    close(bn.weight.grad, ref.weight.grad, rtol=1e-4)
    close(bn.bias.grad, ref.bias.grad, rtol=1e-4)
    close(bn.running_mean, ref.running_mean, rtol=1e-5)
    close(bn.running_var, ref.running_var, rtol=1e-5)
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == n // group
    close(cb.grad, cbr.grad, rtol=0, scale=1e-4 * float(xb.grad.abs().max()) * h * w)     # exactly 0 vs round-off
    # eval mode: running statistics, gradients included
    bn.eval(); ref.eval()
    for tns in (xa, xb, cb, cbr, bn.weight, bn.bias, ref.weight, ref.bias):
        tns.grad = None
    ye = ops.gbn_relu_pool(xa, bn, group, conv_bias=cb)
    yer = torch.nn.functional.max_pool2d(torch.relu(ref(xb + cbr.view(1, -1, 1, 1))), 3, 3)
    close(ye, yer, rtol=1e-5)
    ye.backward(gy); yer.backward(gy)
    close(xa.grad, xb.grad, rtol=1e-4)
    close(cb.grad, cbr.grad, rtol=1e-4)
    close(bn.weight.grad, ref.weight.grad, rtol=1e-4)
    close(bn.bias.grad, ref.bias.grad, rtol=1e-4)


def test_batched_encoder_fused_vs_per_episode(ops):
    """Hybrid encoder on [E,N,1,128,157] views (fused stages, grouped statistics) == one eager call per (episode, view)."""
    import copy
    import afsl_b200.models.main_modules as mm
    torch.manual_seed(3)
    enc = mm.EncoderModule({"encoder_name": "Hybrid"}, {"Hybrid": {"in_channels": 1, "seq_layers": 1, "seq_type": "RNN",
                           "bidirectional": False, "hidden_channels": 64, "pool_dim": [3, 3], "out_dim": 64}}).cuda()
    for m in enc.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ref = copy.deepcopy(enc)
    enc.train(); ref.train()
    torch.backends.cudnn.allow_tf32 = False
    try:
        views = [torch.randn(2, 5, 1, 128, 157, device="cuda") for _ in range(2)]
        got = enc(views)
        (sum(g.square().sum() for g in got)).backward()
        mm.FUSED_STAGES = False
        want = [[ref([views[v][e]])[0] for e in range(2)] for v in range(2)]
        (sum(w.square().sum() for row in want for w in row)).backward()
    finally:
        mm.FUSED_STAGES = True
        torch.backends.cudnn.allow_tf32 = True
    for v in range(2):
        for e in range(2):
            close(got[v][e], want[v][e], rtol=2e-4)
    params, refs = dict(enc.named_parameters()), dict(ref.named_parameters())
    for k, a in params.items():
        if k.startswith("encoder.conv_encoder") and k.endswith(".0.bias"):
            # a convolution bias feeding batch-statistics BatchNorm has an exactly zero gradient; the eager
            # chain returns round-off noise (sum of d_u), so compare against the weight-gradient scale
            scale = float(refs[k.replace("bias", "weight")].grad.abs().max())
            close(a.grad, refs[k].grad, rtol=0, scale=1e-4 * scale)
        else:
            close(a.grad, refs[k].grad, rtol=2e-3)
    for (k, a), (_, b) in zip(enc.named_buffers(), ref.named_buffers()):
        close(a.float(), b.float(), rtol=1e-4)


# ------------------------------------------------------------------ view fusion (transformer encoder layer)
def _fusion_reference(layer, x, masks):
    """The layer's arithmetic written out with explicit dropout masks (plain torch fp32)."""
    a = layer.self_attn
    d = a.embed_dim
    qkv = x @ a.in_proj_weight.t() + a.in_proj_bias
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    p = torch.softmax(q @ k.transpose(-1, -2) / d ** 0.5, dim=-1)
    if masks[0] is not None:
        p = p * masks[0]
    sa = (p @ v) @ a.out_proj.weight.t() + a.out_proj.bias
    if masks[1] is not None:
        sa = sa * masks[1]
    x1 = layer.norm1(x + sa)
    pre = layer.linear1(x1)
    h = torch.relu(pre)
    if masks[2] is not None:
        h = h * masks[2]
    ff = layer.linear2(h)
    if masks[3] is not None:
        ff = ff * masks[3]
    return layer.norm2(x1 + ff), pre


def test_view_fusion_vs_reference_fixture(ops):
    """Eval-mode and dropout-free train-mode outputs / gradients of the REAL reference SelfAttention module."""
    g = load_golden("modules_fusion")
    layer = torch.nn.TransformerEncoderLayer(64, 1, 256, 0.0, batch_first=True).cuda()
    layer.load_state_dict({k[len("w_encoder_layer."):]: dev(v) for k, v in g.items() if k.startswith("w_encoder_layer.")})
    x = dev(g["x"]).requires_grad_(True)
    y = ops.view_fusion(x, layer).reshape(x.shape[0], -1)
    close(y, t(g["y"]), rtol=1e-5)
    close(y, t(g["y_eval"]), rtol=1e-5)
    y.backward(dev(g["gy"]))
    close(x.grad, t(g["dx"]), rtol=2e-5)
    for k, p in layer.named_parameters():
        close(p.grad, t(g["g_encoder_layer." + k]), rtol=2e-5)


@pytest.mark.parametrize("n,v,with_masks", [(37, 4, False), (37, 4, True), (200, 4, True), (9, 2, True), (5, 8, False), (3, 1, True)])
def test_view_fusion_vs_torch(ops, n, v, with_masks):
    torch.manual_seed(n * 10 + v)
    layer = torch.nn.TransformerEncoderLayer(64, 1, 256, 0.1, batch_first=True).cuda()
    ref = torch.nn.TransformerEncoderLayer(64, 1, 256, 0.1, batch_first=True).cuda()
    ref.load_state_dict(layer.state_dict())
    x = torch.randn(n, v, 64, device="cuda")
    masks = (None, None, None, None)
    if with_masks:
        keep = lambda *s: (torch.rand(*s, device="cuda") >= 0.1).float() / 0.9
        masks = (keep(n, v, v), keep(n, v, 64), keep(n, v, 256), keep(n, v, 64))
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y = ops.view_fusion(xa, layer, masks=masks)
    yr, pre = _fusion_reference(ref, xb, masks)
    close(y, yr, rtol=1e-5)
    # ReLU is discontinuous in its derivative: a hidden unit whose pre-activation is within rounding of 0 may be
    # gated differently by two correct fp32 implementations.  Such samples (expected ~0.3 per run) get no gradient.
    safe = (pre.detach().abs() > 1e-5).all(-1).all(-1)                   # per sample
    gy = torch.randn_like(y) * safe.view(-1, 1, 1)
    y.backward(gy); yr.backward(gy)
    close(xa.grad, xb.grad, rtol=2e-5)
    for (k, a), (_, b) in zip(layer.named_parameters(), ref.named_parameters()):
        close(a.grad, b.grad, rtol=2e-5)


# ------------------------------------------------------------------ fused encoder stage 1 (conv1 + BN + ReLU + pool)
def close_channels(got, ref, rtol, max_outliers=3, scale=None):
    """Per-channel (dim 0) check with the absolute floor rtol * max|ref|: every channel within tolerance, except that up to
    `max_outliers` channels may deviate by a bounded amount (a pooling-winner / ReLU-gate flip, see the test below)."""
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    atol = rtol * float(ref.abs().max() if scale is None else scale)
    err = (got - ref).abs().reshape(got.shape[0], -1).max(1).values
    lim = atol + rtol * ref.abs().reshape(ref.shape[0], -1).max(1).values
    bad = torch.nonzero(err > lim).flatten().tolist()
    assert len(bad) <= max_outliers, (bad, err[bad].tolist(), atol)
    record_observed(float(err.max()), 200 * atol)
    assert float(err.max()) <= 200 * atol, (float(err.max()), atol)


@pytest.mark.parametrize("n,group,h,w", [(10, 5, 128, 157), (6, 3, 128, 126), (4, 4, 33, 40), (50, 25, 128, 157)])
def test_stage1_fused_vs_torch(ops, n, group, h, w):
    """Fused stage 1 vs the eager chain conv2d -> per-group BatchNorm2d -> ReLU -> MaxPool2d(3) (fp32, TF32 off)."""
    gen = torch.Generator().manual_seed(n + h)
    # Module init / BatchNorm scales below use the global generator: fixed so the case is reproducible.  Gradients are compared
    # per channel and up to three of the 64 channels per check may deviate: at the largest case there are 7 M pooling windows, and where two
    # window elements (or a ReLU gate) are within ~2 ulp the two fp32 convolution sums (cuDNN's and this kernel's fma chain)
    # pick different winners - one such flip in a channel with a small BatchNorm scale moves that channel's gradients by more
    # than the tolerance (about 40 % of seeds) while every other channel agrees to ~2e-6.  tools/dbg_stage1_case.py scans seeds.
    torch.manual_seed(1000 + n + w)
    x = (torch.randn(n, 1, h, w, generator=gen) * 1.3 + 0.2).cuda()
    conv, bn = torch.nn.Conv2d(1, 64, 3, padding=1).cuda(), torch.nn.BatchNorm2d(64).cuda()
    conv_r, bn_r = torch.nn.Conv2d(1, 64, 3, padding=1).cuda(), torch.nn.BatchNorm2d(64).cuda()
    with torch.no_grad():
        bn.weight.uniform_(-1.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    conv_r.load_state_dict(conv.state_dict()); bn_r.load_state_dict(bn.state_dict())
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        import copy
        with torch.no_grad():       # no-gradient variant (no argmax codes, affine applied to the window's extreme u): same bits
            y_ng = ops.stage1_conv_bn_relu_pool(x, conv, copy.deepcopy(bn), group)
        y = ops.stage1_conv_bn_relu_pool(x, conv, bn, group)
        assert torch.equal(y_ng, y)
        yr = torch.cat([torch.nn.functional.max_pool2d(torch.relu(bn_r(conv_r(x[i:i + group]))), 3, 3) for i in range(0, n, group)])
        close(y, yr, rtol=2e-5)
        gy = torch.randn(y.shape, generator=gen).cuda()
        y.backward(gy); yr.backward(gy)
        wscale = float(conv_r.weight.grad.abs().max())
        close_channels(conv.weight.grad, conv_r.weight.grad, rtol=1e-4)
        close(conv.bias.grad, conv_r.bias.grad, rtol=0, scale=1e-4 * wscale)          # exactly 0 vs round-off
        close_channels(bn.weight.grad, bn_r.weight.grad, rtol=1e-4)
        close_channels(bn.bias.grad, bn_r.bias.grad, rtol=1e-4)
        close(bn.running_mean, bn_r.running_mean, rtol=1e-5)
        close(bn.running_var, bn_r.running_var, rtol=1e-5)
        assert int(bn.num_batches_tracked) == int(bn_r.num_batches_tracked)
        # eval mode (running statistics), gradients included
        for mod in (conv, bn, conv_r, bn_r):
            mod.zero_grad(); mod.eval()
        ye = ops.stage1_conv_bn_relu_pool(x, conv, bn, group)
        with torch.no_grad():
            assert torch.equal(ops.stage1_conv_bn_relu_pool(x, conv, bn, group), ye)
        yer = torch.nn.functional.max_pool2d(torch.relu(bn_r(conv_r(x))), 3, 3)
        close(ye, yer, rtol=2e-5)
        ye.backward(gy); yer.backward(gy)
        close_channels(conv.weight.grad, conv_r.weight.grad, rtol=1e-4)
        close_channels(conv.bias.grad, conv_r.bias.grad, rtol=1e-4)
        close_channels(bn.weight.grad, bn_r.weight.grad, rtol=1e-4)
        close_channels(bn.bias.grad, bn_r.bias.grad, rtol=1e-4)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


# ------------------------------------------------------------------ batched episode runner: CUDA-graph replay == eager
def test_runner_cuda_graph_matches_eager(ops):
    """EpisodeRunner.train_step replayed from a CUDA graph gives the losses and parameter updates of the eager
    step on the same batches and host-drawn randomness (dropout off: its device RNG stream differs under capture)."""
    import copy
    import random
    import bench
    from afsl_b200.episodes import EpisodeRunner, synthetic_batch
    cfg = copy.deepcopy(bench.EXPERIMENT_CONFIG)
    mcfg = copy.deepcopy(bench.MODEL_CONFIG)
    mcfg["Attention"]["dropout"] = 0.0
    dev = torch.device("cuda", 0)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        results = []
        for graph in (False, True):
            torch.manual_seed(7); np.random.seed(7); random.seed(7)
            from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, SelfAttention
            from afsl_b200.models.prototypical import ContrastivePrototypicalNetworks
            model = ContrastivePrototypicalNetworks(EncoderModule(cfg, mcfg), SelfAttention(mcfg), ProjectionHead(mcfg)).to(dev)
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
            runner = EpisodeRunner(model, cfg, opt, replay_reference_rng=False, use_cuda_graph=graph)
            torch.manual_seed(11); np.random.seed(11); random.seed(11)
            losses = []
            for step in range(3):
                batch = synthetic_batch(3, 5, 5, 5, 157, seed=100 + step)
                losses.append(runner.train_step(batch)["loss"].clone())
            results.append((torch.stack(losses), [p.detach().clone() for p in model.parameters()]))
            if graph:
                assert runner.launches_per_replay > 10
        (l0, p0), (l1, p1) = results
        # first step: same parameters, same batch, same randomness -> same losses.  Later steps go through Adam, whose
        # normalised update turns last-bit differences of near-zero gradients (cuDNN picks its algorithms afresh under
        # capture) into +-lr parameter differences, so they are only checked to stay within a few lr of each other.
        close(l1[0], l0[0], rtol=1e-5)
        close(l1[1:], l0[1:], rtol=1e-2)
        for a, b in zip(p1, p0):
            close(a, b, rtol=0, scale=3 * 2 * cfg["lr"] + 1e-6)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


# ------------------------------------------------------------------ training / validation loops vs the reference's own loops
class _FakeDataset:
    """Same construction as tests/golden/make_golden.py::FakeDataset (single-segment clips)."""

    def __init__(self, seed, classes=6, per_class=12, t_len=157):
        import pandas as pd
        g = torch.Generator().manual_seed(seed)
        self.items = torch.randn(classes * per_class, 1, 1, 128, t_len, generator=g)
        names = [f"c{i}" for i in range(classes)]
        self.class_to_label = {n: i for i, n in enumerate(names)}
        self.data_df = pd.DataFrame({"label": [names[i // per_class] for i in range(classes * per_class)],
                                     "index_column": list(range(classes * per_class))})
        self.multi_segm, self.input_type, self.specaug_use, self.waveaug_use = False, "spec", False, False
        self.experiment_config = {"specaug_params": {"use": False}}

    def __getitem__(self, i):
        return self.items[i], 0


def test_training_epoch_and_validation_vs_reference_loops(ops):
    """loops.training_epoch (two episodes, one optimizer step each) and loops.evaluate_single_segment on the GPU
    reproduce what the REFERENCE's loops/loops.py produced on the same fake dataset, seeds and initial weights
    (tests/golden/epoch_cnn_plain.npz: Conv4 ProtoNet, no views - BASELINE config 1)."""
    import random
    from afsl_b200.loops.loops import evaluate_single_segment, training_epoch
    from afsl_b200.loops.loss import FSL_Loss
    from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, StandardCNN
    from afsl_b200.models.prototypical import ContrastivePrototypicalNetworksWithoutAttention
    g = load_golden("epoch_cnn_plain")
    ds = _FakeDataset(int(g["items_seed"]))
    assert abs(float(ds.items.double().sum()) - float(g["items_checksum"])) < 1e-6
    mc = {"Projection": {"input_dim": 64, "hidden_dim": 32, "output_dim": 64}}
    backbone = EncoderModule({"encoder_name": "CNN"}, {}, encoder=StandardCNN(1, (1, 1, 128, 157), 64, [3, 3], 64))
    net = ContrastivePrototypicalNetworksWithoutAttention(backbone, ProjectionHead(mc))
    net.load_state_dict({k[len("init_"):]: t(v) for k, v in g.items() if k.startswith("init_")})
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    net = net.cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    init = {k[len("init_"):]: t(v) for k, v in g.items() if k.startswith("init_")}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        random.seed(1234); np.random.seed(1234); torch.manual_seed(1234)
        first = training_epoch(net, ds, opt, 1, "cuda", FSL_Loss(), None, 0.0, False, False, 5, 5, 5, None, False, False)
        # restart from the initial weights for the two-episode epoch of the fixture
        net.load_state_dict(init)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3)
        random.seed(1234); np.random.seed(1234); torch.manual_seed(1234)
        msg = training_epoch(net, ds, opt, 2, "cuda", FSL_Loss(), None, 0.0, False, False, 5, 5, 5, None, False, False)
        random.seed(99)
        val_mean, val_std = evaluate_single_segment(net, ds, 3, "cuda", 5, 5, 5, None, False)
        after = {k: v.detach().clone() for k, v in net.state_dict().items()}
        # the same validation with the REFERENCE's post-training weights: identical weights -> identical accuracies
        net.load_state_dict({k[len("after_"):]: t(v) for k, v in g.items() if k.startswith("after_")})
        random.seed(99)
        ref_mean, ref_std = evaluate_single_segment(net, ds, 3, "cuda", 5, 5, 5, None, False)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    # the first episode does not go through the optimizer: tight.  The second one follows an Adam step, whose
    # normalised update turns last-bit differences of near-zero gradients into +-lr parameter differences.
    g1 = load_golden("epoch_cnn_first")
    assert abs(first["loss"] - float(g1["loss"])) <= 1e-5 * abs(float(g1["loss"]))
    assert abs(msg["loss"] - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    assert abs(msg["fsl_loss"] - float(g["fsl_loss"])) <= 1e-3 * abs(float(g["fsl_loss"]))
    assert np.isnan(msg["cpl_loss"])
    assert abs(ref_mean - float(g["val_mean"])) < 1e-9 and abs(ref_std - float(g["val_std"])) < 1e-9
    # with the weights trained HERE (two Adam steps away from the reference's by +-lr in entries whose near-zero gradient
    # changes sign in the last bit, see below) a near-tie among the 75 validation queries may fall the other way
    print(f"validation accuracy with the weights trained here: {val_mean:.6f} (reference's run: {float(g['val_mean']):.6f})")
    assert abs(val_mean - float(g["val_mean"])) <= 3.0 / 75 + 1e-9
    for k, v in g.items():
        if k.startswith("after_") and "num_batches_tracked" not in k:
            # two Adam steps: +-lr per step where a near-zero gradient's sign differs in the last bit
            close(after[k[len("after_"):]], t(v), rtol=0, scale=2 * 2 * 1e-3 + 1e-6)


# ------------------------------------------------------------------ one whole training episode vs the CPU oracle
@pytest.mark.parametrize("loss_kind", ["cpl", "angular", "concat_cpl"])
def test_runner_episode_vs_oracle_episode(ops, loss_kind):
    """EpisodeRunner.train_step on ONE episode with replay_reference_rng=True == oracle/episode.py::train_step (the
    restatement of loops/loops.py:26-61 pinned to the reference) on the same seeds: SpecAugment draws, view shuffle,
    CPL negatives, losses and the parameter update.  Config 2 (CPL) and config 3 (angular, NSynth-shaped)."""
    import copy
    import random
    import bench
    from afsl_b200.episodes import EpisodeRunner, synthetic_batch
    from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, SelfAttention
    from afsl_b200.models.prototypical import ContrastivePrototypicalNetworks, ContrastivePrototypicalNetworksWithoutAttention
    from oracle import episode as oep
    from oracle import modules as om
    cfg = copy.deepcopy(bench.EXPERIMENT_CONFIG)
    mcfg = copy.deepcopy(bench.MODEL_CONFIG)
    mcfg["Attention"]["dropout"] = 0.0
    t_len = 157
    if loss_kind == "angular":
        cfg["loss"]["cpl"]["use"] = False
        cfg["loss"]["angular"] = {"use": True, "angle": 0, "prototypes_as_anchors": True}
        cfg["loss"]["l_param"] = 1.0
        cfg["specaug_params"] = {"use": True, "mask_param": 9, "W": 36, "num_mask": 1, "mask_value": 0, "p": 0.42157}
        mcfg["Projection"] = {"input_dim": 256, "hidden_dim": 64, "output_dim": 64}
        t_len = 126
    else:
        cfg["loss"]["cpl"]["m_param"] = 3          # sampled negatives: exercises the host-drawn keep mask
    if loss_kind == "concat_cpl":                  # no view fusion: views stacked along the sample axis, labels repeated
        cfg["use_attention"] = False
        mcfg["Projection"] = {"input_dim": 64, "hidden_dim": 128, "output_dim": 64}
    pc = mcfg["Projection"]
    torch.manual_seed(42)
    if loss_kind == "concat_cpl":
        ref = om.ConcatViewsNet(om.ViewEncoder(om.build_encoder("Hybrid", t_len)),
                                om.Projection(pc["input_dim"], pc["hidden_dim"], pc["output_dim"]))
        net = ContrastivePrototypicalNetworksWithoutAttention(EncoderModule(cfg, mcfg), ProjectionHead(mcfg))
    else:
        ref = om.FusedViewsNet(om.ViewEncoder(om.build_encoder("Hybrid", t_len)), om.ViewFusion(64, 1, 256, 0.0),
                               om.Projection(pc["input_dim"], pc["hidden_dim"], pc["output_dim"]))
        net = ContrastivePrototypicalNetworks(EncoderModule(cfg, mcfg), SelfAttention(mcfg), ProjectionHead(mcfg))
    net.load_state_dict(ref.state_dict())
    for m in list(ref.modules()) + list(net.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    net = net.cuda()
    lr = 1e-3
    ropt, nopt = torch.optim.Adam(ref.parameters(), lr=lr), torch.optim.Adam(net.parameters(), lr=lr)
    batch = synthetic_batch(1, 5, 5, 5, t_len, seed=321)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(9); np.random.seed(9); random.seed(9)
        r_total, r_fsl, r_extra = oep.train_step(ref, ropt, batch.support[0], batch.support_labels[0], batch.query[0],
                                                 batch.query_labels[0], cfg)
        torch.manual_seed(9); np.random.seed(9); random.seed(9)
        out = EpisodeRunner(net, cfg, nopt, replay_reference_rng=True).train_step(batch)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert abs(float(out["fsl_loss"][0]) - r_fsl) <= 2e-5 * abs(r_fsl)
    assert abs(float(out["cpl_loss"][0]) - r_extra) <= 1e-4 * abs(r_extra) + 1e-7
    assert abs(float(out["loss"][0]) - r_total) <= 5e-5 * abs(r_total)
    for (name, a), b in zip(net.state_dict().items(), ref.state_dict().values()):
        if "num_batches_tracked" not in name:
            close(a, b, rtol=0, scale=2 * lr + 1e-6)         # one Adam step: +-lr where a ~0 gradient flips sign


# ------------------------------------------------------------------ plain ProtoNet class (M3) and the cosine logits (P3)
@pytest.mark.parametrize("use_softmax,norm", [(False, None), (True, 2.0)])
def test_prototypical_networks_vs_torch(ops, use_softmax, norm):
    """PrototypicalNetworks (models/prototypical.py:16-43 + few_shot_classifier.py:82-126): feature centering /
    normalisation, prototypes, -cdist scores, optional softmax, and cosine_distance_to_prototypes, vs plain torch."""
    from afsl_b200.models.prototypical import PrototypicalNetworks
    gen = torch.Generator().manual_seed(3)
    ways, shots, nq, din, d = 5, 4, 7, 40, 64
    backbone = torch.nn.Linear(din, d).cuda()
    center = torch.randn(d, generator=gen).cuda() * 0.1
    model = PrototypicalNetworks(backbone=backbone, use_softmax=use_softmax, feature_centering=center, feature_normalization=norm)
    xs = torch.randn(ways * shots, din, generator=gen).cuda()
    xq = torch.randn(ways * nq, din, generator=gen).cuda()
    ys = torch.arange(ways).repeat_interleave(shots)[torch.randperm(ways * shots, generator=gen)].cuda()
    with torch.no_grad():
        model.process_support_set(xs, ys)
        scores = model(xq)
        cos = model.cosine_distance_to_prototypes(model.compute_features(xq))
        fs, fq = backbone(xs) - center, backbone(xq) - center
        if norm is not None:
            fs, fq = torch.nn.functional.normalize(fs, p=norm, dim=1), torch.nn.functional.normalize(fq, p=norm, dim=1)
        protos = torch.stack([fs[ys == w].mean(0) for w in range(ways)])
        ref = -torch.cdist(fq, protos)
        if use_softmax:
            ref = ref.softmax(-1)
        ref_cos = torch.nn.functional.normalize(fq, dim=1) @ torch.nn.functional.normalize(protos, dim=1).T
    close(model.prototypes, protos)
    close(scores, ref)
    close(cos, ref_cos)
    assert not model.is_transductive()
    with pytest.raises(ValueError, match="Illegal backbone or feature shape"):
        model._raise_error_if_features_are_multi_dimensional(torch.zeros(2, 3, 4, 5))


def test_runner_eval_cuda_graph_matches_eager(ops):
    """Single-segment eval_step replayed from a CUDA graph returns the eager step's per-task accuracies."""
    import random
    import bench
    from afsl_b200.episodes import EpisodeRunner, synthetic_batch
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    accs = []
    for graph in (False, True):
        runner = EpisodeRunner(model, bench.EXPERIMENT_CONFIG, None, replay_reference_rng=False, use_cuda_graph=graph)
        torch.manual_seed(5); np.random.seed(5); random.seed(5)
        accs.append(np.concatenate([runner.eval_step(synthetic_batch(6, 5, 5, 5, 157, seed=40 + i), augment_query=True)
                                    for i in range(3)]))
    assert accs[0].shape == (18,) and np.array_equal(accs[0], accs[1])


@pytest.mark.parametrize("path", ["warp", "cta"])
def test_head_edge_cases(ops, monkeypatch, path):
    """One episode, one query per class, a class without support rows (NaN prototype, as the reference's empty mean),
    and a single query row: same outputs as the oracle, NaNs included."""
    monkeypatch.setenv("AFSL_HEAD_WARP", "1" if path == "warp" else "0")
    gen = torch.Generator().manual_seed(8)
    ways, dim = 5, 256
    # (a) a class with no support rows -> its prototype is NaN, every score against it is NaN
    s = torch.randn(1, 8, dim, generator=gen)
    sl = torch.tensor([[0, 1, 1, 2, 4, 4, 0, 2]])                 # class 3 is empty
    q = torch.randn(1, 6, dim, generator=gen)
    ql = torch.tensor([[0, 1, 2, 3, 4, 0]])
    protos = ops.prototypes(s.cuda(), sl.cuda(), n_way=ways)[0].cpu()
    ref = torch.stack([s[0][sl[0] == w].mean(0) for w in range(ways)])
    assert torch.isnan(protos[3]).all() and torch.isnan(ref[3]).all()
    torch.testing.assert_close(protos, ref, rtol=1e-6, atol=1e-6, equal_nan=True)
    loss, _, _ = ops.proto_head(s.cuda(), sl.cuda(), q.cuda(), ql.cuda(), n_way=ways)
    assert torch.isnan(loss).all() and torch.isnan(ohead.fsl_loss(ref, q[0], ql[0]))
    # (b) a single query row, single episode, 1-shot
    s1 = torch.randn(1, ways, dim, generator=gen)
    q1 = torch.randn(1, 1, dim, generator=gen)
    sl1, ql1 = torch.arange(ways).view(1, -1), torch.tensor([[2]])
    sg, qg = s1.cuda().requires_grad_(True), q1.cuda().requires_grad_(True)
    loss1, protos1, correct1 = ops.proto_head(sg, sl1.cuda(), qg, ql1.cuda(), n_way=ways)
    loss1.sum().backward()
    sc, qc = s1[0].clone().requires_grad_(True), q1[0].clone().requires_grad_(True)
    lo = ohead.fsl_loss(ohead.prototypes(sc, sl1[0]), qc, ql1[0])
    lo.backward()
    close(loss1[0], lo)
    close(sg.grad[0], sc.grad)
    close(qg.grad[0], qc.grad)
    assert int(correct1[0]) == ohead.evaluate_task(ohead.l2_scores(qc, ohead.prototypes(sc, sl1[0])), ql1[0])[0]


# ------------------------------------------------------------------ what bench.py times, against the reference / the oracle
def _mirror_model(kind, weights=None):
    """The afsl_b200 counterpart of tests/golden/make_golden.py::_multiseg_models (same configs, same state-dict keys)."""
    from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, SelfAttention, StandardCNN
    from afsl_b200.models.prototypical import ContrastivePrototypicalNetworks, ContrastivePrototypicalNetworksWithoutAttention
    if kind == "concat":
        cfg = {"encoder_name": "CNN", "use_attention": False, "use_contrastive": False, "train_query_augmentations": False,
               "specaug_params": {"use": False}}
        mc = {"Projection": {"input_dim": 64, "hidden_dim": 16, "output_dim": 64}}
        net = ContrastivePrototypicalNetworksWithoutAttention(
            EncoderModule(cfg, {}, encoder=StandardCNN(1, (1, 1, 128, 157), 64, [3, 3], 64)), ProjectionHead(mc))
    else:
        cfg = {"encoder_name": "Hybrid", "use_attention": True, "use_contrastive": False, "train_query_augmentations": True,
               "specaug_params": {"use": True, "mask_param": 16, "W": 22, "num_mask": 1, "mask_value": 0, "p": 0.282}}
        mc = {"Hybrid": {"in_channels": 1, "seq_layers": 1, "seq_type": "RNN", "bidirectional": False, "hidden_channels": 64,
                         "pool_dim": [3, 3], "out_dim": 64},
              "Attention": {"embed_dim": 64, "num_heads": 1, "ffn_dim": 256, "dropout": 0.1},
              "Projection": {"input_dim": 256, "hidden_dim": 16, "output_dim": 64}}
        net = ContrastivePrototypicalNetworks(EncoderModule(cfg, mc), SelfAttention(mc), ProjectionHead(mc))
    if weights is not None:
        net.load_state_dict({k[2:]: t(v) for k, v in weights.items() if k.startswith("w_")})
    return net.cuda().eval(), cfg


def _replay_test_episodes(ds, n_tasks, ways, shots, queries, seed, is_test):
    """The episodes ``n_tasks`` successive sample_episode calls select under ``random.seed(seed)`` (raw sets, no views:
    SpecAugment draws come from other generators and are replayed by the runner)."""
    import random
    from afsl_b200.datasets.batch_creation import sample_episode
    use = ds.specaug_use
    ds.specaug_use = False
    try:
        random.seed(seed)
        return [sample_episode(ds, ways, shots, queries, is_test, "cpu", None, False) for _ in range(n_tasks)]
    finally:
        ds.specaug_use = use


@pytest.mark.parametrize("kind", ["concat", "fused"])
def test_multisegment_evaluation_vs_reference_loop(ops, kind):
    """Config 4, the path bench.py times as multiseg_tasks_per_s: (a) loops.evaluate_multisegment_loop and (b)
    EpisodeRunner.eval_step on ALL tasks packed into one batch (ragged query rows + CSR offsets, draw_ragged SpecAugment
    parameters, ragged proto_eval, eval_vote) return the per-task accuracies / mean / std that the REFERENCE's own
    evaluate_multisegment_loop produced on the same fake dataset, weights and seeds, for the three tie strategies
    (tests/golden/multiseg_eval_*.npz; 'fused' = Hybrid + SpecAugment support/query views + view fusion)."""
    import random
    from afsl_b200.episodes import EpisodeBatch, EpisodeRunner
    from afsl_b200.loops.loops import evaluate_multisegment_loop
    g = load_golden("multiseg_eval_" + kind)
    mg = golden_module()
    net, cfg = _mirror_model(kind, g)
    ds = mg.FakeMultiSegSpecDataset(cfg, seed=int(g["dataset_seed"]))
    n_tasks, ways, shots, queries, seed = (int(g[k]) for k in ("n_tasks", "ways", "shots", "queries", "rng_seed"))
    augment = bool(cfg["specaug_params"]["use"])
    print(f"[{kind}] smallest top-2 score margin in the reference run: {float(g['margins'].min()):.3e} over {g['margins'].size} rows")
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for strat in ("", "min_label", "max_posterior"):
            key = strat or "first"
            # (a) the mirrored loop, one task at a time like the reference
            random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
            msg = evaluate_multisegment_loop(ds, ways, shots, queries, n_tasks, net, "cuda", strat, None, augment)
            assert abs(msg["mean_accuracy"] - float(g["mean_" + key])) < 1e-12, (kind, strat, msg)
            assert abs(msg["accuracy_std"] - float(g["std_" + key])) < 1e-12, (kind, strat, msg)
            # (b) all tasks in one packed batch through the runner
            eps = _replay_test_episodes(ds, n_tasks, ways, shots, queries, seed, True)
            support = torch.stack([e[0][0] for e in eps])
            sl = torch.stack([e[1] for e in eps])
            rows = [e[2][0].shape[0] for e in eps]
            assert rows == g["rows_per_task"].tolist()
            query = torch.cat([e[2][0] for e in eps]).unsqueeze(0)
            ql = torch.cat([e[3] for e in eps]).unsqueeze(0)
            clip_ids = torch.cat([e[4] for e in eps])
            offsets = torch.tensor(np.concatenate([[0], np.cumsum(rows)]), dtype=torch.int64)
            batch = EpisodeBatch(support, sl, query, ql, ways)
            runner = EpisodeRunner(net, cfg, None, replay_reference_rng=True)
            np.random.seed(seed); torch.manual_seed(seed)
            acc = runner.eval_step(batch, augment_query=augment, clip_ids=clip_ids, seg_offsets=offsets, tie_strategy=strat)
            assert np.array_equal(acc, g["acc_" + key]), (kind, strat, acc, g["acc_" + key])
            assert np.mean(acc) == float(g["mean_" + key]) and np.std(acc) == float(g["std_" + key])
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def _structured_batch(episodes, ways, shots, queries, t_len, seed, scale=0.4):
    """synthetic_batch plus a rank-one class pattern per (episode, class): tasks are not pure chance and the argmax over
    prototypes is not decided by the last bits of near-identical distances."""
    from afsl_b200.episodes import synthetic_batch
    b = synthetic_batch(episodes, ways, shots, queries, t_len, seed=seed)
    pat = golden_module().class_patterns(episodes * ways, t_len, seed, scale).view(episodes, ways, 1, 128, t_len)
    b.support += pat[torch.arange(episodes).unsqueeze(1), b.support_labels]
    b.query += pat[torch.arange(episodes).unsqueeze(1), b.query_labels]
    return b


def _oracle_pair(kind, t_len=157, hidden=512, out=256, dropout=0.0):
    """(oracle model on the CPU, afsl_b200 model on the GPU) with identical weights; kind 'fused' or 'concat'."""
    import copy
    import bench
    from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, SelfAttention
    from afsl_b200.models.prototypical import ContrastivePrototypicalNetworks, ContrastivePrototypicalNetworksWithoutAttention
    from oracle import modules as om
    cfg = copy.deepcopy(bench.EXPERIMENT_CONFIG)
    mcfg = copy.deepcopy(bench.MODEL_CONFIG)
    mcfg["Attention"]["dropout"] = dropout
    if kind == "concat":
        cfg["use_attention"] = False
        mcfg["Projection"] = {"input_dim": 64, "hidden_dim": 128, "output_dim": 64}
        pc = mcfg["Projection"]
        ref = om.ConcatViewsNet(om.ViewEncoder(om.build_encoder("Hybrid", t_len)),
                                om.Projection(pc["input_dim"], pc["hidden_dim"], pc["output_dim"]))
        net = ContrastivePrototypicalNetworksWithoutAttention(EncoderModule(cfg, mcfg), ProjectionHead(mcfg))
    else:
        mcfg["Projection"] = {"input_dim": 256, "hidden_dim": hidden, "output_dim": out}
        ref = om.FusedViewsNet(om.ViewEncoder(om.build_encoder("Hybrid", t_len)), om.ViewFusion(64, 1, 256, dropout),
                               om.Projection(256, hidden, out))
        net = ContrastivePrototypicalNetworks(EncoderModule(cfg, mcfg), SelfAttention(mcfg), ProjectionHead(mcfg))
    net.load_state_dict(ref.state_dict())
    for m in list(ref.modules()) + list(net.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return ref, net.cuda(), cfg


def test_eval_step_batched_single_segment_vs_oracle(ops):
    """EpisodeRunner.eval_step on E = 5 tasks in one batch (config 2 model, SpecAugment support + query views, eval-mode
    encoder), eagerly and replayed from a CUDA graph, == oracle.episode.eval_task task by task on the same seeds (the
    restatement of loops/loops.py:66-81,84-121): per-task accuracies bit-equal."""
    import random
    from afsl_b200.episodes import EpisodeRunner
    from oracle import episode as oep
    torch.manual_seed(77)
    ref, net, cfg = _oracle_pair("fused")
    golden_module()._perturb_bn(ref, 5)
    net.load_state_dict(ref.state_dict())
    ref.eval(); net.eval()
    e, ways = 5, 5
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for attempt in range(8):                                    # first seed whose tasks are well conditioned
            seed = 900 + attempt
            batch = _structured_batch(e, ways, 5, 5, 157, seed)
            torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
            want, margin = [], float("inf")
            for i in range(e):
                s_views = oep.make_views(batch.support[i], cfg, True)
                q_views = oep.make_views(batch.query[i], cfg, True)
                want.append(oep.eval_task(ref, s_views, batch.support_labels[i], q_views, batch.query_labels[i]))
                with torch.no_grad():
                    top2 = torch.topk(ref(q_views, inference=True), 2, dim=1).values
                margin = min(margin, float((top2[:, 0] - top2[:, 1]).min() / top2.abs().max()))
            if margin >= 1e-3:
                break
        print(f"oracle accuracies {want}, smallest relative top-2 margin {margin:.3e} (seed {seed})")
        assert margin >= 1e-3
        for graph in (False, True):
            runner = EpisodeRunner(net, cfg, None, replay_reference_rng=True, use_cuda_graph=graph)
            torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
            got = runner.eval_step(batch, augment_query=True)
            assert np.array_equal(got, np.array(want)), (graph, got, want)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_encoder_nchw_path_equals_channels_last(ops, mode):
    """The OPTIONAL fp32 path (AFSL_NCHW_FP32=1, bench.py's labelled nchw_conv_variant): convolutions of blocks 2-4 on NCHW
    tensors (afsl_transpose_f32 after the fused first block, the NCHW family of the BatchNorm/ReLU/pool kernels) against the
    default channels-last path on the same weights and
    inputs, two groups of 25 samples: embeddings 1e-5, running statistics 1e-6, every parameter gradient 2e-3 in L2
    (observed: <= 6e-5 for most, 5e-4 for the first block's weights - the two convolution algorithms round differently,
    so a few pooling windows of blocks 2-4 pick another winner and route their gradient to a neighbouring tap)."""
    import afsl_b200.models.main_modules as mm
    from afsl_b200.models.main_modules import StandardCNN
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    flag = mm.NCHW_FP32_CONVS
    try:
        enc = StandardCNN(1, (1, 1, 128, 157), 64, [3, 3], 64).cuda()
        for m in enc.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if hasattr(m, "group_size"):
                m.group_size = 25
        x = torch.randn(50, 1, 128, 157, device="cuda")
        wts = torch.linspace(-1, 1, 50 * 64, device="cuda").view(50, 64)
        init = {k: v.clone() for k, v in enc.state_dict().items()}
        res = {}
        for nchw in (True, False):
            mm.NCHW_FP32_CONVS = nchw
            enc.load_state_dict(init)
            enc.train(mode == "train")
            enc.zero_grad()
            y = enc(x)
            grads = {}
            if mode == "train":
                (y * wts).sum().backward()
                grads = {n: p.grad.clone() for n, p in enc.named_parameters()}
            res[nchw] = (y.detach().clone(), grads, {k: v.clone() for k, v in enc.state_dict().items() if "running" in k})
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        mm.NCHW_FP32_CONVS = flag
    close(res[True][0], res[False][0], rtol=1e-5)
    for k, v in res[True][2].items():
        close(v, res[False][2][k], rtol=1e-6)
    for n, ga in res[True][1].items():
        gb = res[False][1][n]
        if float(gb.abs().max()) == 0.0:
            assert float(ga.abs().max()) == 0.0
            continue
        rel = float((ga - gb).norm() / gb.norm())
        record_observed(rel, 2e-3)
        assert rel < 2e-3 or float((ga - gb).abs().max()) < 1e-4, (n, rel)


@pytest.mark.parametrize("graph", [False, True])
def test_train_step_batched_vs_oracle(ops, graph):
    """EpisodeRunner.train_step on E = 3 episodes at once (config 2: views, fusion, projection, CPL with sampled negatives;
    eager and CUDA-graph replay) against oracle.episode.train_step episode by episode on the same weights and seeds:
    per-episode losses within 2e-5, the step's gradient (mean over episodes) within 1e-3 relative in L2 norm per parameter
    (elementwise 5e-3 of max|grad|; observed 5e-4).  The optional NCHW convolution path (AFSL_NCHW_FP32=1) does NOT meet
    these bounds (projection-head gradients 9e-3, attention norm2.bias 7e-3, first block 1.1e-3), which is why it is off
    by default; test_encoder_nchw_path_equals_channels_last states what it does meet."""
    import random
    from afsl_b200.episodes import EpisodeRunner
    from oracle import episode as oep
    torch.manual_seed(78)
    ref, net, cfg = _oracle_pair("fused")
    cfg["loss"]["cpl"]["m_param"] = 3
    e = 3
    batch = _structured_batch(e, 5, 5, 5, 157, seed=555)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.0)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(21); np.random.seed(21); random.seed(21)
        want, grads = [], None
        for i in range(e):
            want.append(oep.train_step(ref, ropt, batch.support[i], batch.support_labels[i], batch.query[i],
                                       batch.query_labels[i], cfg))
            g_i = [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in ref.parameters()]
            grads = g_i if grads is None else [a + b for a, b in zip(grads, g_i)]
        torch.manual_seed(21); np.random.seed(21); random.seed(21)
        runner = EpisodeRunner(net, cfg, None, replay_reference_rng=True, use_cuda_graph=graph)
        for p in net.parameters():
            p.grad = None
        out = runner.train_step(batch)
        torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    want = torch.tensor(want)
    close(out["loss"], want[:, 0], rtol=2e-5)
    close(out["fsl_loss"], want[:, 1], rtol=2e-5)
    close(out["cpl_loss"], want[:, 2], rtol=1e-4)
    for (name, p), g_ref in zip(net.named_parameters(), grads):
        if p.grad is None:
            assert float(g_ref.abs().max()) == 0.0, name
            continue
        if name.endswith("0.bias") and "conv_encoder" in name:
            continue                       # convolution bias under batch statistics: exactly 0 here, round-off noise in eager
        # aggregate agreement is tight; single taps may move when a pooling winner flips between the cuDNN and the CPU sums
        # (DESIGN 2), hence the looser elementwise bound
        g_want = g_ref / e
        rel = float((p.grad.cpu() - g_want).norm() / g_want.norm().clamp_min(1e-30))
        record_observed(rel, 1e-3)
        assert rel < 1e-3, (name, rel)
        close(p.grad, g_want, rtol=5e-3)


def test_training_epoch_batched_equals_sequential(ops):
    """loops.training_epoch(episodes_per_step = 2, runner) reports the losses of two sequential one-episode steps when the
    weights do not move (lr = 0): same sampler draws, same per-episode arithmetic (config-1 shaped Conv4 model; the first
    episode is the one tests/golden/epoch_cnn_first.npz pins to the reference's own loop)."""
    import random
    from afsl_b200.episodes import EpisodeRunner
    from afsl_b200.loops.loops import training_epoch
    from afsl_b200.loops.loss import FSL_Loss
    from afsl_b200.models.main_modules import EncoderModule, ProjectionHead, StandardCNN
    from afsl_b200.models.prototypical import ContrastivePrototypicalNetworksWithoutAttention
    g = load_golden("epoch_cnn_plain")
    ds = _FakeDataset(int(g["items_seed"]))
    net = ContrastivePrototypicalNetworksWithoutAttention(
        EncoderModule({"encoder_name": "CNN"}, {}, encoder=StandardCNN(1, (1, 1, 128, 157), 64, [3, 3], 64)),
        ProjectionHead({"Projection": {"input_dim": 64, "hidden_dim": 32, "output_dim": 64}}))
    net.load_state_dict({k[len("init_"):]: t(v) for k, v in g.items() if k.startswith("init_")})
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    net = net.cuda()
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    cfg = {"encoder_name": "CNN", "use_attention": False, "use_contrastive": False, "train_query_augmentations": False,
           "specaug_params": {"use": False}, "project_prototypes": False, "normalize_prototypes": False,
           "loss": {"l_param": 0.0, "cpl": {"use": False}, "angular": {"use": False}}}
    args = (FSL_Loss(), None, 0.0, False, False, 5, 5, 5, None, False, False)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        random.seed(1234); np.random.seed(1234); torch.manual_seed(1234)
        seq = training_epoch(net, ds, opt, 2, "cuda", *args)
        random.seed(1234); np.random.seed(1234); torch.manual_seed(1234)
        bat = training_epoch(net, ds, opt, 2, "cuda", *args, episodes_per_step=2, runner=EpisodeRunner(net, cfg, opt))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    record_observed(abs(bat["loss"] - seq["loss"]), 1e-5 * abs(seq["loss"]))
    assert abs(bat["loss"] - seq["loss"]) <= 1e-5 * abs(seq["loss"]), (bat, seq)
    assert abs(bat["fsl_loss"] - seq["fsl_loss"]) <= 1e-5 * abs(seq["fsl_loss"])
    assert np.isnan(bat["cpl_loss"]) and np.isnan(seq["cpl_loss"])


@pytest.mark.parametrize("ways,shots", [(5, 1), (5, 5), (20, 1), (20, 5)])
def test_2000_task_accuracy_identity(ops, ways, shots):
    """North-star target: reference-identical test accuracy over 2000 tasks at 5/20-way, 1/5-shot (config 5).
    tests/golden/tasks2000_cnn.npz holds what the REFERENCE's evaluate_single_segment produced (Conv4 ProtoNet, eval mode,
    fixed weights, 2000 sampled tasks per configuration) plus its eval-mode embedding of every dataset item.
    (1) Head identity: the GPU head on the reference's own embeddings.  5-way (direct-form distances on both sides):
        bit-identical per-task accuracies, mean and std for all 2000 tasks.  20-way (the reference's cdist is in its matmul
        form, 9e-6 of cancellation noise on these embeddings, 18-22 exact ties per configuration): every one of the 200 000
        row predictions whose exact (float64) top-2 margin exceeds 2e-5 equals the exact argmax, and every task made only of
        such rows has the reference's accuracy.
    (2) Whole path (GPU encoder + head through EpisodeRunner.eval_step, 250 tasks per launch): identical per-task
        accuracy for every task all of whose rows have a top-2 margin above the rounding noise (fp32 convolution sums
        differ between cuDNN and the CPU in the last bits, so nearer ties are not comparable); the number of near-tie
        tasks and of mismatches among them is reported, and the mean accuracy is bounded accordingly."""
    from afsl_b200.episodes import EpisodeBatch, EpisodeRunner
    g = load_golden("tasks2000_cnn")
    key = f"{ways}w{shots}s"
    n_tasks = int(g["n_tasks"])
    mg = golden_module()
    net, cfg = _mirror_model("concat", g)
    ds = mg.FakeManyClassDataset(cfg, seed=int(g["dataset_seed"]))
    # which items each task uses: replay the sampler's index picks only (single-segment clips: nothing else is drawn)
    import random
    from afsl_b200.datasets.batch_creation import _episode_classes
    random.seed(1000 + ways * 10 + shots)
    s_idx = np.empty((n_tasks, ways * shots), dtype=np.int64)
    q_idx = np.empty((n_tasks, ways * 5), dtype=np.int64)
    for i in range(n_tasks):
        s_rows, q_rows = [], []
        for _, s, q in _episode_classes(ds, ways, shots, 5):
            s_rows += s
            q_rows += q
        s_idx[i], q_idx[i] = s_rows, q_rows
    sl = torch.arange(ways).repeat_interleave(shots).expand(n_tasks, -1).contiguous()
    ql = torch.arange(ways).repeat_interleave(5).expand(n_tasks, -1).contiguous()
    want = g["acc_" + key]
    margins = g["margin_" + key]
    # (1) head on the reference's embeddings
    emb = t(g["embeddings"])
    s_idx_t, q_idx_t = torch.from_numpy(s_idx), torch.from_numpy(q_idx)
    pred, _, correct, _ = ops.proto_eval(emb.cuda()[s_idx_t.cuda()], sl.cuda(), emb.cuda()[q_idx_t.cuda()], ql.cuda(), n_way=ways)
    acc_head = correct.cpu().numpy().astype(np.float64) / (ways * 5)
    pred = pred.cpu().view(n_tasks, ways * 5).long()
    # exact (float64) scores from the same embeddings: the arbiter between two fp32 roundings of near-tied distances
    e64 = emb.double()
    protos64 = e64[s_idx_t].view(n_tasks, ways, shots, -1).mean(2)
    d64 = torch.cdist(e64[q_idx_t], protos64)                               # [tasks, Nq, W]
    top2 = torch.topk(-d64, 2, dim=2)
    row_margin = (top2.values[:, :, 0] - top2.values[:, :, 1])              # >= 0
    # embeddings of a random-init network sit close together (norm ~1.6, distances ~0.16): the matmul form
    # |q|^2 + |p|^2 - 2 q.p that cdist uses beyond 25 rows (and the kernels with it) cancels to ~9e-6 absolute in fp32
    # (measured against float64 on these embeddings), the direct form of the 5-way path to ~5e-8
    noise = 2e-5 if ways > 5 else 2e-7
    clear = row_margin > noise
    wrong_rows = (pred != top2.indices[:, :, 0]) & clear
    task_clear = clear.all(1).numpy()
    mism_head = acc_head != want
    print(f"[{key}] head on the reference's embeddings: rows with an exact top-2 margin > {noise:g}: {int(clear.sum())} of "
          f"{clear.numel()}, GPU argmax differs from the exact one on {int(wrong_rows.sum())} of them; tasks made only of "
          f"such rows: {int(task_clear.sum())} of {n_tasks}, per-task accuracy differs from the reference's on "
          f"{int((mism_head & task_clear).sum())} of them ({int(mism_head.sum())} of all {n_tasks} tasks)")
    assert not wrong_rows.any()
    assert not (mism_head & task_clear).any(), np.nonzero(mism_head & task_clear)[0][:10]
    if ways == 5:
        assert not mism_head.any() and np.mean(acc_head) == float(g["mean_" + key]) and np.std(acc_head) == float(g["std_" + key])
    # whole path: the GPU encoder's own rounding (cuDNN vs the CPU's convolution sums, ~1e-6 on the embeddings) comes on top
    solid = (row_margin > max(noise, 1e-5)).all(1).numpy() & (margins > 1e-5)
    # (2) whole path: encoder on the GPU
    runner = EpisodeRunner(net, cfg, None)
    items = ds.items[:, 0].pin_memory()                     # [288, 1, 128, T]
    acc = []
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        step = 250 if ways == 5 else 50
        for lo in range(0, n_tasks, step):
            hi = min(n_tasks, lo + step)
            batch = EpisodeBatch(items[torch.from_numpy(s_idx[lo:hi])], sl[lo:hi], items[torch.from_numpy(q_idx[lo:hi])],
                                 ql[lo:hi], ways)
            acc.append(runner.eval_step(batch))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    acc = np.concatenate(acc)
    mism = acc != want
    print(f"[{key}] reference mean/std {float(g['mean_' + key]):.6f}/{float(g['std_' + key]):.6f}; GPU mean/std "
          f"{np.mean(acc):.6f}/{np.std(acc):.6f}; smallest top-2 margin {margins.min():.3e}; tasks with a row inside the "
          f"rounding noise: {int((~solid).sum())}, mismatching among them: {int((mism & ~solid).sum())}; mismatching among "
          f"the other {int(solid.sum())}: {int((mism & solid).sum())}")
    assert not (mism & solid).any(), np.nonzero(mism & solid)[0][:10]
    assert abs(np.mean(acc) - float(g["mean_" + key])) <= (int((~solid).sum()) + 1e-9) / (n_tasks * ways * 5)


def test_tf32_convolutions_measured_deviation(ops):
    """cuDNN TF32 convolutions (PyTorch's default, what the reference itself would run on an Ampere-or-later GPU) against
    the fp32 convolutions every parity test uses: per-episode losses of one config-2 training step, same weights, batch and
    host-drawn randomness.  The deviation is what bench.py's second (labelled) TF32 line carries; bounds 1e-2 (loss) and 0.15 of max|grad|
    (measured on B200: 7.8e-4 and 5.2e-2 - the gradients are what TF32 costs, which is why the headline is fp32)."""
    import random
    from afsl_b200.episodes import EpisodeRunner
    torch.manual_seed(79)
    _, net, cfg = _oracle_pair("fused")
    batch = _structured_batch(4, 5, 5, 5, 157, seed=556)
    outs = []
    tf32 = torch.backends.cudnn.allow_tf32
    try:
        for allow in (False, True):
            torch.backends.cudnn.allow_tf32 = allow
            torch.manual_seed(31); np.random.seed(31); random.seed(31)
            for p in net.parameters():
                p.grad = None
            out = EpisodeRunner(net, cfg, None, replay_reference_rng=True).train_step(batch)
            grad = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])
            outs.append((out["loss"].clone(), grad.clone()))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    (l32, g32), (ltf, gtf) = outs
    dl = float(((ltf - l32).abs() / l32.abs()).max())
    dg = float((gtf - g32).abs().max() / g32.abs().max())
    print(f"TF32 vs fp32 convolutions: max relative loss deviation {dl:.3e}, max gradient deviation / max|grad| {dg:.3e}")
    record_observed(dl, 1e-2)
    record_observed(dg, 0.15)
    assert dl < 1e-2 and dg < 0.15          # measured on B200: 7.8e-4 and 5.2e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_follow_the_tensors_device(ops):
    """The reference selects cuda:{gpu_index} without set_device (src/train_test.py:44-45): an op on cuda:1 tensors while
    cuda:0 is current must launch on cuda:1 (and mixed-device arguments are refused)."""
    from afsl_b200._lib import AfslError
    torch.cuda.set_device(0)
    gen = torch.Generator().manual_seed(1)
    s, q = torch.randn(4, 25, 64, generator=gen), torch.randn(4, 25, 64, generator=gen)
    lab = torch.arange(5).repeat_interleave(5).expand(4, -1).contiguous()
    l0, p0, c0 = ops.proto_head(s.cuda(0), lab.cuda(0), q.cuda(0), lab.cuda(0), n_way=5)
    l1, p1, c1 = ops.proto_head(s.cuda(1), lab.cuda(1), q.cuda(1), lab.cuda(1), n_way=5)
    assert l1.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(l0.cpu(), l1.cpu()) and torch.equal(p0.cpu(), p1.cpu()) and torch.equal(c0.cpu(), c1.cpu())
    with pytest.raises(AfslError, match="different devices"):
        ops.proto_head(s.cuda(0), lab.cuda(0), q.cuda(1), lab.cuda(1), n_way=5)


# ------------------------------------------------------------------ config-driven driver (src/train_test.py)
@pytest.mark.parametrize("multi_segm,eps", [(False, 1), (True, 4)])
def test_train_test_driver_on_synthetic_episodes(ops, tmp_path, multi_segm, eps):
    """afsl_b200.train_test (mirror of src/train_test.py:20-181) end to end from the two JSON files: model built from the
    configs, contrastive_training_loop with early stopping and checkpoint reload, then the single- or multi-segment test.
    On the class-structured synthetic dataset two short epochs must beat chance clearly (5-way: 0.2)."""
    import copy
    import json
    import bench
    from afsl_b200 import train_test
    cfg = copy.deepcopy(bench.EXPERIMENT_CONFIG)
    cfg.update({"dataset_name": "synthetic", "device": "cuda", "gpu_index": 0, "multi_segm": multi_segm, "tie_strategy": "min_label",
                "n_way_train": 5, "n_way_validation": 5, "n_way_test": 5, "n_shot_train": 3, "n_shot_validation": 3,
                "n_shot_test": 3, "n_query_train": 3, "n_query_validation": 3, "n_query_test": 3, "num_epochs": 2,
                "n_training_tasks": 8, "n_testing_tasks": 6, "patience": 5, "scheduler_milestones": [1], "scheduler_gamma": 0.5,
                "experiment_folder": str(tmp_path / "exp"), "validation_query_augmentations": True, "test_query_augmentations": True,
                "waveaug_params": {"use": False},
                "synthetic": {"classes": 8, "per_class": 8, "t_len": 157, "max_segments": 3 if multi_segm else 1, "scale": 1.0}})
    mcfg = copy.deepcopy(bench.MODEL_CONFIG)
    mcfg["Projection"] = {"input_dim": 256, "hidden_dim": 64, "output_dim": 64}
    e_path, m_path = tmp_path / "experiment_config.json", tmp_path / "model_config.json"
    e_path.write_text(json.dumps(cfg)); m_path.write_text(json.dumps(mcfg))
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        msgs = train_test.main(["-e", str(e_path), "-m", str(m_path), "--runs", "1", "--episodes-per-step", str(eps), "--seed", "3"])
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert (tmp_path / "exp" / "model.pt").exists()
    msg = msgs[0]
    mean = msg["mean_accuracy"] if multi_segm else msg[0]
    print("driver test message:", msg)
    assert 0.35 < float(mean) <= 1.0


# ------------------------------------------------------------------ F2: projection head on the libafsl Linear kernels
def test_projection_head_vs_reference_fixture(ops):
    """ProjectionHead (fc1 -> ReLU -> fc2 -> L2 normalise, all libafsl kernels: afsl_linear_* + afsl_l2_normalize_*) against
    what the REFERENCE's module produced (tests/golden/modules_projection.npz): output, input gradient, parameter gradients."""
    from afsl_b200.models.main_modules import ProjectionHead
    g = load_golden("modules_projection")
    proj = ProjectionHead({"Projection": {"input_dim": 256, "hidden_dim": 128, "output_dim": 256}})
    proj.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith("w_")})
    proj = proj.cuda()
    x = dev(g["x"]).requires_grad_(True)
    before = ops.launch_count()
    y = proj(x)
    y.backward(dev(g["gy"]))
    assert ops.launch_count() - before >= 6          # 3 forward launches + normalise / two Linear backward passes
    close(y, t(g["y"]))
    close(x.grad, t(g["dx"]))
    for name, prm in proj.named_parameters():
        if "g_" + name in g:
            close(prm.grad, t(g["g_" + name]))
        else:
            assert prm.grad is None                  # ln1 / ln2 exist in the state dict but are never applied


@pytest.mark.parametrize("m,n,k,relu,bias", [(960, 512, 256, True, True), (960, 256, 512, False, True), (37, 64, 64, True, False),
                                             (9001, 130, 70, True, True), (1, 256, 256, False, True), (20000, 64, 256, True, True)])
def test_linear_vs_torch(ops, m, n, k, relu, bias):
    """afsl_linear_{fwd,bwd}_f32 (register-tiled fp32 SGEMM; dw / db row-split with a fixed-order reduce beyond 4096 rows)
    against F.linear (+ ReLU) in float64 on the CPU: output and all three gradients."""
    gen = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=gen)
    w = torch.randn(n, k, generator=gen) / k ** 0.5
    b = torch.randn(n, generator=gen) if bias else None
    gy = torch.randn(m, n, generator=gen)
    xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bg = b.cuda().requires_grad_(True) if bias else None
    y = ops.linear(xg, wg, bg, relu=relu)
    y.backward(gy.cuda())
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True) if bias else None
    yd = torch.nn.functional.linear(xd, wd, bd)
    if relu:
        yd = torch.relu(yd)
    yd.backward(gy.double())
    close(y, yd.float())
    close(xg.grad, xd.grad.float())
    close(wg.grad, wd.grad.float())
    if bias:
        close(bg.grad, bd.grad.float())
    # determinism (fixed-order split reduction)
    xg2, wg2 = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    ops.linear(xg2, wg2, bg.detach() if bias else None, relu=relu).backward(gy.cuda())
    assert torch.equal(wg2.grad, wg.grad) and torch.equal(xg2.grad, xg.grad)


# ------------------------------------------------------------------ f-4: waveform front end
def test_log_mel_front_end_vs_torchaudio_and_reference(ops):
    """afsl_logmel_f32 (STFT + power + mel filters + dB + z-normalisation in one launch) against the eager torchaudio chain
    of the reference (datasets/batch_creation.py:138-143,215-218) on the GPU and on the CPU, and - through sample_episode on
    'wav' input - against what the REFERENCE's sample_episode produced (tests/golden/sampler_wav.npz).  Values are dB /
    13.25 (about +-3): bound 2e-5 absolute on the normalised log-mel (fp32 FFT rounding differs from cuFFT's / pocketfft's)."""
    import random
    torchaudio = pytest.importorskip("torchaudio")
    from afsl_b200.datasets.batch_creation import sample_episode
    from test_host_logic import _FakeWavDataset
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_mels=128, n_fft=1024, hop_length=512, power=2.0)
    gen = torch.Generator().manual_seed(4)
    wave = torch.randn(7, 80000, generator=gen) * 0.1
    wave[3] *= 1e-3                                                  # a quiet clip: mel energies near the eps floor
    wave[5, 40000:] = 0.0                                            # digital silence: exactly log10(eps)
    mean, std = -21.5, 13.25
    want_cpu = ((20.0 / 2 * torch.log10(mel(wave) + torch.finfo(torch.float32).eps) - mean) / std).unsqueeze(1)
    mel_gpu = mel.cuda()
    want_gpu = ((20.0 / 2 * torch.log10(mel_gpu(wave.cuda()) + torch.finfo(torch.float32).eps) - mean) / std).unsqueeze(1)
    before = ops.launch_count()
    got = ops.log_mel(wave.cuda(), mel_gpu, mean, std)
    assert ops.launch_count() - before == 1 and got.shape == (7, 1, 128, 157)
    close(got, want_gpu, rtol=0, scale=2e-5)
    close(got, want_cpu, rtol=0, scale=2e-5)
    # the sampler on waveform input, on the GPU, against the reference's own sample_episode
    g = load_golden("sampler_wav")
    ds = _FakeWavDataset(int(g["dataset_seed"]))
    fp = lambda x: x[:, 0, ::16, ::20].reshape(x.shape[0], -1)
    for name, is_test in (("train", False), ("test", True)):
        random.seed(int(g[f"{name}_seed"]))
        s_list, s_lab, q_list, q_lab, ids = sample_episode(ds, 4, 2, 3, is_test, "cuda", mel_gpu, False)
        assert list(q_list[0].shape) == g[f"{name}_shape"].tolist() and q_list[0].is_cuda
        close(fp(s_list[0]), t(g[f"{name}_support"]), rtol=0, scale=2e-5)
        close(fp(q_list[0]), t(g[f"{name}_query"]), rtol=0, scale=2e-5)
        assert torch.equal(q_lab, t(g[f"{name}_query_labels"])) and torch.equal(ids, t(g[f"{name}_audio_ids"]))
