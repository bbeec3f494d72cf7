"""A/B of the encoder's two fp32 layouts (NCHW after the first block vs channels-last throughout) at the shape of
tests::test_train_step_batched_vs_oracle: per-block outputs and the final embeddings.   python tools/nchw_ab_probe.py [n]"""
import sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
import afsl_b200.models.main_modules as mm
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
torch.manual_seed(0)
enc = mm.get_backbone_model("Hybrid", bench.MODEL_CONFIG).to(dev)
for m in enc.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, "group_size"): m.group_size = 25
x = torch.randn(n, 1, 128, 157, device=dev)
init = {k: v.clone() for k, v in enc.state_dict().items()}
res = {}
for flag in (True, False):
    mm.NCHW_FP32_CONVS = flag
    enc.load_state_dict(init)
    enc.train()
    outs = []
    hooks = [blk.register_forward_hook(lambda m, i, o: outs.append(o.detach().contiguous().clone())) for blk in enc.conv_encoder]
    y = enc(x)
    for h in hooks: h.remove()
    res[flag] = (outs, y.detach().clone())
for i, (a, b) in enumerate(zip(res[True][0], res[False][0])):
    d = (a - b).abs()
    per_group = d.view(n // 25, -1).max(dim=1).values
    print(f"block {i}: shape {tuple(a.shape)} max abs diff {float(d.max()):.3e} (scale {float(b.abs().max()):.3e}); worst groups {per_group.topk(min(3, per_group.numel())).indices.tolist()} nonzero-diff elems {int((d > 1e-4).sum())}")
d = (res[True][1] - res[False][1]).abs()
print(f"embeddings: max abs diff {float(d.max()):.3e} scale {float(res[False][1].abs().max()):.3e}")
