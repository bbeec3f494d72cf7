"""How far the CPU oracle's OWN gradients of a 3-episode config-2 step move when the weights of convolution blocks 2-4 are
perturbed by 1e-7 relative (a last-bit change of the convolution sums, what separates two convolution algorithms): the
first block's weights move by ~4e-4 in L2 (pooling windows pick another winner) and attention norm2.bias by ~2e-4 (the
prototype loss is invariant to a common shift of all fused tokens, so that gradient is the remainder of a cancelling sum);
every other parameter stays within ~4e-6.  Justifies the bounds of tests/test_gpu_parity.py::test_train_step_batched_vs_oracle.
    python tools/grad_sensitivity.py        (CPU only)"""
import sys, random, copy
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import bench
from oracle import episode as oep, modules as om
torch.manual_seed(78)
cfg = copy.deepcopy(bench.EXPERIMENT_CONFIG)
ref = om.FusedViewsNet(om.ViewEncoder(om.build_encoder("Hybrid", 157)), om.ViewFusion(64, 1, 256, 0.0), om.Projection(256, 512, 256))
for m in ref.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
cfg["loss"]["cpl"]["m_param"] = 3
import importlib.util
spec = importlib.util.spec_from_file_location("tgp", "/root/repo/tests/test_gpu_parity.py")
# structured batch helper re-implemented: use synthetic batch from episodes (CPU tensors)
from afsl_b200.episodes import synthetic_batch
batch = synthetic_batch(3, 5, 5, 5, 157, seed=555)
def run(model, dtype):
    model = copy.deepcopy(model).to(dtype)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    torch.manual_seed(21); np.random.seed(21); random.seed(21)
    grads = None
    for i in range(3):
        oep.train_step(model, opt, batch.support[i].to(dtype), batch.support_labels[i], batch.query[i].to(dtype), batch.query_labels[i], cfg)
        g_i = [p.grad.detach().clone().double() if p.grad is not None else torch.zeros_like(p).double() for p in model.parameters()]
        grads = g_i if grads is None else [a + b for a, b in zip(grads, g_i)]
    return grads
g32 = run(ref, torch.float32)
ref2 = copy.deepcopy(ref)
torch.manual_seed(5)
with torch.no_grad():
    for n, p in ref2.named_parameters():
        if "conv_encoder" in n and n.endswith("0.weight") and "conv_encoder.0." not in n:
            p.mul_(1 + 1e-7 * torch.randn_like(p))
g32b = run(ref2, torch.float32)
for (name, p), a, b in zip(ref.named_parameters(), g32, g32b):
    nb = float(b.norm())
    print(f"{name:60s} rel L2 after 1e-7 perturbation of conv 2-4 weights {float((a-b).norm())/max(nb,1e-30):.3e}   |g| {nb:.3e}")
