"""fp32 (TF32 off) 64->64 3x3 convolutions of encoder stages 2-4 at the bench step's shapes (6400 sample-views per step:
32 episodes x 50 samples x 4 views; MaxPool 3): cuDNN forward + backward time in NCHW and channels-last layouts
(cudnn.benchmark on), to see whether another layout / algorithm beats the channels-last kernels the step uses.
    python tools/conv_fp32_probe.py"""
import torch
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
n = 6400
tot = {"nchw": 0.0, "nhwc": 0.0}
for (h, w) in ((42, 52), (14, 17), (4, 5)):
    for fmt_name, fmt in (("nchw", torch.contiguous_format), ("nhwc", torch.channels_last)):
        conv = torch.nn.Conv2d(64, 64, 3, padding=1, bias=False).to(dev).to(memory_format=fmt)
        x = torch.randn(n, 64, h, w, device=dev).to(memory_format=fmt).requires_grad_(True)
        def step():
            y = conv(x)
            y.backward(y.detach())
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        tot[fmt_name] += ms
        flops = 3 * 2.0 * n * h * w * 64 * 64 * 9
        print(f"{fmt_name} [{n},64,{h},{w}] fwd+bwd {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
    if (h, w) == (42, 52):
        x = torch.randn(n, 64, h, w, device=dev).to(memory_format=torch.channels_last)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            y = x.contiguous()
        b.record()
        torch.cuda.synchronize()
        print(f"nhwc -> nchw copy of [{n},64,{h},{w}]: {a.elapsed_time(b) / 5:.3f} ms", flush=True)
print(tot)
