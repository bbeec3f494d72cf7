import torch, time
torch.backends.cudnn.benchmark = True
dev = "cuda"
def bench(fmt, n=3200, h=42, w=52):
    x = torch.randn(n, 64, h, w, device=dev).contiguous(memory_format=fmt).requires_grad_(True)
    conv = torch.nn.Conv2d(64, 64, 3, padding=1, bias=False).to(dev).to(memory_format=fmt)
    gy = torch.randn(n, 64, h, w, device=dev).contiguous(memory_format=fmt)
    for _ in range(3):
        y = conv(x); y.backward(gy)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        y = conv(x); y.backward(gy)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5, y.is_contiguous(memory_format=fmt)
for shape in ((3200, 42, 52), (3200, 14, 17), (3200, 4, 5)):
    print(shape, "nchw", bench(torch.contiguous_format, *shape), "nhwc", bench(torch.channels_last, *shape))
