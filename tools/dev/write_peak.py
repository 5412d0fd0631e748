"""dev: write-only and copy bandwidth of this GPU with torch's own kernels (the denominators a write-only kernel can be held to)."""
import torch
x = torch.empty(1 << 28, device="cuda")  # 1 GiB
y = torch.empty(1 << 28, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
tf = t(lambda: x.fill_(1.0)); tc = t(lambda: y.copy_(x))
print("fill  1 GiB: %.1f us -> %.0f GB/s written" % (tf * 1e6, x.numel() * 4 / tf / 1e9))
print("copy  1 GiB: %.1f us -> %.0f GB/s read+written" % (tc * 1e6, 2 * x.numel() * 4 / tc / 1e9))
