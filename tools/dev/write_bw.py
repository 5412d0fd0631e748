import torch, time
x = torch.empty(1 << 28, device="cuda", dtype=torch.float32)  # 1 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.zero_()); print("fill 1 GiB", ms, "ms", x.numel() * 4 / ms / 1e6, "GB/s written")
ms = t(lambda: y.copy_(x)); print("copy 1 GiB", ms, "ms", 2 * x.numel() * 4 / ms / 1e6, "GB/s read+write")
ms = t(lambda: x.sum()); print("read 1 GiB", ms, "ms", x.numel() * 4 / ms / 1e6, "GB/s read")
