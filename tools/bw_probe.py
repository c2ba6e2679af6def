import torch, time
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda").normal_()   # 4 GiB
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.sum())
print("read-only sum: %.2f TB/s" % (x.numel() * 4 / ms / 1e9))
ms = t(lambda: y.copy_(x))
print("copy (r+w): %.2f TB/s" % (2 * x.numel() * 4 / ms / 1e9))
ms = t(lambda: torch.max(x))
print("read-only max: %.2f TB/s" % (x.numel() * 4 / ms / 1e9))
xs = x[: 1 << 26]
ms = t(lambda: xs.sum(), 20)
print("read-only sum 256MB: %.2f TB/s" % (xs.numel() * 4 / ms / 1e9))
xs = x[: 42 * (1 << 20)]
ms = t(lambda: xs.sum(), 50)
print("read-only sum 168MB (per-kernel launch incl.): %.2f TB/s, %.1f us" % (xs.numel() * 4 / ms / 1e9, ms * 1e3))
