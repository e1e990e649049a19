"""HBM bandwidth by access mix (torch ops): write-only, read-only, copy."""
import torch
n = 1 << 30  # 8 GiB of fp64
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, nbytes, name):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:12s} {nbytes / best / 1e6:8.1f} GB/s  ({best:.2f} ms)")
t(lambda: a.fill_(1.0), n * 8, "write-only")
t(lambda: a.zero_(), n * 8, "memset")
t(lambda: a.sum(), n * 8, "read-only")
t(lambda: b.copy_(a), n * 16, "copy r+w")
