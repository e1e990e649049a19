"""Library GEMM rates at the fitting-net shapes (decides the split-integer / split-TF32 design).
usage: python tools/gemm_probe.py"""
import torch

dev = "cuda"
N = 1 << 17


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for K, Nn in [(1600, 240), (3200, 240), (9600, 240), (240, 240), (1440, 240), (240, 1600), (1440, 1600), (1600, 256), (256, 1600)]:
    a8 = torch.randint(-64, 64, (N, K), dtype=torch.int8, device=dev)
    b8 = torch.randint(-64, 64, (K, Nn), dtype=torch.int8, device=dev)
    try:
        ms = t(lambda: torch._int_mm(a8, b8))
        print(f"int8  [{N}x{K}]x[{K}x{Nn}] {ms:8.3f} ms  {2 * N * K * Nn / ms / 1e9:9.1f} TOP/s", flush=True)
        b8t = b8.t().contiguous().t()
        ms = t(lambda: torch._int_mm(a8, b8t))
        print(f"int8 (B col-major)            {ms:8.3f} ms  {2 * N * K * Nn / ms / 1e9:9.1f} TOP/s", flush=True)
    except Exception as e:
        print("int8 failed", K, Nn, repr(e)[:300], flush=True)
    del a8, b8
    for dt, name in [(torch.float64, "f64"), (torch.float32, "f32"), (torch.bfloat16, "bf16")]:
        if K > 3200 and dt == torch.float64:
            continue
        a = torch.randn(N, K, dtype=dt, device=dev)
        b = torch.randn(K, Nn, dtype=dt, device=dev)
        torch.backends.cuda.matmul.allow_tf32 = False
        ms = t(lambda: a @ b)
        print(f"{name:5s} [{N}x{K}]x[{K}x{Nn}] {ms:8.3f} ms  {2 * N * K * Nn / ms / 1e9:9.1f} TFLOP/s", flush=True)
        if dt == torch.float32:
            torch.backends.cuda.matmul.allow_tf32 = True
            ms = t(lambda: a @ b)
            print(f"tf32  [{N}x{K}]x[{K}x{Nn}] {ms:8.3f} ms  {2 * N * K * Nn / ms / 1e9:9.1f} TFLOP/s", flush=True)
            torch.backends.cuda.matmul.allow_tf32 = False
        del a, b
