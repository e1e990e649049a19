"""Numerics and timing of the tcgen05 int8-split fitting GEMMs (csrc/fit_tc.cu) against torch fp64 matmul.
Run on a B200: python tools/fit_tc_check.py [--big]"""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from deepmd_kit_b200 import ops  # noqa: E402
from deepmd_kit_b200.model import FittingNet, split_i8_cols  # noqa: E402

dev = torch.device("cuda:0")


def pack(w, ns=6):
    sl, ce = split_i8_cols(w, ns)
    K, N = w.shape
    Kp = (K + 63) // 64 * 64
    b = torch.zeros((ns, N, Kp), dtype=torch.int8)
    b[:, :, :K] = sl.permute(0, 2, 1)
    colv = torch.zeros((N, 4), dtype=torch.float64)
    colv[:, 0] = torch.ldexp(torch.ones(N, dtype=torch.float64), ce - 14)
    colv[:, 2] = 1.0
    return b.contiguous().to(dev), colv.to(dev), Kp


def plain(n, K, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, K, generator=g, dtype=torch.float64) * torch.exp(torch.randn(n, 1, generator=g, dtype=torch.float64))
    w = torch.randn(K, N, generator=g, dtype=torch.float64) / K ** 0.5
    xd = x.to(dev)
    xs, ex = ops.split_i8_rows(xd, 6)
    bsl, cs, Kp = pack(w)
    out = torch.full((n, N), float("nan"), dtype=torch.float64, device=dev)
    ops.fit_gemm_i8(2, n, N, K, xs, K, 6 * K, ex, 0, bsl, Kp, cs, out0=out, ld_out=N)
    torch.cuda.synchronize()
    want = xd @ w.to(dev)
    scale = (xd.abs().amax(1, keepdim=True) * w.to(dev).abs().amax(0, keepdim=True)) * K ** 0.5
    err = ((out - want).abs() / scale).max().item()
    nan = int(torch.isnan(out).sum().item())
    print(f"[plain] n={n} K={K} N={N}: max |err| / (rowmax*colmax*sqrt(K)) = {err:.3e}  nan={nan}", flush=True)
    if nan or err > 1e-11:
        bad = ((out - want).abs() / scale)
        r, c = divmod(int(bad.nan_to_num(1e9).argmax().item()), N)
        print("    worst at", r, c, "got", out[r, c].item(), "want", want[r, c].item())
        print("    row-block errs:", [(bad[i:i + 128].nan_to_num(1e9).max().item()) for i in range(0, min(n, 512), 128)])
        print("    col-tile errs:", [(bad[:, j:j + 16].nan_to_num(1e9).max().item()) for j in range(0, min(N, 256), 16)])
    return err


def full(n, seed=3, K0=1600):
    net = FittingNet(K0, (240, 240, 240), True, seed, torch.float64, dev)
    assert net.prepare_tc(6)
    g = torch.Generator().manual_seed(seed)
    d = (torch.randn(n, K0, generator=g, dtype=torch.float64) * 0.05).to(dev)
    e0, g0 = net.forward_backward(d)
    xs, ex = ops.split_i8_rows(d, 6)
    e1, g1 = net.forward_backward_tc(xs, ex, n)
    torch.cuda.synchronize()
    ee = ((e1 - e0).abs().max() / e0.abs().max()).item()
    ge_ = ((g1 - g0).abs().max() / g0.abs().max()).item()
    print(f"[full] n={n}: energy rel err {ee:.3e}, dE/dD rel err {ge_:.3e}", flush=True)
    return net, d, xs, ex


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    for shape in [(128, 64, 80), (128, 64, 16), (100, 128, 80), (300, 240, 240), (200, 48, 96), (1000, 1600, 240), (257, 240, 1600),
                  (4096, 1600, 240)]:
        try:
            plain(*shape)
        except Exception:
            traceback.print_exc()
            return 1
    try:
        full(1000)
        net, d, xs, ex = full(20000)
    except Exception:
        traceback.print_exc()
        return 1
    if "--big" in sys.argv:
        n = 131072
        g = torch.Generator().manual_seed(5)
        d = (torch.randn(n, 1600, generator=g, dtype=torch.float64) * 0.05).to(dev)
        xs, ex = ops.split_i8_rows(d, 6)
        net.prepare_split(6)
        t_tc = timeit(lambda: net.forward_backward_tc(xs, ex, n))
        t_sp = timeit(lambda: net.forward_backward_split(xs, ex, n))
        t_pl = timeit(lambda: net.forward_backward(d))
        print(f"[time] n={n}: tcgen05 {t_tc:.3f} ms | cuBLASLt int8 L0 + DGEMM {t_sp:.3f} ms | DGEMM {t_pl:.3f} ms", flush=True)
        # per-GEMM timing
        tc = net.tc
        nb = ops.fit_blocked_rows(n)
        t = torch.empty(nb * 240, dtype=torch.float64, device=dev)
        y = torch.empty(nb * 240, dtype=torch.float64, device=dev)
        sl = torch.empty((n, 6 * 240), dtype=torch.int8, device=dev)
        bsl, cs, Kp = tc["fw"][0]
        w, b, idt = net.layers[0]
        t0 = timeit(lambda: ops.fit_gemm_i8(0, n, 240, 1600, xs, 1600, xs.stride(0), ex, 0, bsl, Kp, cs,
                                            out0=t, out1=y, slices_out=sl, ld_slices=1440, kp_out=240, out_exp=tc["exp"][0]))
        bsl1, cs1, Kp1 = tc["fw"][1]
        exr0 = torch.zeros(n, dtype=torch.int32, device=dev)
        w1, b1, idt1 = net.layers[1]
        t2 = torch.empty_like(t)
        y2 = torch.empty_like(t)
        sl2 = torch.empty_like(sl)
        t1 = timeit(lambda: ops.fit_gemm_i8(0, n, 240, 240, sl, 240, 1440, None, tc["exp"][0], bsl1, Kp1, cs1,
                                            skip=y, out0=t2, out1=y2, slices_out=sl2, ld_slices=1440, kp_out=240,
                                            out_exp=tc["exp"][1]))
        bslh, csh, Kph = tc["bw"][1]
        t1b = timeit(lambda: ops.fit_gemm_i8(1, n, 240, 240, sl, 240, 1440, exr0, 0, bslh, Kph, csh, skip=y, t_in=t,
                                             out0=t2, out1=y2))
        bslb, csb, Kpb = tc["bw"][0]
        gd = torch.empty((n, 1600), dtype=torch.float64, device=dev)
        exr = torch.zeros(n, dtype=torch.int32, device=dev)
        t3 = timeit(lambda: ops.fit_gemm_i8(2, n, 1600, 240, sl, 240, 1440, exr, 0, bslb, Kpb, csb, out0=gd, ld_out=1600))
        t4 = timeit(lambda: ops.fit_slice_rows(y, n, 240, 240))
        t5 = timeit(lambda: ops.fit_head(t, y, tc["w_head"], idt, tc["b_head"], n, 240, 240))
        f0 = 2.0 * n * 1600 * 240 * 21 / 1e12
        f1 = 2.0 * n * 240 * 240 * 21 / 1e12
        print(f"[time] L0 fwd {t0:.3f} ms ({f0 / t0 * 1e3:.0f} TOP/s int8) | hidden fwd {t1:.3f} ms ({f1 / t1 * 1e3:.0f}) | hidden bwd {t1b:.3f} ms | "
              f"L0 bwd {t3:.3f} ms ({f0 / t3 * 1e3:.0f}) | slice {t4:.3f} ms | head {t5:.3f} ms", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
