"""Per-role cycle counters of the tcgen05 fitting GEMM (dpb200_fit_gemm_debug): where does a tile's time go?"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from deepmd_kit_b200 import ops  # noqa: E402
from deepmd_kit_b200._lib import lib  # noqa: E402
from deepmd_kit_b200.model import FittingNet  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("N", 131072))
net = FittingNet(1600, (240, 240, 240), True, 3, torch.float64, dev)
assert net.prepare_tc(6)
g = torch.Generator().manual_seed(5)
d = (torch.randn(n, 1600, generator=g, dtype=torch.float64) * 0.05).to(dev)
xs, ex = ops.split_i8_rows(d, 6)
tc = net.tc
nb = ops.fit_blocked_rows(n)
t = torch.empty(nb * 240, dtype=torch.float64, device=dev)
y = torch.empty(nb * 240, dtype=torch.float64, device=dev)
t2 = torch.empty_like(t)
y2 = torch.empty_like(t)
sl = torch.empty((n, 6 * 240), dtype=torch.int8, device=dev)
sl2 = torch.empty_like(sl)
gd = torch.empty((n, 1600), dtype=torch.float64, device=dev)
exr = torch.zeros(n, dtype=torch.int32, device=dev)
w, b, idt = net.layers[0]
w1, b1, idt1 = net.layers[1]
bsl, cs, Kp = tc["fw"][0]
bsl1, cs1, Kp1 = tc["fw"][1]
bslb, csb, Kpb = tc["bw"][0]
bslh, csh, Kph = tc["bw"][1]
cases = {
    "L0 fwd": lambda: ops.fit_gemm_i8(0, n, 240, 1600, xs, 1600, xs.stride(0), ex, 0, bsl, Kp, cs,
                                      out0=t, out1=y, slices_out=sl, ld_slices=1440, kp_out=240, out_exp=tc["exp"][0]),
    "hidden fwd": lambda: ops.fit_gemm_i8(0, n, 240, 240, sl, 240, 1440, None, tc["exp"][0], bsl1, Kp1, cs1,
                                          skip=y, out0=t2, out1=y2, slices_out=sl2, ld_slices=1440, kp_out=240,
                                          out_exp=tc["exp"][1]),
    "hidden bwd": lambda: ops.fit_gemm_i8(1, n, 240, 240, sl, 240, 1440, exr, 0, bslh, Kph, csh, skip=y,
                                          t_in=t, out0=t2, out1=y2),
    "L0 bwd": lambda: ops.fit_gemm_i8(2, n, 1600, 240, sl, 240, 1440, exr, 0, bslb, Kpb, csb, out0=gd, ld_out=1600),
}
dbg = torch.zeros((148, 16), dtype=torch.int64, device=dev)
names = ["prod.total", "prod.wait_empty", "mma.total", "mma.wait_tempty", "mma.wait_full", "epi2.total", "epi2.wait_tfull",
         "epi2.drain", "epi2.math", "epi21.total", "epi21.wait_tfull", "epi21.drain", "epi21.math"]
for name, fn in cases.items():
    fn()
    torch.cuda.synchronize()
    dbg.zero_()
    lib().cdll.dpb200_fit_gemm_debug(C.c_void_p(dbg.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib().cdll.dpb200_fit_gemm_debug(None)
    m = dbg.double().mean(0).tolist()
    tiles = (n + 127) // 128 * (240 if "L0 bwd" != name else 1600) // 80 / 148
    print(f"== {name}: tiles/CTA {tiles:.1f}")
    for k, nm in enumerate(names):
        print(f"   {nm:18s} {m[k]:12.0f} cycles   {m[k] / tiles:9.0f} /tile")
