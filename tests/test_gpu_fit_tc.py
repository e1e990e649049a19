"""Fitting net on the tcgen05 tensor cores (csrc/fit_tc.cu, through the C ABI): every int8 split GEMM against the
fp64 product it replaces, and the whole forward + backward chain against the plain fp64 net
(deepmd/pt/model/network/mlp.py algebra: y = tanh(x W + b) * idt (+ x), energy head, dE/dD).

Tolerances: the split product drops terms below 2^-48 of (row scale x column scale) per K element, so the error
bound used here is 1e-12 x rowmax x colmax x sqrt(K) (measured 1.5e-13); the chain is held to 5e-12 of the
largest reference magnitude (measured 4e-13) -- two orders of magnitude inside the 1e-10 of the end-to-end tests.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g

    return g.load_package()


def _pack(w, ns=6):
    """[K, N] fp64 weights -> B operand slices [ns, N, Kp], colv [N, 4] (scale, 0, 1, 0), Kp."""
    from deepmd_kit_b200.model import split_i8_cols

    sl, ce = split_i8_cols(w, ns)
    K, N = w.shape
    Kp = (K + 63) // 64 * 64
    b = torch.zeros((ns, N, Kp), dtype=torch.int8)
    b[:, :, :K] = sl.permute(0, 2, 1)
    colv = torch.zeros((N, 4), dtype=torch.float64)
    colv[:, 0] = torch.ldexp(torch.ones(N, dtype=torch.float64), ce - 14)
    colv[:, 2] = 1.0
    return b.contiguous().to(DEV), colv.to(DEV), Kp


@pytest.mark.parametrize("n,K,N", [(1, 64, 16), (127, 64, 80), (128, 48, 96), (129, 240, 240), (300, 1600, 240),
                                   (257, 240, 1600), (700, 128, 320)])
def test_fit_gemm_plain_matches_fp64(pkg, n, K, N):
    """mode 2 (row-major product): ragged row counts, K tails (zero-filled by TMA), every cluster width
    (N / 80 column tiles: 1, 2, 3 -> cluster 3, 4 / 20 -> cluster 4)."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(n + K + N)
    x = torch.randn(n, K, generator=g, dtype=torch.float64) * torch.exp(2 * torch.randn(n, 1, generator=g, dtype=torch.float64))
    w = torch.randn(K, N, generator=g, dtype=torch.float64) / K ** 0.5 * torch.exp(torch.randn(1, N, generator=g, dtype=torch.float64))
    xd = x.to(DEV)
    xs, ex = ops.split_i8_rows(xd, 6)
    bsl, colv, Kp = _pack(w)
    out = torch.full((n, N), float("nan"), dtype=torch.float64, device=DEV)
    ops.fit_gemm_i8(2, n, N, K, xs, K, 6 * K, ex, 0, bsl, Kp, colv, out0=out, ld_out=N)
    want = xd @ w.to(DEV)
    scale = xd.abs().amax(1, keepdim=True) * w.to(DEV).abs().amax(0, keepdim=True) * K ** 0.5
    assert not torch.isnan(out).any()
    assert ((out - want).abs() / scale).max().item() < 1e-12


@pytest.mark.parametrize("n,K,N", [(130, 64, 80), (300, 1600, 240), (257, 240, 1600), (600, 128, 160)])
def test_fit_gemm_four_slices(pkg, n, K, N):
    """nslice = 4 (the fp32 model's operands: 31 fraction bits below 2^E with max in [2^(E-2), 2^(E-1)), orders >= 4
    dropped: a few 2^-28 per term; measured 2^-26.5): product to 2^-25 of the row x column scale, stored as
    fp64 (mode 2) and as float32 (mode 3).  Column-tile counts 1, 2 (cluster 1), 3 (no 3-wide cluster for 4 slices)
    and 20 (cluster 4)."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(n + K + N)
    x = torch.randn(n, K, generator=g, dtype=torch.float64) * torch.exp(2 * torch.randn(n, 1, generator=g, dtype=torch.float64))
    w = torch.randn(K, N, generator=g, dtype=torch.float64) / K ** 0.5
    xd = x.to(DEV)
    xs, ex = ops.split_i8_rows(xd, 4)
    bsl, colv, Kp = _pack(w, 4)
    want = xd @ w.to(DEV)
    scale = xd.abs().amax(1, keepdim=True) * w.to(DEV).abs().amax(0, keepdim=True) * K ** 0.5
    out = torch.full((n, N), float("nan"), dtype=torch.float64, device=DEV)
    ops.fit_gemm_i8(2, n, N, K, xs, K, 4 * K, ex, 0, bsl, Kp, colv, out0=out, ld_out=N, nslice=4)
    assert ((out - want).abs() / scale).max().item() < 2.0 ** -25
    out32 = torch.full((n, N), float("nan"), dtype=torch.float32, device=DEV)
    ops.fit_gemm_i8(3, n, N, K, xs, K, 4 * K, ex, 0, bsl, Kp, colv, out0=out32, ld_out=N, nslice=4)
    assert torch.equal(out32, out.to(torch.float32))


@pytest.mark.parametrize("n", [1, 200, 4097])
def test_fit_net_tc_fp32_model(pkg, n):
    """An fp32 net through the same kernels with 4 slices: energies and dE/dD (returned as float32) against the
    fp64 evaluation of the same (float32-valued) weights, 2e-6 of the largest magnitude -- the fp32 model's end-to-end
    tolerance is 1e-5."""
    from deepmd_kit_b200.model import FittingNet

    ops = pkg.ops
    K0 = 1600
    net = FittingNet(K0, (240, 240, 240), True, 11, torch.float32, DEV)
    assert net.prepare_tc(4)
    assert not net.prepare_tc(6)
    ref = FittingNet(K0, (240, 240, 240), True, 11, torch.float64, DEV)
    ref.layers = [(w.double(), b.double(), None if i is None else i.double()) for w, b, i in net.layers]
    ref.head = (net.head[0].double(), net.head[1].double())
    g = torch.Generator().manual_seed(n)
    d = (torch.randn(n, K0, generator=g, dtype=torch.float64) * 0.05 *
         torch.exp(torch.randn(n, 1, generator=g, dtype=torch.float64))).to(DEV)
    e0, g0 = ref.forward_backward(d)
    xs, ex = ops.split_i8_rows(d, 4)
    e1, g1 = net.forward_backward_tc(xs, ex, n)
    assert e1.dtype == torch.float32 and g1.dtype == torch.float32
    assert ((e1.double() - e0).abs().max() / e0.abs().max()).item() < 2e-6
    assert ((g1.double() - g0).abs().max() / g0.abs().max()).item() < 2e-6


def test_fit_blocked_roundtrip_and_slices(pkg):
    """Row-blocked layout conversion is exact both ways; fit_slice_rows reproduces the rows to 2^-47 of the row
    maximum with |digit| <= 128 and the documented exponent rule |x| < 2^(E-1)."""
    ops = pkg.ops
    torch.manual_seed(4)
    n, N = 333, 240
    x = torch.randn(n, N, dtype=torch.float64, device=DEV) * torch.exp(3 * torch.randn(n, 1, dtype=torch.float64, device=DEV))
    x[5] = 0.0  # an all-zero row must not overflow the exponent
    xb = ops.fit_blocked(x, n, N, True)
    assert torch.equal(ops.fit_blocked(xb, n, N, False), x)
    sl, ex = ops.fit_slice_rows(xb, n, N, N)
    d = sl.reshape(n, 6, N).to(torch.float64)
    assert int(d.abs().max()) <= 128
    wgt = torch.tensor([2.0 ** (-7 - 8 * s) for s in range(6)], dtype=torch.float64, device=DEV)
    rec = (d * wgt[None, :, None]).sum(1) * torch.ldexp(torch.ones(n, dtype=torch.float64, device=DEV), ex)[:, None]
    rowmax = x.abs().amax(1, keepdim=True).clamp_min(1e-300)
    assert ((rec - x).abs() / rowmax).max().item() <= 2.0 ** -46
    live = x.abs().amax(1) > 0
    assert bool((x.abs().amax(1)[live] < torch.ldexp(torch.ones_like(ex[live], dtype=torch.float64), ex[live] - 1)).all())


@pytest.mark.parametrize("n", [1, 200, 4097])
@pytest.mark.parametrize("resnet_dt", [True, False])
def test_fit_net_tc_matches_plain_fp64(pkg, n, resnet_dt):
    """The whole chain (3 forward GEMMs with tanh / idt / skip epilogues, energy head, 3 backward GEMMs) against the
    plain fp64 net with the same weights."""
    from deepmd_kit_b200.model import FittingNet

    ops = pkg.ops
    K0 = 1600
    net = FittingNet(K0, (240, 240, 240), resnet_dt, 11, torch.float64, DEV)
    assert net.prepare_tc(6)
    g = torch.Generator().manual_seed(n)
    d = (torch.randn(n, K0, generator=g, dtype=torch.float64) * 0.05 *
         torch.exp(torch.randn(n, 1, generator=g, dtype=torch.float64))).to(DEV)
    e0, g0 = net.forward_backward(d)
    xs, ex = ops.split_i8_rows(d, 6)
    e1, g1 = net.forward_backward_tc(xs, ex, n)
    assert ((e1 - e0).abs().max() / e0.abs().max()).item() < 5e-12
    assert ((g1 - g0).abs().max() / g0.abs().max()).item() < 5e-12


def test_fit_net_tc_other_widths(pkg):
    """A narrower net (descriptor 512 -> 128 -> 128): one and two column tiles per layer, cluster width 1."""
    from deepmd_kit_b200.model import FittingNet

    ops = pkg.ops
    net = FittingNet(512, (128, 128), True, 5, torch.float64, DEV)
    assert net.prepare_tc(6)
    torch.manual_seed(9)
    n = 515
    d = torch.randn(n, 512, dtype=torch.float64, device=DEV) * 0.1
    e0, g0 = net.forward_backward(d)
    xs, ex = ops.split_i8_rows(d, 6)
    e1, g1 = net.forward_backward_tc(xs, ex, n)
    assert ((e1 - e0).abs().max() / e0.abs().max()).item() < 5e-12
    assert ((g1 - g0).abs().max() / g0.abs().max()).item() < 5e-12


def test_fit_tc_rejects_unsupported(pkg):
    from deepmd_kit_b200.model import FittingNet

    ops = pkg.ops
    assert not FittingNet(240, (240, 240), True, 1, torch.float64, DEV).prepare_tc(6)   # skip connection on layer 0
    assert not FittingNet(1600, (240, 120), True, 1, torch.float64, DEV).prepare_tc(6)  # unequal hidden widths
    assert not FittingNet(1600, (240, 240), True, 1, torch.float32, DEV).prepare_tc(6)  # fp32 nets take 4 slices
    assert not FittingNet(1600, (240, 240), True, 1, torch.float64, DEV).prepare_tc(4)  # fp64 nets take 6
    x = torch.zeros((4, 6 * 64), dtype=torch.int8, device=DEV)
    ex = torch.zeros(4, dtype=torch.int32, device=DEV)
    b = torch.zeros((6, 24, 64), dtype=torch.int8, device=DEV)
    colv = torch.zeros((24, 4), dtype=torch.float64, device=DEV)
    out = torch.zeros((4, 24), dtype=torch.float64, device=DEV)
    with pytest.raises(ValueError):  # N must be a multiple of 16
        ops.fit_gemm_i8(2, 4, 24, 64, x, 64, 6 * 64, ex, 0, b, 64, colv, out0=out, ld_out=24)
    with pytest.raises(RuntimeError):  # CPU tensors are rejected: there is no CPU path
        ops.fit_gemm_i8(2, 4, 24, 64, x.cpu(), 64, 6 * 64, ex, 0, b, 64, colv, out0=out, ld_out=24)
