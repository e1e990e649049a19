"""Compressed se_e2_a energy / force / virial evaluation — the per-step hot path, driven from
PyTorch host code through the dpb200 operators.

Pipeline per step (the decomposed graph of deepmd/tf/descriptor/se_a.py:665-681,753-783 with the
descriptor algebra of deepmd/pt/model/descriptor/se_a.py:838-850):

    [every `nlist_every` steps] normalize_coord -> copy_coord (ghost images) -> build_nlist(rcut+skin)
    prod_env_mat_a (format with the true rcut + env-mat)      -> em, em_deriv, rij, nlist
    tabulate_fusion_se_a over the type sections               -> xyz_scatter [nloc,4,M]
    se_a_descriptor: (GR/nnei)^T (GR/nnei)[:, :axis]; fitting net fwd + hand-written bwd (cuBLAS GEMMs)
    tabulate_fusion_se_a_grad over the type sections          -> net_deriv = dE/d(em)
    use_nlist_map (ghost -> owner)                            -> prod_force_a + prod_virial_a

``DeepPotB200.eval`` is the public call (argument meaning of deepmd.infer.DeepPot.eval): host
coordinates in, host energy / force / virial out.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .compress import EmbeddingNet, compress_se_a

# davg / dstd of the reference's se_e2_a water model (source/tests/infer/deeppot_sea.yaml), per
# centre type: (davg of component 0, dstd of component 0, dstd of components 1..3)
WATER_STATS = [(0.05033, 0.13984, 0.08580), (0.04810, 0.12388, 0.07672)]


@dataclass
class SeAConfig:
    """examples/water/se_e2_a/input.json by default."""
    ntypes: int = 2
    sel: Sequence[int] = (46, 92)
    rcut: float = 6.0
    rcut_smth: float = 0.5
    neuron: Sequence[int] = (25, 50, 100)
    axis_neuron: int = 16
    fitting_neuron: Sequence[int] = (240, 240, 240)
    fitting_resnet_dt: bool = True
    stats: Sequence[Sequence[float]] = field(default_factory=lambda: list(WATER_STATS))
    seed: int = 1
    stride0: float = 0.01
    stride1: float = 0.1
    extrapolate: float = 5.0
    min_nbor_dist: float = 0.9

    @property
    def nnei(self) -> int:
        return int(sum(self.sel))

    @property
    def sec(self) -> List[int]:
        s = [0]
        for v in self.sel:
            s.append(s[-1] + int(v))
        return s


COPPER_CONFIG = dict(ntypes=1, sel=(512,), rcut=8.0, rcut_smth=2.0, stats=[(0.06, 0.12, 0.07)], min_nbor_dist=2.0)


def split_i8_cols(w: torch.Tensor, nslice: int):
    """Host-side split of an fp64 weight matrix [K, N] into balanced base-256 digit slices per COLUMN scale:
    w[:, c] = 2^col_exp[c] * sum_j slice_j[:, c] 2^(-7-8j).  Returns (slices int8 [nslice, K, N], most
    significant first; col_exp int32 [N]).  Same digit convention as csrc/fitting.cu / fit_tc.cu / tabulate.cu."""
    w = w.detach().to("cpu", torch.float64)
    m = w.abs().amax(0)
    _, ex = torch.frexp(torch.where(m > 0, m, torch.ones_like(m)))  # m = mant * 2^ex, mant in [0.5, 1)
    E = (ex + 1).to(torch.int32)  # |w| * 2^-E < 0.5
    P = 7 + 8 * (nslice - 1)
    q = torch.round(torch.ldexp(w, (P - E).to(torch.int32).unsqueeze(0).expand_as(w))).to(torch.int64)
    digits = []
    for _ in range(nslice):  # least significant first: d = ((q + 128) mod 256) - 128 in [-128, 127]
        d = ((q + 128) & 255) - 128
        digits.append(d.to(torch.int8))
        q = (q - d) >> 8
    return torch.stack(digits[::-1]), E


def _tf32_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away, as cvt.rna.tf32.f32) of fp32 values to 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    return torch.where(torch.isfinite(x), r, x)


def split_tf32_weight(w: torch.Tensor):
    w = w.to(torch.float32)
    hi = _tf32_round(w)
    lo = _tf32_round(w - hi)
    return hi, lo


class _tf32_matmul:
    """TF32 tensor-core GEMMs for the 3xTF32 products only (operands are pre-split, so the products are exact)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *a):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


class FittingNet:
    """Energy fitting net of one atom type: D -> neuron... -> 1, tanh, `resnet_dt` skip connections
    on equal-width layers (deepmd/pt/model/network/mlp.py), default normal init, bias_atom_e = 0."""

    def __init__(self, dim_in: int, neuron: Sequence[int], resnet_dt: bool, seed: int, dtype, device):
        g = torch.Generator().manual_seed(seed)
        self.layers = []
        n_in = dim_in
        for n_out in neuron:
            w = torch.empty(n_in, n_out, dtype=torch.float64).normal_(0.0, 1.0 / math.sqrt(n_in + n_out), generator=g)
            b = torch.empty(n_out, dtype=torch.float64).normal_(0.0, 1.0, generator=g)
            idt = torch.empty(n_out, dtype=torch.float64).normal_(0.1, 0.001, generator=g) if resnet_dt else None
            self.layers.append((w.to(device, dtype), b.to(device, dtype), None if idt is None else idt.to(device, dtype)))
            n_in = n_out
        w = torch.empty(n_in, 1, dtype=torch.float64).normal_(0.0, 1.0 / math.sqrt(n_in + 1), generator=g)
        b = torch.empty(1, dtype=torch.float64).normal_(0.0, 1.0, generator=g)
        self.head = (w.to(device, dtype), b.to(device, dtype))
        self.dtype = dtype
        self.split = None

    def prepare_split(self, nslice: int = 7):
        """Pre-split weights for the tensor-core GEMMs (see csrc/fitting.cu):
        fp64: first layer as reversed int8 slice stack [nslice*K, N] + column exponents;
        fp32: every layer as 3xTF32 stacks [Whi; Whi; Wlo] (forward) and of W^T (backward)."""
        dev = self.layers[0][0].device
        if self.dtype == torch.float64:
            w0 = self.layers[0][0]
            sl, ce = split_i8_cols(w0, nslice)  # [ns, K, N]
            wrev = torch.cat([sl[k] for k in range(nslice - 1, -1, -1)], 0).contiguous()  # W_{ns-1}; ...; W_0
            self.split = dict(nslice=nslice, wrev=wrev.to(dev), col_exp=ce.to(dev), K=w0.shape[0])
        else:
            fw, bw = [], []
            for li, (w, b, idt) in enumerate(self.layers):
                hi, lo = split_tf32_weight(w)
                if li == 0:
                    fw.append((torch.cat([hi, hi], 0).contiguous(), lo.contiguous()))  # operand [hi | lo]
                else:
                    fw.append(torch.cat([hi, hi, lo], 0).contiguous())  # operand [hi | lo | hi]
                ht, lt = hi.t().contiguous(), lo.t().contiguous()
                bw.append(torch.cat([ht, ht, lt], 0).contiguous())
            self.split = dict(fw=fw, bw=bw, K=self.layers[0][0].shape[0])
        return self

    def prepare_tc(self, nslice: int = 6):
        """Weights of all GEMMs of the net (forward and backward) as balanced base-256 int8 digit slices in the B-operand
        layout of dpb200_fit_gemm_i8 ([nslice][N][K padded to 64], K contiguous) + column scales 2^(col_exp-12),
        and the fixed exponents of the hidden activations (|y_l| < sum_k<=l max|idt_k|).  Returns False when the
        architecture is outside what csrc/fit_tc.cu covers (first layer without skip connection, equal-width hidden
        layers; 6 slices for an fp64 net, 4 for an fp32 net: 47 / 31 fraction bits per operand -- the fp32 net runs
        through the same kernels, its intermediates are fp64, and dE/dD leaves the last GEMM as float32)."""
        if (self.dtype, nslice) not in ((torch.float64, 6), (torch.float32, 4)):
            return False
        dev = self.layers[0][0].device
        widths = [w.shape for w, _, _ in self.layers]
        if widths[0][0] in (widths[0][1], 2 * widths[0][1]) or widths[0][1] == 2 * widths[0][0] or widths[0][0] % 16:
            return False
        for (k_in, k_out) in widths[1:]:
            if k_in != k_out:
                return False

        def pack(w, add=None, mul=None):  # w [K, N] -> B slices [ns, N, Kp], colv [N, 4] = {scale, add, mul, 0}, Kp
            sl, ce = split_i8_cols(w, nslice)
            K, N = w.shape
            Kp = (K + 63) // 64 * 64
            b = torch.zeros((nslice, N, Kp), dtype=torch.int8)
            b[:, :, :K] = sl.permute(0, 2, 1)
            colv = torch.zeros((N, 4), dtype=torch.float64)
            colv[:, 0] = torch.ldexp(torch.ones(N, dtype=torch.float64), ce - 14)
            if add is not None:
                colv[:, 1] = add.detach().to("cpu", torch.float64)
            colv[:, 2] = 1.0 if mul is None else mul.detach().to("cpu", torch.float64)
            return b.contiguous().to(dev), colv.contiguous().to(dev), Kp

        if any(w.shape[1] % 16 for w, _, _ in self.layers):
            return False
        tc = dict(nslice=nslice, fw=[], bw=[], exp=[])
        bound = 0.0
        nl = len(self.layers)
        for li, (w, b, idt) in enumerate(self.layers):
            w64 = w.detach().to("cpu", torch.float64)
            tc["fw"].append(pack(w64, add=b, mul=idt))
            if li == 0:
                tc["bw"].append(pack(w64.t().contiguous()))
            else:  # epilogue of the backward GEMM of layer li: g_{li-1} (+ head weights on top), times idt_{li-1}
                tc["bw"].append(pack(w64.t().contiguous(), add=self.head[0][:, 0] if li == nl - 1 else None,
                                     mul=self.layers[li - 1][2]))
            amp = 1.0 if idt is None else float(idt.abs().max())
            bound = amp if li == 0 else bound + amp
            tc["exp"].append(int(math.floor(math.log2(bound))) + 2)
        tc["w_head"] = self.head[0][:, 0].to(torch.float64).contiguous()
        tc["b_head"] = float(self.head[1][0])
        idt_last = self.layers[-1][2]
        tc["idt_last"] = None if idt_last is None else idt_last.to(torch.float64).contiguous()
        self.tc = tc
        return True

    @torch.no_grad()
    def forward_backward_tc(self, xs: torch.Tensor, row_exp: torch.Tensor, n: int, grad_cols: Optional[int] = None):
        """forward_backward on the pre-split descriptor rows with every GEMM on the int8 tensor cores
        (csrc/fit_tc.cu): per layer ONE kernel does the split product, the recombination and the layer's
        elementwise chain; the activations travel between layers as int8 slices.  Returns (e [n], dE/dD [n, K0]);
        grad_cols (a multiple of 16): only the first grad_cols input columns of the gradient are formed (se_atten: the
        descriptor part of [D | type embedding])."""
        tc = self.tc
        ns = tc["nslice"]
        dev = xs.device
        nb = ops.fit_blocked_rows(n)
        K0 = self.layers[0][0].shape[0]
        ts, ys = [], []
        a, a_ss, a_rs, a_exp, a_fixed, K = xs, K0, xs.stride(0), row_exp, 0, K0
        nl = len(self.layers)
        for li, (w, b, idt) in enumerate(self.layers):
            N = w.shape[1]
            bsl, cs, Kp = tc["fw"][li]
            t = torch.empty(nb * N, dtype=torch.float64, device=dev)
            y = torch.empty(nb * N, dtype=torch.float64, device=dev)
            last = li == nl - 1
            kp_out = (N + 15) // 16 * 16
            sl = None if last else torch.empty((n, ns * kp_out), dtype=torch.int8, device=dev)
            ops.fit_gemm_i8(0, n, N, K, a, a_ss, a_rs, a_exp, a_fixed, bsl, Kp, cs, skip=ys[-1] if li > 0 else None,
                            out0=t, out1=y, slices_out=sl, ld_slices=0 if last else ns * kp_out, kp_out=kp_out,
                            out_exp=tc["exp"][li], nslice=ns)
            ts.append(t)
            ys.append(y)
            if not last:
                a, a_ss, a_rs, a_exp, a_fixed, K = sl, kp_out, ns * kp_out, None, tc["exp"][li], kp_out
        N = self.layers[-1][0].shape[1]
        kp = (N + 15) // 16 * 16
        e, dz, dz_exp = ops.fit_head(ts[-1], ys[-1], tc["w_head"], tc["idt_last"], tc["b_head"], n, N, kp, ns)
        g_prev = None
        for li in range(nl - 1, 0, -1):
            w, b, idt = self.layers[li]
            n_in = w.shape[0]
            bsl, cs, Kp = tc["bw"][li]
            need_g = li - 1 > 0
            g = torch.empty(nb * n_in, dtype=torch.float64, device=dev) if need_g else None
            dzn = torch.empty(nb * n_in, dtype=torch.float64, device=dev)
            ops.fit_gemm_i8(1, n, n_in, kp, dz, kp, ns * kp, dz_exp, 0, bsl, Kp, cs, skip=g_prev, t_in=ts[li - 1],
                            out0=g, out1=dzn, nslice=ns)
            g_prev = g
            kp = (n_in + 15) // 16 * 16
            dz, dz_exp = ops.fit_slice_rows(dzn, n, n_in, kp, ns)
        bsl, cs, Kp = tc["bw"][0]
        ng = K0 if grad_cols is None else int(grad_cols)
        if ng != K0:  # the weight rows of the wanted input columns as their own operand (slice stride = ng rows)
            key = ("bw0", ng)
            if key not in tc:
                tc[key] = (bsl[:, :ng, :].contiguous(), cs[:ng].contiguous(), Kp)
            bsl, cs, Kp = tc[key]
        gd = torch.empty((n, ng), dtype=self.dtype, device=dev)
        ops.fit_gemm_i8(2 if self.dtype == torch.float64 else 3, n, ng, kp, dz, kp, ns * kp, dz_exp, 0, bsl, Kp, cs,
                        out0=gd, ld_out=ng, nslice=ns)
        return e.to(self.dtype), gd

    def to(self, device, dtype):
        self.layers = [(w.to(device, dtype), b.to(device, dtype), None if i is None else i.to(device, dtype))
                       for w, b, i in self.layers]
        self.head = (self.head[0].to(device, dtype), self.head[1].to(device, dtype))
        return self

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        for w, b, idt in self.layers:
            y = torch.tanh(torch.addmm(b, x, w))
            if idt is not None:
                y = y * idt
            if w.shape[0] == w.shape[1]:
                y = y + x
            elif w.shape[1] == 2 * w.shape[0]:
                y = y + torch.cat([x, x], 1)
            x = y
        return torch.addmm(self.head[1], x, self.head[0]).reshape(-1)

    @torch.no_grad()
    def forward_backward(self, x: torch.Tensor):
        """Atomic energies e[n] and dE/dx[n, dim_in] with a hand-written backward: six cuBLAS GEMMs; every
        elementwise chain (tanh, resnet_dt, skip / its derivative) is one dpb200 pass, the skip connection
        of the backward rides in the GEMM epilogue (beta = 1)."""
        acts = []
        h = x
        for w, b, idt in self.layers:
            a = torch.addmm(b, h, w)
            same = w.shape[0] == w.shape[1]
            y = ops.mlp_tanh_fwd(a, h if same else None, idt)  # a <- tanh(a)
            if w.shape[1] == 2 * w.shape[0]:
                y.add_(torch.cat([h, h], 1))
            acts.append(a)
            h = y
        e = torch.addmv(self.head[1], h, self.head[0][:, 0])
        g = self.head[0][:, 0].unsqueeze(0).expand(x.shape[0], -1)  # stride-0 rows: never materialised
        for (w, b, idt), a in zip(reversed(self.layers), reversed(acts)):
            t = ops.mlp_tanh_bwd(g, a, idt)
            if w.shape[0] == w.shape[1]:
                g = torch.addmm(g, t, w.t())
            elif w.shape[1] == 2 * w.shape[0]:
                n_in = w.shape[0]
                g = torch.addmm(g[:, :n_in] + g[:, n_in:], t, w.t())
            else:
                g = t @ w.t()
        return e, g


    @torch.no_grad()
    def forward_backward_split(self, xs: torch.Tensor, row_exp, n: int):
        """forward_backward on a PRE-SPLIT descriptor block (rows of tabulate_sections_desc mode 2; xs may hold
        more than n rows, the extra ones are ignored).  fp64: the first GEMM runs as `nslice` error-free int8
        tensor-core GEMMs (one per order, K-concatenated) recombined in fp64; fp32: every GEMM is a 3xTF32 product."""
        sp = self.split
        if self.dtype == torch.float64:
            ns, K = sp["nslice"], sp["K"]
            w0, b0, idt0 = self.layers[0]
            N = w0.shape[1]
            m = max(n, 32)  # torch._int_mm needs more than 16 rows
            acc = torch.empty((ns, m, N), dtype=torch.int32, device=xs.device)
            for d in range(ns):
                torch._int_mm(xs[:m, : (d + 1) * K], sp["wrev"][(ns - 1 - d) * K:], out=acc[d])
            if m != n:
                acc = acc[:, :n].contiguous()
            a0, h = ops.split_i8_combine(acc, row_exp, sp["col_exp"], bias=b0, idt=idt0)
            del acc
            acts = [a0]
            for w, b, idt in self.layers[1:]:
                a = torch.addmm(b, h, w)
                same = w.shape[0] == w.shape[1]
                y = ops.mlp_tanh_fwd(a, h if same else None, idt)
                if w.shape[1] == 2 * w.shape[0]:
                    y.add_(torch.cat([h, h], 1))
                acts.append(a)
                h = y
            e = torch.addmv(self.head[1], h, self.head[0][:, 0])
            g = self.head[0][:, 0].unsqueeze(0).expand(n, -1)
            for (w, b, idt), a in zip(reversed(self.layers), reversed(acts)):
                t = ops.mlp_tanh_bwd(g, a, idt)
                if w.shape[0] == w.shape[1]:
                    g = torch.addmm(g, t, w.t())
                elif w.shape[1] == 2 * w.shape[0]:
                    n_in = w.shape[0]
                    g = torch.addmm(g[:, :n_in] + g[:, n_in:], t, w.t())
                else:
                    g = t @ w.t()
            return e, g
        K = sp["K"]
        with _tf32_matmul():
            acts = []
            h = None
            hs = None
            for li, (w, b, idt) in enumerate(self.layers):
                if li == 0:
                    wcat, wlo = sp["fw"][0]
                    a = torch.addmm(b, xs[:n, : 2 * K], wcat)
                    a.addmm_(xs[:n, :K], wlo)
                else:
                    a = torch.addmm(b, hs, sp["fw"][li])
                same = w.shape[0] == w.shape[1]
                y, ys = ops.mlp_tanh_fwd(a, h if (same and li > 0) else None, idt, split3=True)
                if li > 0 and w.shape[1] == 2 * w.shape[0]:
                    y.add_(torch.cat([h, h], 1))
                    ys = ops.split_tf32(y, 3)
                acts.append(a)
                h, hs = y, ys
            e = torch.addmv(self.head[1], h, self.head[0][:, 0])
            g = self.head[0][:, 0].unsqueeze(0).expand(n, -1)
            for li in range(len(self.layers) - 1, -1, -1):
                w, b, idt = self.layers[li]
                t3 = ops.mlp_tanh_bwd(g, acts[li], idt, split3=True)
                if li > 0 and w.shape[0] == w.shape[1]:
                    g = torch.addmm(g, t3, sp["bw"][li])
                elif li > 0 and w.shape[1] == 2 * w.shape[0]:
                    n_in = w.shape[0]
                    g = torch.addmm(g[:, :n_in] + g[:, n_in:], t3, sp["bw"][li])
                else:
                    g = t3 @ sp["bw"][li]
        return e, g


class SeAModel:
    """Random-init compressed se_e2_a model (weights of the named architecture, tabulated with the
    restated `dp compress`)."""

    def __init__(self, cfg: SeAConfig, dtype=torch.float64, device="cuda", weights: Optional[dict] = None):
        """weights=None: random-init nets of the architecture in `cfg` with the WATER_STATS-style `cfg.stats`.
        weights = dict(davg, dstd [ntypes, nnei, 4], embed=[(weights, biases) per neighbour type], fit=[dict(layers=[(w, b,
        idt|None)], head=(w, b)) per centre type], bias_atom_e [ntypes]): a trained model (see from_reference)."""
        self.cfg = cfg
        self.dtype = dtype
        self.device = torch.device(device)
        nnei = cfg.nnei
        if weights is None:
            davg = np.zeros((cfg.ntypes, nnei, 4))
            dstd = np.ones((cfg.ntypes, nnei, 4))
            for t, (a0, s0, s1) in enumerate(cfg.stats):
                davg[t, :, 0] = a0
                dstd[t, :, 0] = s0
                dstd[t, :, 1:] = s1
            self.embed = [EmbeddingNet(cfg.neuron, cfg.seed + 17 * t) for t in range(cfg.ntypes)]
        else:
            davg = np.asarray(weights["davg"], np.float64).reshape(cfg.ntypes, nnei, 4)
            dstd = np.asarray(weights["dstd"], np.float64).reshape(cfg.ntypes, nnei, 4)
            self.embed = []
            for ws, bs in weights["embed"]:
                net = EmbeddingNet(cfg.neuron, 0)
                net.weights = [torch.as_tensor(np.asarray(w, np.float64)) for w in ws]
                net.biases = [torch.as_tensor(np.asarray(b, np.float64)).reshape(-1) for b in bs]
                assert [tuple(w.shape) for w in net.weights] == [
                    (a, b) for a, b in zip([1] + list(cfg.neuron[:-1]), cfg.neuron)], "embedding net shape"
                self.embed.append(net)
        self.bias_atom_e = torch.zeros(cfg.ntypes, dtype=dtype, device=self.device)
        if weights is not None and weights.get("bias_atom_e") is not None:
            self.bias_atom_e = torch.as_tensor(np.asarray(weights["bias_atom_e"], np.float64).reshape(-1), dtype=dtype,
                                               device=self.device)
        self.davg_np, self.dstd_np = davg, dstd
        self.davg = torch.as_tensor(davg.reshape(cfg.ntypes, -1), dtype=dtype, device=self.device)
        self.dstd = torch.as_tensor(dstd.reshape(cfg.ntypes, -1), dtype=dtype, device=self.device)
        tables, infos = compress_se_a(self.embed, davg[:, 0, :], dstd[:, 0, :], cfg.sel, cfg.min_nbor_dist,
                                      cfg.rcut_smth, cfg.rcut, cfg.stride0, cfg.stride1, cfg.extrapolate)
        self.tables64 = tables
        self.tables = [t.to(self.device, dtype).contiguous() for t in tables]
        self.infos = [i.to(dtype) for i in infos]  # host
        self.M = int(cfg.neuron[-1])
        dim_d = self.M * cfg.axis_neuron
        self.fit = [FittingNet(dim_d, cfg.fitting_neuron, cfg.fitting_resnet_dt, cfg.seed + 101 * t, dtype, self.device)
                    for t in range(cfg.ntypes)]
        if weights is not None and weights.get("fit") is not None:  # (descriptor-only weight sets keep the random fit)
            for f, fw in zip(self.fit, weights["fit"]):
                f.layers = [(torch.as_tensor(np.asarray(w, np.float64)).to(self.device, dtype),
                             torch.as_tensor(np.asarray(b, np.float64)).reshape(-1).to(self.device, dtype),
                             None if i is None else torch.as_tensor(np.asarray(i, np.float64)).reshape(-1).to(self.device, dtype))
                            for w, b, i in fw["layers"]]
                hw, hb = fw["head"]
                f.head = (torch.as_tensor(np.asarray(hw, np.float64)).reshape(-1, 1).to(self.device, dtype),
                          torch.as_tensor(np.asarray(hb, np.float64)).reshape(1).to(self.device, dtype))
        import os as _os

        self.fit_chunk = int(_os.environ.get("DPB200_FIT_CHUNK", str(1 << 17)))
        # Tensor-core fitting net (csrc/fitting.cu): the descriptor leaves the tabulate forward already split
        # (int8 slices in fp64, TF32 head/tail in fp32).  Needs axis == 16 (fp64) / axis % 4 == 0 (fp32),
        # M <= 128 and a first fitting layer without skip connection.
        self.nslice = 6  # 7 + 5*8 = 47 fraction bits per operand: GEMM error ~1e-14 of the row*column scale
        w0 = self.fit[0].layers[0][0]
        ok_axis = cfg.axis_neuron == 16 if dtype == torch.float64 else cfg.axis_neuron % 4 == 0
        self.use_split = bool(ok_axis and self.M <= 128 and cfg.axis_neuron <= 32 and w0.shape[1] not in
                              (w0.shape[0], 2 * w0.shape[0]) and self.device.type == "cuda")
        # fp64: every GEMM of the net on the tcgen05 int8 path (csrc/fit_tc.cu); DPB200_FIT_TC=0 keeps the
        # library GEMMs (cuBLASLt int8 first layer + DGEMM) of round 1 as the comparison path.
        # fp32: the same kernels with 4 slices (31 fraction bits); the descriptor leaves the table kernel as int8
        # slices of its 32-bit fixed-point image (mode 3).  Without them: 3xTF32 library GEMMs (mode 2).
        self.use_tc = False
        self.desc_mode, self.desc_nslice = 2, self.nslice
        if self.use_split:
            for f in self.fit:
                f.prepare_split(self.nslice)
            if _os.environ.get("DPB200_FIT_TC", "1") != "0":
                ns_tc = 6 if dtype == torch.float64 else 4
                if cfg.axis_neuron == 16 and all([f.prepare_tc(ns_tc) for f in self.fit]):
                    self.use_tc = True
                    self.desc_nslice = ns_tc
                    self.desc_mode = 2 if dtype == torch.float64 else 3
        # Compressed table coefficients (include/dpb200.h DPB200_TAB_COMPRESSED_COEF): the table kernels are bound by
        # the on-chip coefficient stream; for a dp-compress table the high-order terms tolerate fp32 / fp16 storage
        # at the 1e-12 level.  Checked per table here, on the host, once.
        self.coef_flags = None
        import os

        if self.device.type == "cuda" and os.environ.get("DPB200_TAB_COMPRESS", "1") != "0":
            gate = ops.compressed_coef_flags if dtype == torch.float64 else ops.compressed_coef_flags_f32
            fl = [gate(t, i) for t, i in zip(self.tables64, infos)]
            if all(fl):
                self.coef_flags = fl

    @classmethod
    def from_reference(cls, data: dict, dtype=torch.float64, device="cuda", min_nbor_dist: float = 0.8, **compress):
        """A trained reference se_e2_a + energy model (type_one_side) from its serialized weights: `data` is the dict
        written by tests/golden/make_deeppot_sea.py from a deepmd.pt state_dict (atomic_model.descriptor.sea.{mean,
        stddev, filter_layers.networks.<neighbour type>}, atomic_model.fitting_net.{filter_layers.networks.<centre
        type>, bias_atom_e}; deepmd/pt/model/descriptor/se_a.py, deepmd/pt/model/task/fitting.py) plus the
        model_def_script entries sel, rcut, rcut_smth, neuron, axis_neuron, fitting neuron / resnet_dt.  The embedding
        nets are tabulated here (compress.py = `dp compress`), everything else is used as is."""
        d = data["descriptor"]
        f = data["fitting_net"]
        cfg = SeAConfig(ntypes=len(d["sel"]), sel=tuple(d["sel"]), rcut=float(d["rcut"]), rcut_smth=float(d["rcut_smth"]),
                        neuron=tuple(d["neuron"]), axis_neuron=int(d["axis_neuron"]),
                        fitting_neuron=tuple(f["neuron"]), fitting_resnet_dt=bool(f["resnet_dt"]),
                        min_nbor_dist=float(min_nbor_dist), **compress)
        return cls(cfg, dtype, device, weights=data["weights"])

    # -- descriptor contraction (dpb200 kernels) + fitting net (cuBLAS GEMMs through torch, hand-written
    #    backward).  The fitting net is library code here; SURVEY 8f-1 lists its fusion as the next item.
    def energy_and_dy(self, xyz: torch.Tensor, type_perm: torch.Tensor, type_ranges):
        """xyz: [nloc,4,M] raw tabulate output. Returns (E_total, atom_energy[nloc], dE/d(xyz))."""
        cfg = self.cfg
        nloc = xyz.shape[0]
        dy = torch.empty_like(xyz)
        e_atom = torch.empty(nloc, dtype=self.dtype, device=xyz.device)
        type_perm32 = type_perm.to(torch.int32)
        inv = 1.0 / cfg.nnei
        for t, (a, b) in enumerate(type_ranges):
            for c0 in range(a, b, self.fit_chunk):
                c1 = min(b, c0 + self.fit_chunk)
                idx = type_perm32[c0:c1]
                d = ops.se_a_descriptor(xyz, cfg.axis_neuron, inv, rows=idx)  # gathers the atoms of this type
                e, gd = self.fit[t].forward_backward(d)
                del d
                ops.se_a_descriptor_grad(gd, xyz, cfg.axis_neuron, inv, rows=idx, out=dy)  # scatters back
                e_atom.index_copy_(0, type_perm[c0:c1], e + self.bias_atom_e[t])
        return e_atom.sum(), e_atom, dy

    def energy_and_dy_split(self, xyz, desc, row_exp, type_perm, type_ranges):
        """energy_and_dy on the pre-split descriptor rows emitted by tabulate_sections_desc (type-sorted order)."""
        cfg = self.cfg
        nloc = xyz.shape[0]
        dy = torch.empty_like(xyz)
        e_atom = torch.empty(nloc, dtype=self.dtype, device=xyz.device)
        type_perm32 = type_perm.to(torch.int32)
        inv = 1.0 / cfg.nnei
        for t, (a, b) in enumerate(type_ranges):
            for c0 in range(a, b, self.fit_chunk):
                c1 = min(b, c0 + self.fit_chunk)
                if self.use_tc:
                    e, gd = self.fit[t].forward_backward_tc(desc[c0:c1], row_exp[c0:c1], c1 - c0)
                else:
                    e, gd = self.fit[t].forward_backward_split(desc[c0:], None if row_exp is None else row_exp[c0:c1],
                                                               c1 - c0)
                ops.se_a_descriptor_grad(gd, xyz, cfg.axis_neuron, inv, rows=type_perm32[c0:c1], out=dy)
                e_atom.index_copy_(0, type_perm[c0:c1], e + self.bias_atom_e[t])
        return e_atom.sum(), e_atom, dy

    def bytes_per_atom(self) -> int:
        """Device bytes of per-atom intermediates of one evaluation (env-mat, its derivative, rij, nlist, dE/d(em),
        table output and its gradient, the descriptor in GEMM operand form, fitting activations)."""
        cfg = self.cfg
        F = 8 if self.dtype == torch.float64 else 4
        K = self.M * cfg.axis_neuron
        desc = (self.desc_nslice * K if self.desc_mode == 3 or self.dtype == torch.float64 else 8 * K) if self.use_split else 0
        fit = max((sum(int(n) for n in cfg.fitting_neuron) * 3 + K) * F, 0) * min(1.0, self.fit_chunk / 1e6)
        return int(cfg.nnei * (4 + 12 + 3 + 4) * F + cfg.nnei * 4 + 3 * 4 * self.M * F + desc + fit)

    def evaluate_chunked(self, ext_coord, ext_type, numneigh, rows, mapping, nloc, chunks, atom_virial=False):
        """evaluate() in slabs of centre atoms: only one slab's intermediates are alive at a time (BASELINE config 3,
        4 M copper atoms with sel 512, needs 95 KB of them per atom: 380 GB).  `chunks` = [(a, b, perm, ranges, inv)]
        with the type partition of centre atoms a..b-1 (indices relative to a).  Forces / virial accumulate through
        dpb200_prod_force_virial_a_ex."""
        cfg = self.cfg
        nall = ext_type.numel()
        n_out = nloc if mapping is not None else nall
        dev = ext_coord.device
        force = torch.empty(n_out * 3, dtype=self.dtype, device=dev)
        virial = torch.empty(9, dtype=self.dtype, device=dev)
        av = torch.empty(n_out * 9, dtype=self.dtype, device=dev) if atom_virial else None
        e_atom = torch.empty(nloc, dtype=self.dtype, device=dev)
        coord_flat = ext_coord.reshape(-1)
        for ci, (a, b, perm, ranges, inv) in enumerate(chunks):
            em, dv, rij, nlist = ops.prod_env_mat_a(coord_flat, ext_type, numneigh, rows, self.davg, self.dstd, nloc, nall,
                                                    cfg.rcut, cfg.rcut_smth, cfg.sec, row_range=(a, b))
            if self.use_split:
                xyz, desc, row_exp = ops.tabulate_sections_desc(self.tables, self.infos, em, cfg.sec, self.M,
                                                                cfg.axis_neuron, 1.0 / cfg.nnei, desc_row=inv,
                                                                mode=self.desc_mode, nslice=self.desc_nslice, pad_rows=32,
                                                                flags=self.coef_flags)
                _, e_c, dy = self.energy_and_dy_split(xyz, desc, row_exp, perm, ranges)
                del desc
            else:
                xyz = ops.tabulate_sections_fwd(self.tables, self.infos, em, cfg.sec, self.M)
                _, e_c, dy = self.energy_and_dy(xyz, perm, ranges)
            net_deriv = ops.tabulate_sections_grad(self.tables, self.infos, em, dy, cfg.sec, self.M, flags=self.coef_flags)
            if mapping is not None:
                ops.use_nlist_map(nlist, mapping)
            ops.prod_force_virial_a_ex(force, virial, av, net_deriv, dv, rij, nlist, b - a, a, n_out, cfg.nnei,
                                       accumulate=ci > 0)
            e_atom[a:b] = e_c
            del em, dv, rij, nlist, xyz, dy, net_deriv
        return e_atom.sum(), force.reshape(-1, 3), virial, dict(atom_energy=e_atom, atom_virial=av, nlist=None)

    def evaluate(self, ext_coord, ext_type, numneigh, rows, mapping, nloc, type_perm, type_ranges, atom_virial=False,
                 fused=True, type_inv=None, firstneigh=None, max_nbor_size=None):
        """One force evaluation on an extended system. Returns (E, force[nloc,3], virial[9], extras).
        The raw list is either a dense `rows` block or `firstneigh` (device array of device row pointers)."""
        cfg = self.cfg
        nall = ext_type.numel()
        em, dv, rij, nlist = ops.prod_env_mat_a(ext_coord.reshape(-1), ext_type, numneigh, rows, self.davg, self.dstd,
                                                nloc, nall, cfg.rcut, cfg.rcut_smth, cfg.sec, firstneigh=firstneigh,
                                                max_nbor_size=max_nbor_size)
        if self.use_split:
            if type_inv is None:
                type_inv = torch.empty(nloc, dtype=torch.int32, device=type_perm.device)
                type_inv[type_perm] = torch.arange(nloc, dtype=torch.int32, device=type_perm.device)
            xyz, desc, row_exp = ops.tabulate_sections_desc(self.tables, self.infos, em, cfg.sec, self.M,
                                                            cfg.axis_neuron, 1.0 / cfg.nnei, desc_row=type_inv,
                                                            mode=self.desc_mode, nslice=self.desc_nslice, pad_rows=32,
                                                            flags=self.coef_flags)
            energy, e_atom, dy = self.energy_and_dy_split(xyz, desc, row_exp, type_perm, type_ranges)
            del desc
        else:
            xyz = ops.tabulate_sections_fwd(self.tables, self.infos, em, cfg.sec, self.M)
            energy, e_atom, dy = self.energy_and_dy(xyz, type_perm, type_ranges)
        net_deriv = ops.tabulate_sections_grad(self.tables, self.infos, em, dy, cfg.sec, self.M, flags=self.coef_flags)
        if mapping is not None:
            ops.use_nlist_map(nlist, mapping)
            n_out = nloc
        else:
            n_out = nall
        if fused:
            force, virial, av = ops.prod_force_virial_a(net_deriv, dv, rij, nlist, nloc, n_out, cfg.nnei,
                                                        atom_virial=atom_virial)
        else:
            force = ops.prod_force_a(net_deriv, dv, nlist, nloc, n_out, cfg.nnei)
            virial, av = ops.prod_virial_a(net_deriv, dv, rij, nlist, nloc, n_out, cfg.nnei)
        return energy, force.reshape(-1, 3), virial, dict(atom_energy=e_atom, atom_virial=av, nlist=nlist)


@dataclass
class NeighborState:
    nloc: int
    ext_type: torch.Tensor
    mapping: torch.Tensor
    shift: torch.Tensor  # ext_coord - coord[mapping] at build time
    numneigh: torch.Tensor
    rows: torch.Tensor
    type_perm: torch.Tensor
    type_ranges: list
    ago: int = 0
    map64: Optional[torch.Tensor] = None
    type_inv: Optional[torch.Tensor] = None  # int32 inverse of type_perm: descriptor row of atom i
    chunks: Optional[list] = None  # atom slabs [(a, b, perm, ranges, inv)] when the evaluation is chunked
    box: Optional[np.ndarray] = None  # cell the list was built for (host copy)
    ref_coord: Optional[torch.Tensor] = None  # coordinates the list was built from


def type_partition(atype: torch.Tensor, ntypes: int):
    perm = torch.argsort(atype.to(torch.int64), stable=True)
    counts = torch.bincount(atype.to(torch.int64), minlength=ntypes).tolist()
    ranges, a = [], 0
    for c in counts[:ntypes]:
        ranges.append((a, a + int(c)))
        a += int(c)
    return perm, ranges


class DeepPotB200:
    """Inference facade with the calling convention of deepmd.infer.DeepPot.eval: periodic cell,
    host arrays in and out.  The raw neighbour list is rebuilt every `nlist_every` evaluations with a
    `skin` (the reference MD set-up examples/water/lmp/in.lammps:7-8: neighbor 2.0 bin, every 10)."""

    def __init__(self, model: SeAModel, skin: float = 2.0, nlist_every: int = 10, use_graph: bool = True,
                 atom_chunk="auto"):
        self.model = model
        # centre atoms per slab of the evaluation: None = one slab, "auto" = one slab unless the per-atom
        # intermediates would not fit in 60 % of the device memory
        self.atom_chunk = atom_chunk
        self.skin = float(skin)
        self.nlist_every = int(nlist_every)
        # The steady-state step (everything between two list rebuilds) has static shapes and no host
        # decisions: it is captured once into a CUDA graph and replayed, so an MD step costs one graph
        # launch on the host instead of ~400 kernel launches.
        self.use_graph = bool(use_graph)
        self._graph = None
        self._graph_key = None
        self._graph_launches = 0
        self.state: Optional[NeighborState] = None
        self._cache = {}  # persistent device buffers of the neighbour-list rebuild
        self._pin_in = None
        self._pin_out = None
        self._atype_np = None

    def reset(self):
        self.state = None

    def _type_partition(self, atype: torch.Tensor):
        """Atom types do not change between rebuilds: sort them once per type tensor."""
        key = (atype.data_ptr(), atype.numel())
        hit = self._cache.get("perm")
        if hit is None or hit[0] != key:
            perm, ranges = type_partition(atype, self.model.cfg.ntypes)
            inv = torch.empty(perm.numel(), dtype=torch.int32, device=perm.device)
            inv[perm] = torch.arange(perm.numel(), dtype=torch.int32, device=perm.device)
            hit = (key, perm, ranges, inv)
            self._cache["perm"] = hit
            self._cache.pop("chunks", None)
        return hit[1], hit[2], hit[3]

    def _chunk_size(self, nloc: int):
        c = self.atom_chunk
        if c is None:
            return None
        if c == "auto":
            dev = self.model.device
            if dev.type != "cuda":
                return None
            total = torch.cuda.get_device_properties(dev).total_memory
            per = self.model.bytes_per_atom()
            if per * nloc <= 0.6 * total:
                return None
            c = max(1024, int(0.45 * total / per) // 1024 * 1024)
        c = int(c)
        return c if c < nloc else None

    def _chunks(self, atype: torch.Tensor):
        """Type partition of every slab of centre atoms (cached with the type tensor)."""
        nloc = atype.numel()
        c = self._chunk_size(nloc)
        if c is None:
            return None
        hit = self._cache.get("chunks")
        key = (atype.data_ptr(), nloc, c)
        if hit is None or hit[0] != key:
            out = []
            for a in range(0, nloc, c):
                b = min(nloc, a + c)
                perm, ranges = type_partition(atype[a:b], self.model.cfg.ntypes)
                inv = torch.empty(b - a, dtype=torch.int32, device=perm.device)
                inv[perm] = torch.arange(b - a, dtype=torch.int32, device=perm.device)
                out.append((a, b, perm, ranges, inv))
            hit = (key, out)
            self._cache["chunks"] = hit
        return hit[1]

    def build_neighbors(self, coord: torch.Tensor, atype: torch.Tensor, box) -> NeighborState:
        m = self.model
        rc = m.cfg.rcut + self.skin
        self.state = None  # the views below alias the cached buffers of the previous list
        nloc = atype.numel()
        c = ops._buf(self._cache, "norm_c", (nloc, 3), coord.dtype, coord.device)
        c.copy_(coord.reshape(-1, 3))
        ops.normalize_coord(c, box)
        ext_c, ext_t, mapping = ops.copy_coord(c, atype, box, rc, cache=self._cache)
        numneigh, rows = ops.build_nlist(ext_c, nloc, rc, ext_t, cache=self._cache)
        map64 = ops._buf(self._cache, "map64", (mapping.numel(),), torch.int64, coord.device)
        map64.copy_(mapping)
        shift = ops._buf(self._cache, "shift", (mapping.numel(), 3), coord.dtype, coord.device)
        torch.index_select(coord.reshape(-1, 3), 0, map64, out=shift)
        torch.sub(ext_c, shift, out=shift)
        perm, ranges, inv = self._type_partition(atype)
        ref = ops._buf(self._cache, "ref_coord", (nloc, 3), coord.dtype, coord.device)
        ref.copy_(coord.reshape(-1, 3))
        self.state = NeighborState(nloc, ext_t, mapping, shift, numneigh, rows, perm, ranges, map64=map64, type_inv=inv,
                                   chunks=self._chunks(atype), box=np.array(box, dtype=np.float64).reshape(9).copy(),
                                   ref_coord=ref)
        return self.state

    def _list_is_stale(self, coord: torch.Tensor, atype: torch.Tensor, box) -> bool:
        """The raw list (cutoff rcut + skin) and the ghost image shifts stay valid while the cell is the one they
        were built for and no atom has moved more than skin / 2 from its build-time position (the LAMMPS `check yes`
        criterion); `nlist_every` only caps the reuse (examples/water/lmp/in.lammps:7-8: every 10).  Independent
        frames handed to eval() therefore never see another frame's list."""
        st = self.state
        if st is None or st.nloc != atype.numel() or st.ago >= self.nlist_every:
            return True
        if not np.array_equal(st.box, np.asarray(box, dtype=np.float64).reshape(9)):
            return True
        if self.skin <= 0.0:
            return True
        d2 = (coord.reshape(-1, 3) - st.ref_coord).square_().sum(1).max()
        return bool(d2.item() > (0.5 * self.skin) ** 2)

    def _step(self, coord, atom_virial, fused):
        st = self.state
        ext_c = coord.reshape(-1, 3).index_select(0, st.map64).add_(st.shift)
        self._last_ext_coord = ext_c
        if st.chunks is not None:
            return self.model.evaluate_chunked(ext_c, st.ext_type, st.numneigh, st.rows, st.mapping, st.nloc, st.chunks,
                                               atom_virial=atom_virial)
        return self.model.evaluate(ext_c, st.ext_type, st.numneigh, st.rows, st.mapping, st.nloc, st.type_perm,
                                   st.type_ranges, atom_virial=atom_virial, fused=fused, type_inv=st.type_inv)

    def eval_device(self, coord: torch.Tensor, atype: torch.Tensor, box, atom_virial=False, fused=True):
        """Device tensors in, device tensors out (no host copies).  With `use_graph` the returned tensors
        live in the graph's memory pool and are overwritten by the next call."""
        st = self.state
        if self._list_is_stale(coord, atype, box):
            st = self.build_neighbors(coord, atype, box)
        st.ago += 1
        if not self.use_graph or not getattr(self.model, "graph_safe", True):
            # (a model whose launch shapes follow the neighbour counts of the step cannot be replayed from a graph)
            return self._step(coord, atom_virial, fused)
        key = (st.nloc, int(st.ext_type.numel()), int(st.rows.shape[1]), st.rows.data_ptr(), st.ext_type.data_ptr(),
               bool(atom_virial), bool(fused), coord.dtype, st.type_perm.data_ptr(),
               None if st.chunks is None else tuple(c[2].data_ptr() for c in st.chunks))
        if self._graph is None or self._graph_key != key:
            self._graph = None
            self._g_out = None
            self._g_coord = torch.empty_like(coord.reshape(-1, 3))
            self._g_coord.copy_(coord.reshape(-1, 3))
            self._step(self._g_coord, atom_virial, fused)  # eager warm-up (cuBLAS handles, attributes)
            torch.cuda.synchronize(coord.device)
            torch.cuda.empty_cache()
            from ._lib import lib

            n0 = lib().launch_count()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    self._g_out = self._step(self._g_coord, atom_virial, fused)
            except Exception as exc:  # keep working eagerly, but say so
                import warnings

                warnings.warn(f"dpb200: CUDA graph capture of the force step failed ({exc!r}); running eagerly")
                self.use_graph = False
                torch.cuda.synchronize(coord.device)
                return self._step(coord, atom_virial, fused)
            self._graph_launches = lib().launch_count() - n0
            self._graph, self._graph_key = graph, key
            # the graph reads these index tensors on every replay: keep them alive with it
            self._graph_refs = (st.type_perm, st.type_inv, st.chunks, st.rows, st.ext_type, st.numneigh, st.mapping,
                                st.map64, st.shift)
        self._g_coord.copy_(coord.reshape(-1, 3))
        self._graph.replay()
        from ._lib import lib

        lib().cdll.dpb200_count_replayed_launches(self._graph_launches)
        return self._g_out

    def compute_nlist(self, ext_coord: torch.Tensor, ext_type: torch.Tensor, nloc: int, numneigh: torch.Tensor,
                      rows: Optional[torch.Tensor] = None, firstneigh: Optional[torch.Tensor] = None,
                      max_nbor_size: Optional[int] = None, ago: int = 0, atom_virial: bool = False):
        """Device-resident hand-off of an MD code's own neighbour list: the counterpart of
        DeepPotPT::compute(..., nghost, InputNlist, ago) (source/api_cc/src/DeepPotPT.cc:171-377) and of the Kokkos
        pair style (source/lmp/pair_deepmd_kokkos.cpp) WITHOUT the per-step host staging the reference does there
        (host vectors -> padded host list -> H2D; forces D2H): every argument is a DEVICE tensor and so is every result.

        ext_coord [nall, 3], ext_type [nall] int32 (local atoms first, then ghosts; negative type = virtual atom),
        numneigh [nloc] int32, and the raw rows either as a dense row-major block `rows` [nloc, capacity] int32 or as
        `firstneigh`, an int64 device tensor of device row pointers (InputNlist layout, neighbor_list.h:20-57);
        `max_nbor_size` bounds numneigh (default: the row capacity / the largest count).  The list must contain every
        neighbour within rcut (a skin is fine: rows are re-formatted with the true cutoff every call).  `ago` == 0
        tells that atom types / the list changed since the previous call (type partition is recomputed).
        Returns (E, force [nall, 3] INCLUDING the forces on ghost atoms -- the caller's reverse communication folds
        them, as PairDeepMD::compute does (pair_deepmd.cpp:482-488) --, virial [9], extras)."""
        m = self.model
        ext_type = ext_type.to(torch.int32)
        key = (ext_type.data_ptr(), int(ext_type.numel()), int(nloc))
        hit = self._cache.get("nl_perm")
        if ago == 0 or hit is None or hit[0] != key:
            perm, ranges = type_partition(ext_type[:nloc], m.cfg.ntypes)
            inv = torch.empty(nloc, dtype=torch.int32, device=perm.device)
            inv[perm] = torch.arange(nloc, dtype=torch.int32, device=perm.device)
            hit = (key, perm, ranges, inv)
            self._cache["nl_perm"] = hit
        if rows is None and firstneigh is None:
            raise ValueError("dpb200: compute_nlist needs `rows` or `firstneigh`")
        return m.evaluate(ext_coord.reshape(-1, 3), ext_type, numneigh.to(torch.int32), rows, None, int(nloc), hit[1],
                          hit[2], atom_virial=atom_virial, type_inv=hit[3], firstneigh=firstneigh,
                          max_nbor_size=max_nbor_size)

    def eval(self, coords, cells, atom_types, atomic: bool = False):
        """coords [nframes, natoms*3], cells [nframes, 9], atom_types [natoms] (host arrays).
        Returns (energy [nframes,1], force [nframes,natoms,3], virial [nframes,9]) as numpy arrays
        (+ atom_energy, atom_virial when atomic)."""
        m = self.model
        dev = m.device
        coords = np.asarray(coords).reshape(-1, len(atom_types) * 3)
        cells = np.asarray(cells, dtype=np.float64).reshape(-1, 9)
        nf, nat = coords.shape[0], len(atom_types)
        np_dt = np.float64 if m.dtype == torch.float64 else np.float32
        at_np = np.ascontiguousarray(atom_types, dtype=np.int32)
        if self._pin_in is None or self._pin_in.numel() != nat * 3:
            self._pin_in = torch.empty(nat * 3, dtype=m.dtype).pin_memory()
            self._pin_out = torch.empty(nat * 3 + 10, dtype=m.dtype).pin_memory()
            self._atype_np = None
        if self._atype_np is None or not np.array_equal(self._atype_np, at_np):
            self._atype_np = at_np.copy()
            self._atype = torch.as_tensor(at_np).to(dev)
            self.state = None  # new types: new list, new type partition
        e_out = np.empty((nf, 1), np_dt)
        f_out = np.empty((nf, nat, 3), np_dt)
        v_out = np.empty((nf, 9), np_dt)
        aes, avs = [], []
        for f in range(nf):
            self._pin_in.copy_(torch.from_numpy(np.ascontiguousarray(coords[f], dtype=np_dt)))
            c = self._pin_in.to(dev, non_blocking=True)
            e, force, virial, ex = self.eval_device(c, self._atype, cells[f], atom_virial=atomic)
            out = torch.cat([force.reshape(-1), virial.reshape(-1), e.reshape(1)])
            self._pin_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            o = self._pin_out.numpy()
            # (torch's CPU copy is multi-threaded; a 37 MB numpy assignment is a single-threaded memcpy)
            torch.from_numpy(f_out[f].reshape(-1)).copy_(self._pin_out[: nat * 3])
            v_out[f] = o[nat * 3: nat * 3 + 9]
            e_out[f, 0] = o[nat * 3 + 9]
            if atomic:
                aes.append(ex["atom_energy"].cpu().numpy())
                avs.append(ex["atom_virial"].cpu().numpy().reshape(nat, 9))
        res = (e_out, f_out, v_out)
        if atomic:
            res = res + (np.stack(aes)[..., None], np.stack(avs))
        return res
