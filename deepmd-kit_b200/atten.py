"""Compressed se_atten (DPA-1, `se_atten_v2`: strip-mode type embedding, smooth, attn_layer = 0) energy / force /
virial evaluation — BASELINE config 5, examples/water/se_atten_compressible/input.json.

Model path of deepmd/pt/model/descriptor/se_atten.py:892-1016 with geometric + type-embedding compression
(dpa1.py:640-700), driven through the dpb200 operators:

    prod_env_mat_a with ONE type-agnostic section (sel = 120, neighbours ordered by distance only: f_type = 0)
    sw(r_ij)                                    smooth switch of every neighbour (se_atten.py: `sw`)
    two_embed = tt_full[center*(nt+1) + nei] * sw      type-pair gate, tt_full = strip net on the type embeddings
                                                       ((ntypes+1)^2 x M table, `type_embd_data` se_atten.py:666-720)
    moment = tabulate_fusion_se_atten(table, em_x, em, two_embed) / nnei        g * (1 + two_embed) folded in the op
    D = moment^T moment[:, :axis]  ++ type embedding of the centre atom   (dpa1.py:769-770 concat_output_tebd)
    one fitting net for all types (mixed types) + bias_atom_e[type]
    backward: tabulate_fusion_se_atten_grad -> dE/d(em) (prod_force_a / prod_virial_a) and dE/d(two_embed) ->
    dE/d(sw) -> the pair force through the switch function.

`SeAttenModel` offers the interface DeepPotB200 expects from a model (cfg, evaluate, bytes_per_atom), so the same
facade, neighbour-list management and CUDA-graph replay serve both descriptors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .compress import EmbeddingNet, build_table, env_mat_range
from .model import FittingNet


@dataclass
class SeAttenConfig:
    """examples/water/se_atten_compressible/input.json."""
    ntypes: int = 2
    nsel: int = 120
    rcut: float = 6.0
    rcut_smth: float = 0.5
    neuron: Sequence[int] = (25, 50, 100)
    axis_neuron: int = 16
    tebd_dim: int = 8
    fitting_neuron: Sequence[int] = (240, 240, 240)
    fitting_resnet_dt: bool = True
    stats: Sequence[Sequence[float]] = ((0.05033, 0.13984, 0.08580), (0.04810, 0.12388, 0.07672))
    seed: int = 1
    stride0: float = 0.01
    stride1: float = 0.1
    extrapolate: float = 5.0
    min_nbor_dist: float = 0.9
    # neighbour-gated self-attention on the per-neighbour embedding (DPA-1 proper; 0 = the compressible model).
    # Reference defaults of deepmd/pt/model/descriptor/dpa1.py: attn 128, attn_dotr, normalised q / k / v, one head,
    # scaling (attn * scaling_factor)^-1/2, smooth shift 20, layer-norm eps 1e-5.
    attn_layer: int = 0
    attn: int = 128
    attn_dotr: bool = True
    attn_normalize: bool = True
    scaling_factor: float = 1.0
    ln_eps: float = 1e-5
    attnw_shift: float = 20.0

    @property
    def sel(self):
        return (self.nsel,)

    @property
    def nnei(self) -> int:
        return int(self.nsel)

    @property
    def sec(self):
        return [0, int(self.nsel)]


def _mlp_init(widths, seed):
    g = torch.Generator().manual_seed(seed)
    ws, bs = [], []
    for n_in, n_out in zip(widths[:-1], widths[1:]):
        ws.append(torch.empty(n_in, n_out, dtype=torch.float64).normal_(0.0, 1.0 / math.sqrt(n_in + n_out), generator=g))
        bs.append(torch.empty(n_out, dtype=torch.float64).normal_(0.0, 1.0, generator=g))
    return ws, bs


def _mlp_tanh_resnet(x, ws, bs):
    """tanh MLP with the embedding-net skip rules (same width: + x, doubled width: + concat(x, x))."""
    for w, b in zip(ws, bs):
        y = torch.tanh(x @ w + b)
        if w.shape[1] == w.shape[0]:
            y = y + x
        elif w.shape[1] == 2 * w.shape[0]:
            y = y + torch.cat([x, x], 1)
        x = y
    return x


def switch_and_derivative(r: torch.Tensor, rmin: float, rmax: float):
    """spline5 switch (source/lib/include/switcher.h:61-84) and d sw / d r, elementwise."""
    uu = ((r - rmin) / (rmax - rmin)).clamp(0.0, 1.0)
    sw = uu * uu * uu * (-6.0 * uu * uu + 15.0 * uu - 10.0) + 1.0
    dsw = (3.0 * uu * uu * (-6.0 * uu * uu + 15.0 * uu - 10.0) + uu * uu * uu * (-12.0 * uu + 15.0)) / (rmax - rmin)
    inside = (r >= rmin) & (r < rmax)
    sw = torch.where(r < rmin, torch.ones_like(sw), torch.where(inside, sw, torch.zeros_like(sw)))
    dsw = torch.where(inside, dsw, torch.zeros_like(dsw))
    return sw, dsw


# The dense products of the attention layers go to the library (cuBLAS through torch): plain GEMMs.  They are module
# level names so that bench.py can time them as one row ("attn_library_gemm") next to the dpb200 stage kernels.
def _addmm(bias, a, b):
    return torch.addmm(bias, a, b)


def _bmm(a, b):
    return torch.bmm(a, b)


def _mm(a, b):
    return torch.mm(a, b)


class SeAttenModel:
    """Random-init compressed se_atten_v2 model (weights of the named architecture; geometric table through the
    restated `dp compress`, type-pair table through the strip net)."""

    def __init__(self, cfg: SeAttenConfig, dtype=torch.float64, device="cuda", weights: Optional[dict] = None):
        """weights (optional; tests/golden/dpa1_strip.json "weights", taken from the reference's DescrptDPA1):
        embed / strip = [{w, b, resnet}] of the geometric net and of the two-side strip net, tebd = the type-embedding
        table [(ntypes + 1), tebd_dim] with the zero padding row last.  Without it: random init of the same shapes."""
        self.cfg = cfg
        self.dtype = dtype
        self.device = torch.device(device)
        nt, nnei = cfg.ntypes, cfg.nnei
        davg = np.zeros((nt, nnei, 4))
        dstd = np.ones((nt, nnei, 4))
        for t, (a0, s0, s1) in enumerate(cfg.stats):
            davg[t, :, 0] = a0
            dstd[t, :, 0] = s0
            dstd[t, :, 1:] = s1
        self.davg = torch.as_tensor(davg.reshape(nt, -1), dtype=dtype, device=self.device)
        self.dstd = torch.as_tensor(dstd.reshape(nt, -1), dtype=dtype, device=self.device)
        # one type-agnostic geometric embedding net -> one table over the range of all centre types
        self.embed = EmbeddingNet(cfg.neuron, cfg.seed)
        if weights is not None:
            self.embed.weights = [torch.as_tensor(np.asarray(l["w"], np.float64)) for l in weights["embed"]]
            self.embed.biases = [torch.as_tensor(np.asarray(l["b"], np.float64)) for l in weights["embed"]]
        lower, upper = env_mat_range(davg[:, 0, :], dstd[:, 0, :], cfg.min_nbor_dist, cfg.rcut_smth, cfg.rcut)
        ll, uu = float(lower.min()), float(upper.max())
        self.table64 = build_table(self.embed, ll, uu, cfg.stride0, cfg.stride1, cfg.extrapolate)
        self.table = self.table64.to(self.device, dtype).contiguous()
        self.info = torch.tensor([ll, uu, uu * cfg.extrapolate, cfg.stride0, cfg.stride1, -1.0], dtype=dtype)
        self.M = int(cfg.neuron[-1])
        # type embedding (ntypes + 1 rows, the last one is the zero padding row) and the two-side strip net:
        # tt_full[center * (nt + 1) + nei] = strip([tebd[nei], tebd[center]])   (se_atten.py:698-711)
        g = torch.Generator().manual_seed(cfg.seed + 7)
        tebd = torch.zeros(nt + 1, cfg.tebd_dim, dtype=torch.float64)
        tebd[:nt] = torch.empty(nt, cfg.tebd_dim, dtype=torch.float64).normal_(0.0, 1.0, generator=g)
        ws, bs = _mlp_init([2 * cfg.tebd_dim] + list(cfg.neuron), cfg.seed + 13)
        if weights is not None:
            tebd = torch.as_tensor(np.asarray(weights["tebd"], np.float64)).reshape(nt + 1, cfg.tebd_dim)
            ws = [torch.as_tensor(np.asarray(l["w"], np.float64)) for l in weights["strip"]]
            bs = [torch.as_tensor(np.asarray(l["b"], np.float64)) for l in weights["strip"]]
        nei = tebd.view(1, nt + 1, -1).expand(nt + 1, nt + 1, -1)
        cen = tebd.view(nt + 1, 1, -1).expand(nt + 1, nt + 1, -1)
        tt = _mlp_tanh_resnet(torch.cat([nei, cen], -1).reshape(-1, 2 * cfg.tebd_dim), ws, bs)
        self.tebd = tebd.to(self.device, dtype)
        self.tt_full = tt.to(self.device, dtype).contiguous()  # [(nt+1)^2, M]
        # fitting net input = [D (M*axis), tebd(centre)] zero-padded to a multiple of 16 (the int8 tensor-core
        # GEMMs read 16-byte aligned operand rows; the padded inputs are zeros, so their weight rows are inert)
        self.dim_d = self.M * cfg.axis_neuron
        self.dim_in = (self.dim_d + cfg.tebd_dim + 15) // 16 * 16
        self.fit = FittingNet(self.dim_in, cfg.fitting_neuron, cfg.fitting_resnet_dt, cfg.seed + 101, dtype, self.device)
        self.bias_atom_e = torch.zeros(nt, dtype=dtype, device=self.device)
        if weights is not None and "fit" in weights:
            # the reference's fitting net (deepmd/dpmodel/fitting: one net for all types, input [D | tebd(centre)]);
            # the first weight matrix gets zero rows for the padded input columns
            layers = []
            for li, l in enumerate(weights["fit"]["layers"]):
                w = torch.as_tensor(np.asarray(l["w"], np.float64))
                if li == 0 and w.shape[0] < self.dim_in:
                    w = torch.cat([w, torch.zeros(self.dim_in - w.shape[0], w.shape[1], dtype=torch.float64)], 0)
                b = torch.as_tensor(np.asarray(l["b"], np.float64))
                idt = None if l.get("idt") is None else torch.as_tensor(np.asarray(l["idt"], np.float64))
                layers.append((w.to(self.device, dtype).contiguous(), b.to(self.device, dtype),
                               None if idt is None else idt.to(self.device, dtype)))
            self.fit.layers = layers
            hw = torch.as_tensor(np.asarray(weights["fit"]["head"]["w"], np.float64)).reshape(-1, 1)
            hb = torch.as_tensor(np.asarray(weights["fit"]["head"]["b"], np.float64)).reshape(1)
            self.fit.head = (hw.to(self.device, dtype), hb.to(self.device, dtype))
            self.bias_atom_e = torch.as_tensor(np.asarray(weights["bias_atom_e"], np.float64)).to(self.device, dtype)
        self.fit_chunk = 1 << 17
        # two_embed [atoms * nnei, M] is 96 KB per atom in fp64 (50 GB for the 526 848-atom box): the gated table
        # operator runs over slabs of this many centre atoms, the gate being recomputed for the backward
        self.tab_chunk = 1 << 16
        # default: the pair-indexed gate entry points (dpb200_tabulate_fusion_se_atten_gate*), which form
        # tt_full[pair] * sw inside the kernel; False = the reference-schema op fed with the materialised tensor
        self.use_gate = True
        # compressed table coefficients in the gated backward (same host-side check as SeAModel; fp64 only here)
        import os

        self.coef_flags = 0
        if self.device.type == "cuda" and dtype == torch.float64 and os.environ.get("DPB200_TAB_COMPRESS", "1") != "0":
            self.coef_flags = int(ops.compressed_coef_flags(self.table64, self.info))
        # attention layers: in_proj [M, 3 attn] + bias, out_proj [attn, M] + bias, layer norm scale / shift [M]
        # (weights["attn"] = [{in_w, in_b, out_w, out_b, ln_w, ln_b}] from the reference's NeighborGatedAttention)
        self.attn_layers = []
        self.attn_scaling = float((cfg.attn * cfg.scaling_factor) ** -0.5)
        self.attn_chunk = 4096  # centre atoms per slab: ~2 MB of saved activations per atom and layer pair in fp64
        # Empty neighbour slots trail the formatted list and are all alike (s = -davg / dstd, sw = 0, rhat = 0): a slab
        # is evaluated on its first n_eff slots, n_eff = the slab's largest neighbour count + 1 (rounded up to 8), the
        # last of which stands for every omitted empty slot -- (nnei - n_eff + 1)-fold in the moment, and through
        # `nnei_full` in the softmax denominators.  The slot count follows the step's neighbour counts (one host read
        # per slab), so the step is not replayed from a CUDA graph.
        self.attn_compact = True
        self.graph_safe = cfg.attn_layer == 0
        for li in range(cfg.attn_layer):
            if weights is not None and "attn" in weights:
                lay = {k: torch.as_tensor(np.asarray(v, np.float64)) for k, v in weights["attn"][li].items()}
            else:
                ws_, bs_ = _mlp_init([self.M, 3 * cfg.attn], cfg.seed + 31 + 2 * li)
                wo_, bo_ = _mlp_init([cfg.attn, self.M], cfg.seed + 32 + 2 * li)
                lay = dict(in_w=ws_[0], in_b=bs_[0], out_w=wo_[0], out_b=bo_[0],
                           ln_w=torch.ones(self.M, dtype=torch.float64), ln_b=torch.zeros(self.M, dtype=torch.float64))
            assert lay["in_w"].shape == (self.M, 3 * cfg.attn) and lay["out_w"].shape == (cfg.attn, self.M)
            lay = {k: v.to(self.device, dtype).contiguous() for k, v in lay.items()}
            lay["in_wt"] = lay["in_w"].t().contiguous()
            lay["out_wt"] = lay["out_w"].t().contiguous()
            self.attn_layers.append(lay)
        self.nslice = 6 if dtype == torch.float64 else 4  # int8 digit slices of the fitting-net operands
        self.use_tc = bool(self.device.type == "cuda" and self.fit.prepare_tc(self.nslice))
        # lower bound of the descriptor rows' exponent so that the appended type embedding fits: |tebd| < 2^(E-1)
        self.tebd_exp = int(math.floor(math.log2(max(float(self.tebd.abs().max()), 1e-300)))) + 2

    def bytes_per_atom(self) -> int:
        """Per-atom intermediates alive across the whole evaluation (two_embed and its gradient are only ever
        materialised for one slab of `tab_chunk` atoms at a time)."""
        F = 8 if self.dtype == torch.float64 else 4
        nnei, M = self.cfg.nnei, self.M
        return int(nnei * (4 + 12 + 3 + 4 + 3) * F + nnei * 4 + 3 * 4 * M * F)

    def gate_scalars(self, ext_type, nlist, rij, a, b):
        """Centre atoms a..b-1: (pair int32 [(b-a), nnei] row of tt_full, sw, sw', r) of every neighbour slot."""
        cfg = self.cfg
        nt1 = cfg.ntypes + 1
        nl = nlist[a:b]
        pad = nl < 0
        nei_t = ext_type.to(torch.int64)[nl.clamp_min(0).to(torch.int64)]
        nei_t = torch.where(pad | (nei_t < 0), torch.full_like(nei_t, cfg.ntypes), nei_t)
        cen_t = ext_type[a:b].to(torch.int64)
        cen_t = torch.where(cen_t < 0, torch.full_like(cen_t, cfg.ntypes), cen_t)
        pair = (cen_t.view(-1, 1) * nt1 + nei_t)
        r = rij[a:b].reshape(b - a, cfg.nnei, 3).norm(dim=-1)
        sw, dsw = switch_and_derivative(r, cfg.rcut_smth, cfg.rcut)
        sw = torch.where(pad, torch.zeros_like(sw), sw)
        dsw = torch.where(pad, torch.zeros_like(dsw), dsw)
        return pair, sw, dsw, r

    def gate(self, ext_type, nlist, rij, a, b):
        """Materialised two_embed [(b-a)*nnei, M] = tt_full[pair] * sw (the reference op's operand)."""
        pair, sw, dsw, r = self.gate_scalars(ext_type, nlist, rij, a, b)
        pair = pair.reshape(-1)
        two = self.tt_full.index_select(0, pair) * sw.reshape(-1, 1)
        return two, pair, dsw, r

    def evaluate(self, ext_coord, ext_type, numneigh, rows, mapping, nloc, type_perm=None, type_ranges=None,
                 atom_virial=False, fused=True, type_inv=None):
        """One force evaluation on an extended system: (E, force[n_out, 3], virial[9], extras)."""
        cfg = self.cfg
        if cfg.attn_layer > 0:
            return self._evaluate_attn(ext_coord, ext_type, numneigh, rows, mapping, nloc, atom_virial)
        nall = ext_type.numel()
        nnei, M = cfg.nnei, self.M
        f_type = torch.zeros_like(ext_type)  # one section: the list is ordered by distance only
        em, dv, rij, nlist = ops.prod_env_mat_a(ext_coord.reshape(-1), ext_type, numneigh, rows, self.davg, self.dstd, nloc,
                                                nall, cfg.rcut, cfg.rcut_smth, cfg.sec, f_type=f_type)
        em3 = em.reshape(nloc, nnei, 4)
        inv = 1.0 / nnei
        ctype = ext_type[:nloc].to(torch.int64)
        ctype_e = torch.where(ctype < 0, torch.full_like(ctype, cfg.ntypes), ctype)
        desc = row_exp = None
        if self.use_gate:
            pair32, sw, dswr = ops.se_atten_gate_scalars(nlist, ext_type, rij, nloc, nnei, cfg.ntypes, cfg.rcut_smth,
                                                         cfg.rcut)
            em_x = em3[:, :, 0].reshape(-1, 1).contiguous()
            if self.use_tc and self.dtype == torch.float64:
                # the warp that finishes an atom writes its descriptor straight as the int8 operand of the first fitting
                # GEMM; the centre type embedding is appended behind it at the same row exponent
                xyz, desc, row_exp = ops.tabulate_fusion_se_atten_gate_desc(
                    self.table, self.info, em_x, em3, self.tt_full, pair32, sw, M, cfg.axis_neuron, inv, self.dim_in,
                    self.nslice, self.tebd_exp, pad_rows=32, flags=self.coef_flags)
                ops.fit_slice_cols(desc, self.dim_in, self.dim_d, self.nslice, row_exp, self.tebd,
                                   idx=ctype_e.to(torch.int32))
            else:
                xyz = ops.tabulate_fusion_se_atten_gate(self.table, self.info, em_x, em3, self.tt_full, pair32, sw, M,
                                                        flags=self.coef_flags)
        else:
            xyz = torch.empty((nloc, 4, M), dtype=self.dtype, device=em.device)
            for a in range(0, nloc, self.tab_chunk):
                b = min(nloc, a + self.tab_chunk)
                two, _, _, _ = self.gate(ext_type, nlist, rij, a, b)
                em_xc = em3[a:b, :, 0].reshape(-1, 1).contiguous()
                xyz[a:b] = ops.tabulate_fusion_se_a(self.table, self.info, em_xc, em3[a:b], M, two_embed=two, is_sorted=True)
                del two
        # descriptor + centre type embedding -> fitting net -> dE/dD
        e_atom = torch.empty(nloc, dtype=self.dtype, device=em.device)
        dy = torch.empty_like(xyz)
        for c0 in range(0, nloc, self.fit_chunk):
            c1 = min(nloc, c0 + self.fit_chunk)
            if desc is not None:
                e, gd = self.fit.forward_backward_tc(desc[c0:c1], row_exp[c0:c1], c1 - c0, grad_cols=self.dim_d)
            else:
                g1 = torch.zeros((c1 - c0, self.dim_in), dtype=self.dtype, device=em.device)
                g1[:, :self.dim_d] = ops.se_a_descriptor(xyz[c0:c1], cfg.axis_neuron, inv)
                g1[:, self.dim_d:self.dim_d + cfg.tebd_dim] = self.tebd.index_select(0, ctype_e[c0:c1])
                if self.use_tc:
                    xs, ex = ops.split_i8_rows(g1.to(torch.float64), self.nslice)
                    e, gd = self.fit.forward_backward_tc(xs, ex, c1 - c0, grad_cols=self.dim_d)
                else:
                    e, gd = self.fit.forward_backward(g1)
                del g1
            e_atom[c0:c1] = e + self.bias_atom_e.index_select(0, ctype_e[c0:c1].clamp_max(cfg.ntypes - 1))
            # (the descriptor backward reads the first dim_d columns of the dim_in-wide gradient rows in place)
            dy[c0:c1] = ops.se_a_descriptor_grad(gd, xyz[c0:c1], cfg.axis_neuron, inv) if gd.shape[1] == self.dim_d \
                else ops.se_a_descriptor_grad(gd[:, :self.dim_d].contiguous(), xyz[c0:c1], cfg.axis_neuron, inv)
            del gd
        del desc
        # dE/d(sw_ij) = sum_k dE/d(two_embed)_ijk * tt_full[pair]_k ; pair force through the switch:
        # dE/dr_j = q * sw'(r) * (r_j - r_i) / r  (and the opposite on the centre atom)
        if self.use_gate:
            # the switch-path force -(q * sw'/r) r_ij rides in the force / virial kernel: nothing per pair is
            # materialised between the table backward and the scatter
            _, gem, q = ops.tabulate_fusion_se_atten_gate_grad(self.table, self.info, em_x, em3, self.tt_full, pair32, sw,
                                                               dy, M, fuse_x=True, flags=self.coef_flags)
            if mapping is not None:
                ops.use_nlist_map(nlist, mapping)
            n_out = nloc if mapping is not None else nall
            force, virial, av = ops.prod_force_virial_a_pair(gem.reshape(nloc, -1), dv, rij, nlist, q, dswr, nloc, n_out,
                                                             nnei, atom_virial=atom_virial)
            return e_atom.sum(), force.reshape(-1, 3), virial, dict(atom_energy=e_atom, atom_virial=av, nlist=nlist)
        else:
            net_deriv = torch.empty((nloc, nnei, 4), dtype=self.dtype, device=em.device)
            vec = torch.empty((nloc, nnei, 3), dtype=self.dtype, device=em.device)
            for a in range(0, nloc, self.tab_chunk):
                b = min(nloc, a + self.tab_chunk)
                two, pair, dsw, r = self.gate(ext_type, nlist, rij, a, b)
                em_xc = em3[a:b, :, 0].reshape(-1, 1).contiguous()
                gx, gem, gtwo = ops.tabulate_fusion_se_a_grad(self.table, self.info, em_xc, em3[a:b], dy[a:b], M,
                                                              two_embed=two, is_sorted=True)
                nd = gem.reshape(b - a, nnei, 4)
                nd[:, :, 0] += gx.reshape(b - a, nnei)
                net_deriv[a:b] = nd
                q = (gtwo * self.tt_full.index_select(0, pair)).sum(1).reshape(b - a, nnei)
                coef = torch.where(r > 0, q * dsw / r.clamp_min(1e-30), torch.zeros_like(q))
                vec[a:b] = coef.unsqueeze(-1) * rij[a:b].reshape(b - a, nnei, 3)
                del gtwo, two, gem, gx
        if mapping is not None:
            ops.use_nlist_map(nlist, mapping)
            n_out = nloc
        else:
            n_out = nall
        force, virial, av = ops.prod_force_virial_a(net_deriv.reshape(nloc, -1), dv, rij, nlist, nloc, n_out, nnei,
                                                    atom_virial=atom_virial)
        force = force.reshape(-1, 3)
        idx = nlist.clamp_min(0).to(torch.int64).reshape(-1)
        force.index_add_(0, idx, -vec.reshape(-1, 3))  # (padded slots carry vec = 0)
        force[:nloc] += vec.sum(1)
        # virial of the switch path: - dE/dr_j (x) r_ij per pair (symmetric: the force is along r_ij)
        virial = virial - torch.einsum("pi,pj->ij", vec.reshape(-1, 3), rij.reshape(-1, 3)).reshape(9)
        if atom_virial and av is not None:
            av = av.reshape(-1, 9)
            for a in range(0, nloc, self.tab_chunk):
                b = min(nloc, a + self.tab_chunk)
                w = -(vec[a:b].reshape(-1, 3, 1) * rij[a:b].reshape(-1, 1, 3))
                av.index_add_(0, idx[a * nnei:b * nnei], w.reshape(-1, 9))
            av = av.reshape(-1)
        return e_atom.sum(), force, virial, dict(atom_energy=e_atom, atom_virial=av, nlist=nlist)

    # ---------------------------------------------------------------------------------------- attention layers
    def attention_forward(self, x, sw, rhat, keep=True, nnei_full=None):
        """NeighborGatedAttention (se_atten.py:1058-1447) on x [B, nnei, M]: per layer in_proj -> normalised q, k, v
        -> gated softmax weights -> A v -> out_proj -> residual + layer norm.  Dense products on the library, the stages
        between them in csrc/attn_layers.cu.  Returns (x_out, saved activations for attention_backward)."""
        cfg = self.cfg
        B, n, M = x.shape
        h = cfg.attn
        saved = []
        for lay in self.attn_layers:
            qkv = _addmm(lay["in_b"], x.view(-1, M), lay["in_w"])
            inv = ops.attn_qkv_normalize(qkv, h, self.attn_scaling, cfg.attn_normalize)
            q3 = qkv.view(B, n, 3 * h)
            S = _bmm(q3[:, :, :h], q3[:, :, h:2 * h].transpose(1, 2))
            P, A = ops.attn_weights(S, sw, rhat, cfg.attnw_shift, cfg.attn_dotr, nnei_full=nnei_full)
            O = _bmm(A, q3[:, :, 2 * h:])
            Y = _addmm(lay["out_b"], O.view(-1, h), lay["out_w"])
            del O
            xn, zhat, rstd = ops.attn_residual_layernorm(x.view(-1, M), Y, lay["ln_w"], lay["ln_b"], cfg.ln_eps)
            if keep:
                saved.append((qkv, inv, S, P, A, zhat, rstd))
            x = xn.view(B, n, M)
        return x, saved

    def attention_backward(self, dx, saved, sw, rhat, d_sw, d_rhat):
        """Transpose of attention_forward: dE/d(x_out) [B, nnei, M] -> dE/d(x_in); dE/d(sw) and dE/d(rhat) of all
        layers are accumulated into d_sw / d_rhat."""
        cfg = self.cfg
        B, n, M = dx.shape
        h = cfg.attn
        for lay in reversed(self.attn_layers):
            qkv, inv, S, P, A, zhat, rstd = saved.pop()
            q3 = qkv.view(B, n, 3 * h)
            dz = ops.attn_residual_layernorm_grad(dx.reshape(-1, M), zhat, rstd, lay["ln_w"])  # = dY and the skip
            dO = _mm(dz, lay["out_wt"]).view(B, n, h)
            dqkv = torch.empty((B, n, 3 * h), dtype=dx.dtype, device=dx.device)
            dA = _bmm(dO, q3[:, :, 2 * h:].transpose(1, 2))
            dqkv[:, :, 2 * h:] = _bmm(A.transpose(1, 2), dO)
            del dO, A
            dS = ops.attn_weights_grad(dA, P, S, sw, rhat, d_sw, d_rhat, cfg.attnw_shift, cfg.attn_dotr)
            dqkv[:, :, :h] = _bmm(dS, q3[:, :, h:2 * h])
            dqkv[:, :, h:2 * h] = _bmm(dS.transpose(1, 2), q3[:, :, :h])
            del dS, dA, P, S
            ops.attn_qkv_normalize_grad(dqkv, qkv, inv, h, self.attn_scaling, cfg.attn_normalize)
            dx = _addmm(dz, dqkv.view(-1, 3 * h), lay["in_wt"]).view(B, n, M)
            del dqkv, qkv, dz
        return dx

    def _evaluate_attn(self, ext_coord, ext_type, numneigh, rows, mapping, nloc, atom_virial=False):
        """attn_layer > 0: the per-neighbour embedding is materialised slab by slab (attn_chunk centre atoms): table ->
        g2 -> attention layers -> moment -> descriptor -> fitting net -> and all the way back inside the slab, so
        that only net_deriv [nloc, nnei, 4] and dE/d(sw) [nloc, nnei] outlive it."""
        cfg = self.cfg
        nall = ext_type.numel()
        nnei, M = cfg.nnei, self.M
        f_type = torch.zeros_like(ext_type)
        em, dv, rij, nlist = ops.prod_env_mat_a(ext_coord.reshape(-1), ext_type, numneigh, rows, self.davg, self.dstd, nloc,
                                                nall, cfg.rcut, cfg.rcut_smth, cfg.sec, f_type=f_type)
        em3 = em.reshape(nloc, nnei, 4)
        inv = 1.0 / nnei
        ctype = ext_type[:nloc].to(torch.int64)
        ctype_e = torch.where(ctype < 0, torch.full_like(ctype, cfg.ntypes), ctype)
        pair32, sw, dswr = ops.se_atten_gate_scalars(nlist, ext_type, rij, nloc, nnei, cfg.ntypes, cfg.rcut_smth, cfg.rcut)
        e_atom = torch.empty(nloc, dtype=self.dtype, device=em.device)
        net_deriv = torch.empty((nloc, nnei, 4), dtype=self.dtype, device=em.device)
        q_sw = torch.zeros((nloc, nnei), dtype=self.dtype, device=em.device)
        for a in range(0, nloc, self.attn_chunk):
            b = min(nloc, a + self.attn_chunk)
            n_eff = nnei
            if a == 0:
                self.last_n_eff = []  # slots evaluated per slab (bench.py sizes the rooflines with it)
            if self.attn_compact:
                most = int((nlist[a:b] >= 0).sum(1).max().item())
                n_eff = min(nnei, (most + 1 + 7) // 8 * 8)
            self.last_n_eff.append(n_eff)
            if n_eff < nnei:
                em_c, sw_c, pair_c = em3[a:b, :n_eff].contiguous(), sw[a:b, :n_eff].contiguous(), \
                    pair32[a:b, :n_eff].contiguous()
                em_w = em_c.clone()
                em_w[:, -1, :] *= float(nnei - n_eff + 1)  # the kept empty slot stands for all of them in the moment
            else:
                em_c, sw_c, pair_c = em3[a:b], sw[a:b], pair32[a:b]
                em_w = em_c
            x0, gs, dgs = ops.se_atten_embed(self.table, self.info, em_c, self.tt_full, pair_c, sw_c, M)
            rhat, rinv = ops.se_atten_rhat(em_c)
            x, saved = self.attention_forward(x0, sw_c, rhat, nnei_full=nnei)
            del x0
            xyz = _bmm(em_w.transpose(1, 2), x)  # [B, 4, M], the unscaled moment (se_atten.py:1012)
            g1 = torch.zeros((b - a, self.dim_in), dtype=self.dtype, device=em.device)
            g1[:, :self.dim_d] = ops.se_a_descriptor(xyz, cfg.axis_neuron, inv)
            g1[:, self.dim_d:self.dim_d + cfg.tebd_dim] = self.tebd.index_select(0, ctype_e[a:b])
            if self.use_tc:
                xs, ex = ops.split_i8_rows(g1.to(torch.float64), self.nslice)
                e, gd = self.fit.forward_backward_tc(xs, ex, b - a, grad_cols=self.dim_d)
                del xs
            else:
                e, gd = self.fit.forward_backward(g1)
            del g1
            e_atom[a:b] = e + self.bias_atom_e.index_select(0, ctype_e[a:b].clamp_max(cfg.ntypes - 1))
            dy = ops.se_a_descriptor_grad(gd, xyz, cfg.axis_neuron, inv) if gd.shape[1] == self.dim_d \
                else ops.se_a_descriptor_grad(gd[:, :self.dim_d].contiguous(), xyz, cfg.axis_neuron, inv)
            dy = dy.to(self.dtype)
            del gd
            d_em = _bmm(x, dy.transpose(1, 2)).contiguous()  # through the moment's em factor
            dx = _bmm(em_w, dy)
            del x, dy
            d_rhat = torch.zeros((b - a, n_eff, 3), dtype=self.dtype, device=em.device)
            d_sw = q_sw[a:b] if n_eff == nnei else torch.zeros((b - a, n_eff), dtype=self.dtype, device=em.device)
            dx = self.attention_backward(dx, saved, sw_c, rhat, d_sw, d_rhat)
            ops.se_atten_embed_grad(d_em, d_sw, dx, gs, dgs, self.tt_full, pair_c, sw_c)
            ops.se_atten_rhat_grad(d_em, d_rhat, rhat, rinv)
            if n_eff == nnei:
                net_deriv[a:b] = d_em
            else:  # (the omitted slots are empty: their environment derivative is zero whatever stands here)
                net_deriv[a:b, :n_eff] = d_em
                net_deriv[a:b, n_eff:] = 0
                q_sw[a:b, :n_eff] = d_sw
            del dx, gs, dgs, d_em, d_rhat, d_sw
        if mapping is not None:
            ops.use_nlist_map(nlist, mapping)
        n_out = nloc if mapping is not None else nall
        force, virial, av = ops.prod_force_virial_a_pair(net_deriv.reshape(nloc, -1), dv, rij, nlist, q_sw, dswr, nloc,
                                                         n_out, nnei, atom_virial=atom_virial)
        return e_atom.sum(), force.reshape(-1, 3), virial, dict(atom_energy=e_atom, atom_virial=av, nlist=nlist)
