"""TEST INFRASTRUCTURE ONLY — CPU checker of the se_atten (DPA-1 strip / smooth) model path; `evaluate` = the compressed
attn_layer = 0 model, `evaluate_layers` (end of file) = the model with attention layers:
the reference's CPU operators (oracle.cpu.CpuLib: prod_env_mat_a_cpu with a type-agnostic format type,
tabulate_fusion_se_a_cpu / _grad_cpu with two_embed, prod_force_a_cpu, prod_virial_a_cpu) composed exactly as
deepmd/pt/model/descriptor/se_atten.py:892-1016 + dpa1.py:755-770 compose the torch ops, with the parts that the
reference leaves to torch autograd (type-pair gate x switch function, descriptor algebra, fitting net) done by torch
autograd on the CPU here as well.  `model` is a deepmd_kit_b200.atten.SeAttenModel living on the CPU (weights and
tables only)."""
from __future__ import annotations

import numpy as np
import torch


class _Tab(torch.autograd.Function):
    """tabulate_fusion_se_atten through the CPU library, differentiable in em_x / em / two_embed
    (source/op/pt/tabulate_multi_device.cc:964-1151)."""

    @staticmethod
    def forward(ctx, lib, table, info, em_x, em, two, M):
        ctx.lib, ctx.table, ctx.info, ctx.M = lib, table, info, M
        ctx.save_for_backward(em_x, em, two)
        out = lib.tabulate_fusion_se_a(table, info, em_x.numpy(), em.numpy(), M, two_embed=two.numpy(), is_sorted=True)
        return torch.as_tensor(out)

    @staticmethod
    def backward(ctx, dy):
        em_x, em, two = ctx.saved_tensors
        gx, gem, gtwo = ctx.lib.tabulate_fusion_se_a_grad(ctx.table, ctx.info, em_x.numpy(), em.numpy(),
                                                          np.ascontiguousarray(dy.numpy()), ctx.M, two_embed=two.numpy(),
                                                          is_sorted=True)
        return (None, None, None, torch.as_tensor(gx).reshape(em_x.shape), torch.as_tensor(gem).reshape(em.shape),
                torch.as_tensor(gtwo).reshape(two.shape), None)


def _switch(r, rmin, rmax):
    uu = ((r - rmin) / (rmax - rmin)).clamp(0.0, 1.0)
    sw = uu * uu * uu * (-6.0 * uu * uu + 15.0 * uu - 10.0) + 1.0
    return torch.where(r < rmin, torch.ones_like(sw), torch.where(r < rmax, sw, torch.zeros_like(sw)))


def evaluate(lib, model, lists):
    """(E, force[nloc,3], virial[9], extras) of one evaluation; lists = oracle.pipeline.build_lists(...)."""
    cfg = model.cfg
    nloc, nnei, M, nt = lists["nloc"], cfg.nnei, model.M, cfg.ntypes
    ext_c, ext_t = lists["coord"], lists["atype"]
    avg, std = model.davg.numpy(), model.dstd.numpy()
    em, dv, rij, nl = lib.prod_env_mat_a(ext_c, ext_t, lists["offsets"], lists["neigh"], avg, std, nloc, cfg.rcut,
                                         cfg.rcut_smth, cfg.sec, f_type=np.zeros_like(ext_t))
    nl = nl.reshape(nloc, nnei)
    # torch graph: leaves = em (its gradient is net_deriv, pushed through prod_force_a / prod_virial_a) and the
    # extended coordinates (only the switch function of the gate depends on them here)
    em_t = torch.as_tensor(em.reshape(nloc, nnei, 4)).clone().requires_grad_(True)
    xc = torch.as_tensor(ext_c.reshape(-1, 3)).clone().requires_grad_(True)
    nl_t = torch.as_tensor(nl.astype(np.int64))
    pad = nl_t < 0
    jj = nl_t.clamp_min(0)
    r = (xc[jj] - xc[:nloc].unsqueeze(1)).norm(dim=-1)
    sw = torch.where(pad, torch.zeros_like(r), _switch(r, cfg.rcut_smth, cfg.rcut))
    et = torch.as_tensor(ext_t.astype(np.int64))
    nei_t = torch.where(pad | (et[jj] < 0), torch.full_like(jj, nt), et[jj])
    cen_t = torch.where(et[:nloc] < 0, torch.full_like(et[:nloc], nt), et[:nloc])
    pair = (cen_t.view(-1, 1) * (nt + 1) + nei_t).reshape(-1)
    two = model.tt_full[pair] * sw.reshape(-1, 1)
    em_x = em_t[:, :, 0].reshape(-1, 1)
    moment = _Tab.apply(lib, model.table64.numpy() if model.dtype == torch.float64 else model.table.numpy(),
                        model.info.numpy(), em_x, em_t, two, M)
    xs = moment / nnei
    d = torch.matmul(xs.permute(0, 2, 1), xs[:, :, :cfg.axis_neuron]).reshape(nloc, -1)
    g1 = torch.zeros((nloc, model.dim_in), dtype=d.dtype)
    g1 = torch.cat([d, model.tebd[cen_t], torch.zeros((nloc, model.dim_in - d.shape[1] - cfg.tebd_dim), dtype=d.dtype)], 1)
    e_atom = model.fit(g1) + model.bias_atom_e[cen_t.clamp_max(nt - 1)]
    energy = e_atom.sum()
    energy.backward()
    nd = em_t.grad.numpy().reshape(nloc, nnei * 4)
    mapping = lists["mapping"]
    nl_own = np.where(nl >= 0, mapping[np.maximum(nl, 0)], -1).astype(np.int32)
    force = lib.prod_force_a(nd, dv, nl_own, nloc)
    virial, atom_virial = lib.prod_virial_a(nd, dv, rij, nl_own, nloc)
    # switch path: forces on the extended atoms folded onto their owners; virial = - sum_k dE/dr_k (x) r_k
    gsw = xc.grad.numpy()
    np.add.at(force, mapping, -gsw)
    virial = virial + (-(gsw[:, :, None] * ext_c.reshape(-1, 3)[:, None, :]).sum(0)).reshape(9)
    return float(energy), force, virial, dict(nlist=nl, atom_energy=e_atom.detach().numpy(), xyz=moment.detach().numpy())


# ------------------------------------------------------------------------------------------------------------------
# attn_layer > 0 (DPA-1 proper): restatement of deepmd/pt/model/descriptor/se_atten.py:977-1012 (strip-mode g2 from
# the UNCOMPRESSED geometric embedding net -- the reference cannot compress a model with attention layers) and
# :1058-1447 (NeighborGatedAttention), in plain torch on the CPU with autograd for the derivatives.  Pinned against the
# reference's NumPy and PyTorch backends by tests/golden/dpa1_attn.npz (tests/test_attn_layers.py).
def _embed_mlp(x, ws, bs):
    """deepmd/dpmodel/utils/network.py EmbeddingNet: tanh layers, skip when the width is kept or doubled."""
    for w, b in zip(ws, bs):
        y = torch.tanh(x @ w + b)
        if w.shape[1] == w.shape[0]:
            y = y + x
        elif w.shape[1] == 2 * w.shape[0]:
            y = y + torch.cat([x, x], -1)
        x = y
    return x


def gated_attention_layer(x, sw, rhat, lay, hidden, scaling, shift, dotr, normalize, ln_eps):
    """One NeighborGatedAttentionLayer on x [B, nnei, M] (se_atten.py:1280-1291, 1361-1431); smooth branch only."""
    B, n, M = x.shape
    q, k, v = (x @ lay["in_w"] + lay["in_b"]).chunk(3, dim=-1)
    if normalize:
        q = torch.nn.functional.normalize(q, dim=-1)
        k = torch.nn.functional.normalize(k, dim=-1)
        v = torch.nn.functional.normalize(v, dim=-1)
    q = q * scaling
    w = torch.matmul(q, k.transpose(-2, -1))
    ww = sw[:, :, None] * sw[:, None, :]
    w = (w + shift) * ww - shift
    w = torch.softmax(w, dim=-1)
    w = w * ww
    if dotr:
        w = w * torch.matmul(rhat, rhat.transpose(-2, -1))
    o = torch.matmul(w, v) @ lay["out_w"] + lay["out_b"]
    z = x + o
    return torch.nn.functional.layer_norm(z, (M,), lay["ln_w"], lay["ln_b"], ln_eps)


def evaluate_layers(lib, model, lists, return_g2=False):
    """(E, force[nloc,3], virial[9], extras) of one evaluation of a SeAttenModel with cfg.attn_layer > 0."""
    cfg = model.cfg
    nloc, nnei, M, nt = lists["nloc"], cfg.nnei, model.M, cfg.ntypes
    ext_c, ext_t = lists["coord"], lists["atype"]
    avg, std = model.davg.numpy(), model.dstd.numpy()
    em, dv, rij, nl = lib.prod_env_mat_a(ext_c, ext_t, lists["offsets"], lists["neigh"], avg, std, nloc, cfg.rcut,
                                         cfg.rcut_smth, cfg.sec, f_type=np.zeros_like(ext_t))
    nl = nl.reshape(nloc, nnei)
    em_t = torch.as_tensor(em.reshape(nloc, nnei, 4)).clone().requires_grad_(True)
    xc = torch.as_tensor(ext_c.reshape(-1, 3)).clone().requires_grad_(True)
    nl_t = torch.as_tensor(nl.astype(np.int64))
    pad = nl_t < 0
    jj = nl_t.clamp_min(0)
    r = (xc[jj] - xc[:nloc].unsqueeze(1)).norm(dim=-1)
    sw = torch.where(pad, torch.zeros_like(r), _switch(r, cfg.rcut_smth, cfg.rcut))
    et = torch.as_tensor(ext_t.astype(np.int64))
    nei_t = torch.where(pad | (et[jj] < 0), torch.full_like(jj, nt), et[jj])
    cen_t = torch.where(et[:nloc] < 0, torch.full_like(et[:nloc], nt), et[:nloc])
    pair = cen_t.view(-1, 1) * (nt + 1) + nei_t
    gg_s = _embed_mlp(em_t[:, :, :1], [w.double() for w in model.embed.weights], [b.double() for b in model.embed.biases])
    gg_t = model.tt_full.double()[pair] * sw.unsqueeze(-1)
    gg = gg_s * gg_t + gg_s
    rhat = torch.nn.functional.normalize(em_t[:, :, 1:4], dim=-1)
    for lay in model.attn_layers:
        lay64 = {k: v.double() for k, v in lay.items()}
        gg = gated_attention_layer(gg, sw, rhat, lay64, cfg.attn, model.attn_scaling, cfg.attnw_shift, cfg.attn_dotr,
                                   cfg.attn_normalize, cfg.ln_eps)
    moment = torch.matmul(em_t.transpose(1, 2), gg)
    xs = moment / nnei
    d = torch.matmul(xs.permute(0, 2, 1), xs[:, :, :cfg.axis_neuron]).reshape(nloc, -1)
    g1 = torch.cat([d, model.tebd[cen_t], torch.zeros((nloc, model.dim_in - d.shape[1] - cfg.tebd_dim), dtype=d.dtype)], 1)
    e_atom = model.fit(g1) + model.bias_atom_e[cen_t.clamp_max(nt - 1)]
    energy = e_atom.sum()
    energy.backward()
    nd = em_t.grad.numpy().reshape(nloc, nnei * 4)
    mapping = lists["mapping"]
    nl_own = np.where(nl >= 0, mapping[np.maximum(nl, 0)], -1).astype(np.int32)
    force = lib.prod_force_a(nd, dv, nl_own, nloc)
    virial, atom_virial = lib.prod_virial_a(nd, dv, rij, nl_own, nloc)
    gsw = xc.grad.numpy()
    np.add.at(force, mapping, -gsw)
    virial = virial + (-(gsw[:, :, None] * ext_c.reshape(-1, 3)[:, None, :]).sum(0)).reshape(9)
    extras = dict(nlist=nl, atom_energy=e_atom.detach().numpy(), xyz=moment.detach().numpy())
    if return_g2:
        extras["g2"] = gg.detach().numpy()
    return float(energy.detach()), force, virial, extras
