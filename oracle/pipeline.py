"""TEST INFRASTRUCTURE ONLY — the reference's decomposed CPU pipeline for one compressed se_e2_a
energy / force / virial evaluation, driven through oracle.cpu.CpuLib ("reference" = the unmodified
library compiled from /root/reference, "port" = our C restatement).

Stages (SURVEY.md 8d "CPU baseline"): raw neighbour list -> prod_env_mat_a_cpu ->
tabulate_fusion_se_a_cpu per table -> descriptor algebra + fitting net (torch, CPU) ->
tabulate_fusion_se_a_grad_cpu per table -> prod_force_a_cpu -> prod_virial_a_cpu.

Used by tests/ (end-to-end parity), __graft_entry__.smoke() and bench.py (`cpu_baseline`,
`--impl reference`).  The product never imports this module.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import cpu as ocpu


def build_lists(lib, coord, atype, box, rc_list):
    """normalize -> periodic ghost copy -> raw list with the strict cutoff rc_list."""
    dtype = coord.dtype
    c = lib.normalize_coord(np.asarray(coord, dtype=dtype).reshape(-1, 3), box)
    ext_c, ext_t, mapping = lib.copy_coord(c, atype, box, rc_list)
    nloc = len(atype)
    numneigh, rows = lib.build_nlist(ext_c, nloc, rc_list, atype=ext_t, mem_size=None) if lib.kind == "port" \
        else _reference_lists(lib, ext_c, ext_t, nloc, rc_list)
    off, neigh = ocpu.dense_to_csr(rows, numneigh)
    return dict(coord=ext_c, atype=ext_t, mapping=mapping, nloc=nloc, offsets=off, neigh=neigh, numneigh=numneigh)


def _reference_lists(lib, ext_c, ext_t, nloc, rc_list):
    # build_nlist_cpu of the reference is O(nloc*nall): fine for the bounded baseline samples
    return lib.build_nlist(ext_c, nloc, rc_list, atype=ext_t)


def reference_energy_and_dy(model, xyz, perm, ranges):
    """Descriptor algebra + fitting net exactly as the reference's PyTorch backend writes them
    (deepmd/pt/model/descriptor/se_a.py:843-850: xyz_scatter /= nnei; matmul(xyz_scatter_1,
    xyz_scatter_2); per-type fitting MLP with resnet_dt, deepmd/pt/model/network/mlp.py), with
    dE/d(xyz_scatter) from torch.autograd.  Plain torch on the CPU: the checker for the dpb200
    descriptor kernels and the hand-written MLP backward."""
    cfg = model.cfg
    x = xyz.detach().clone().requires_grad_(True)
    xs = x / cfg.nnei
    d = torch.matmul(xs.permute(0, 2, 1), xs[:, :, :cfg.axis_neuron]).reshape(x.shape[0], -1)
    e_atom = torch.zeros(x.shape[0], dtype=x.dtype)
    for t, (a, b) in enumerate(ranges):
        idx = perm[a:b]
        if idx.numel():
            e_atom = e_atom.index_add(0, idx, model.fit[t](d.index_select(0, idx)) + model.bias_atom_e[t].to(x.dtype))
    energy = e_atom.sum()
    (dy,) = torch.autograd.grad(energy, x)
    return energy.detach(), e_atom.detach(), dy


def evaluate(lib, model, lists, timings=None):
    """One force evaluation. `model` is a deepmd_kit_b200.model.SeAModel living on the CPU (only its
    weights, tables and torch fitting net are used).  Returns (E, force[nloc,3], virial[9], extras)."""
    cfg = model.cfg
    np_dt = np.float64 if model.dtype == torch.float64 else np.float32
    sec = cfg.sec
    nloc = lists["nloc"]
    nnei = cfg.nnei
    M = model.M

    def tick(name, t0):
        if timings is not None:
            timings[name] = timings.get(name, 0.0) + (time.perf_counter() - t0)

    t0 = time.perf_counter()
    avg = model.davg.cpu().numpy().astype(np_dt)
    std = model.dstd.cpu().numpy().astype(np_dt)
    em, dv, rij, nl = lib.prod_env_mat_a(lists["coord"].astype(np_dt), lists["atype"], lists["offsets"], lists["neigh"],
                                         avg, std, nloc, cfg.rcut, cfg.rcut_smth, sec)
    tick("prod_env_mat_a", t0)
    t0 = time.perf_counter()
    em3 = em.reshape(nloc, nnei, 4)
    tables = [t.cpu().numpy().astype(np_dt) for t in model.tables]
    infos = [i.numpy().astype(np_dt) for i in model.infos]
    xyz = np.zeros((nloc, 4, M), np_dt)
    for t in range(cfg.ntypes):
        sl = em3[:, sec[t]:sec[t + 1], :]
        if sl.shape[1] == 0:
            continue
        xyz += lib.tabulate_fusion_se_a(tables[t], infos[t], np.ascontiguousarray(sl[:, :, 0]).reshape(-1, 1),
                                        np.ascontiguousarray(sl), M)
    tick("tabulate_fusion_se_a", t0)
    t0 = time.perf_counter()
    at = torch.as_tensor(lists["atype"][:nloc]).to(torch.int64)
    perm = torch.argsort(at, stable=True)
    cnt = torch.bincount(at, minlength=cfg.ntypes).tolist()
    ranges, a0 = [], 0
    for c in cnt[:cfg.ntypes]:
        ranges.append((a0, a0 + int(c)))
        a0 += int(c)
    energy, e_atom, dy = reference_energy_and_dy(model, torch.as_tensor(xyz), perm, ranges)
    dy = dy.numpy()
    tick("fitting_net", t0)
    t0 = time.perf_counter()
    nd = np.zeros((nloc, nnei, 4), np_dt)
    for t in range(cfg.ntypes):
        sl = em3[:, sec[t]:sec[t + 1], :]
        if sl.shape[1] == 0:
            continue
        gx, gem, _ = lib.tabulate_fusion_se_a_grad(tables[t], infos[t], np.ascontiguousarray(sl[:, :, 0]).reshape(-1, 1),
                                                   np.ascontiguousarray(sl), dy, M)
        gem = gem.reshape(nloc, -1, 4)
        gem[:, :, 0] += gx.reshape(nloc, -1)
        nd[:, sec[t]:sec[t + 1], :] = gem
    nd = nd.reshape(nloc, nnei * 4)
    tick("tabulate_fusion_se_a_grad", t0)
    t0 = time.perf_counter()
    mapping = lists["mapping"]
    nl_own = np.where(nl >= 0, mapping[np.maximum(nl, 0)], -1).astype(np.int32)
    force = lib.prod_force_a(nd, dv, nl_own, nloc)
    tick("prod_force_a", t0)
    t0 = time.perf_counter()
    virial, atom_virial = lib.prod_virial_a(nd, dv, rij, nl_own, nloc)
    tick("prod_virial_a", t0)
    return float(energy), force, virial, dict(nlist=nl, em=em, em_deriv=dv, rij=rij, net_deriv=nd, xyz=xyz,
                                              atom_energy=e_atom.numpy(), atom_virial=atom_virial)
