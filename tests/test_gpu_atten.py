"""Compressed se_atten (DPA-1 strip / smooth, attn_layer = 0; BASELINE config 5) end to end on the GPU:
DeepPotB200 over SeAttenModel against the CPU checker oracle/pipeline_atten.py (reference CPU operators composed as
deepmd/pt/model/descriptor/se_atten.py:892-1016 composes the torch ops, autograd for the gate / switch / fitting
parts), fp64 1e-10."""
import numpy as np
import pytest
import torch

import __graft_entry__ as g
from oracle import cpu as ocpu
from oracle import pipeline, pipeline_atten

pytestmark = pytest.mark.gpu


def rel(a, b):
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.abs(np.asarray(a, np.float64).reshape(-1) - b).max() / np.abs(b).max())


@pytest.mark.parametrize("use_gate", [True, False])
@pytest.mark.parametrize("ncopy,jitter", [(1, 0.0), (2, 0.01)])
def test_se_atten_e2e_matches_cpu_pipeline(ncopy, jitter, use_gate):
    """use_gate: the pair-indexed gate entry points (two_embed never materialised) / the reference-schema op fed with
    the materialised tt_full[pair] * sw tensor, slab by slab."""
    g.load_package()
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel
    from deepmd_kit_b200.model import DeepPotB200

    cfg = SeAttenConfig()
    coord, atype, box = g.water_box(ncopy, jitter)
    model = SeAttenModel(cfg, torch.float64, "cuda:0")
    model.tab_chunk = 500  # several slabs of the gated table operator even on the small box
    model.use_gate = use_gate
    dp = DeepPotB200(model, skin=2.0)
    e, f, v, ae, av = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype, atomic=True)
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    lists = pipeline.build_lists(lib, coord, atype, box, cfg.rcut + 2.0)
    we, wf, wv, ex = pipeline_atten.evaluate(lib, SeAttenModel(cfg, torch.float64, "cpu"), lists)
    assert (ex["nlist"] >= 0).sum(1).mean() > 80
    assert abs(e[0, 0] - we) <= 1e-10 * abs(we)
    assert rel(ae[0].reshape(-1), ex["atom_energy"]) <= 1e-10
    assert rel(f[0], wf) <= 1e-10
    assert rel(v[0], wv) <= 1e-10
    assert rel(av[0].sum(0), wv) <= 1e-9
    # graph replay (second call) gives the same answer
    e2, f2, v2 = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    assert rel(f2[0], f[0]) <= 1e-12


def test_se_atten_e2e_fp32_against_fp64_checker():
    """fp32 model (fitting net on the four-slice int8 tensor-core kernels).  The fp32 CPU pipeline is NOT a usable
    yardstick here: it differs from its own fp64 evaluation by 1.8e-4 on the forces (single-precision cancellation in
    the autograd switch path and in prod_force), ten times more than the GPU path does.  So the fp32 GPU result is held
    against the fp64 checker at single-precision level: measured 1.8e-5 (force), 1e-6 (virial), < 1e-7 (energy)."""
    g.load_package()
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel
    from deepmd_kit_b200.model import DeepPotB200

    cfg = SeAttenConfig()
    coord, atype, box = g.water_box(2, 0.01)
    model = SeAttenModel(cfg, torch.float32, "cuda:0")
    assert model.use_tc and model.nslice == 4
    dp = DeepPotB200(model, skin=2.0)
    e, f, v = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    lists = pipeline.build_lists(lib, coord, atype, box, cfg.rcut + 2.0)
    we, wf, wv, ex = pipeline_atten.evaluate(lib, SeAttenModel(cfg, torch.float64, "cpu"), lists)
    assert abs(e[0, 0] - we) <= 1e-6 * abs(we)
    assert rel(f[0], wf) <= 5e-5
    assert rel(v[0], wv) <= 1e-5


def test_se_atten_force_is_energy_gradient():
    g.load_package()
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel
    from deepmd_kit_b200.model import DeepPotB200

    coord, atype, box = g.water_box(1, 0.02)
    dp = DeepPotB200(SeAttenModel(SeAttenConfig(), torch.float64, "cuda:0"), skin=2.0, use_graph=False)
    e0, f0, _ = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    h = 1e-5
    for i, d in ((3, 0), (77, 2)):
        es = []
        for sgn in (1, -1):
            c = coord.copy()
            c[i, d] += sgn * h
            es.append(dp.eval(c.reshape(1, -1), box.reshape(1, 9), atype)[0][0, 0])
        assert abs(-(es[0] - es[1]) / (2 * h) - f0[0, i, d]) < 1e-7


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gate_op_equals_reference_schema_op(dtype):
    """dpb200_tabulate_fusion_se_atten_gate[_grad] against tabulate_fusion_se_atten[_grad] fed with the materialised
    two_embed = tt_full[pair] * sw: same forward, same em gradients, dy_dsw = sum_k dy_dtwo * tt_full[pair]."""
    pkg = g.load_package()
    ops = pkg.ops
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

    torch.manual_seed(3)
    m = SeAttenModel(SeAttenConfig(), dtype, "cuda:0")
    nloc, nnei, M = 300, 120, 100
    em = torch.randn(nloc, nnei, 4, dtype=dtype, device="cuda:0") * 0.3
    em[:, :, 0] = torch.sort(torch.rand(nloc, nnei, dtype=dtype, device="cuda:0") * 6 - 0.5, dim=1, descending=True)[0]
    em[:, 100:, :] = torch.tensor([-0.36, 0.0, 0.0, 0.0], dtype=dtype, device="cuda:0")  # padded tail
    em_x = em[:, :, 0].reshape(-1, 1).contiguous()
    pair = torch.randint(0, 9, (nloc, nnei), dtype=torch.int32, device="cuda:0")
    sw = torch.rand(nloc, nnei, dtype=dtype, device="cuda:0")
    sw[:, 100:] = 0
    two = (m.tt_full[pair.reshape(-1).long()] * sw.reshape(-1, 1)).contiguous()
    dy = torch.randn(nloc, 4, M, dtype=dtype, device="cuda:0")
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    a = ops.tabulate_fusion_se_atten_gate(m.table, m.info, em_x, em, m.tt_full, pair, sw, M)
    b = ops.tabulate_fusion_se_a(m.table, m.info, em_x, em, M, two_embed=two, is_sorted=True)
    assert ((a - b).abs().max() / b.abs().max()).item() <= tol
    gx, gem, gq = ops.tabulate_fusion_se_atten_gate_grad(m.table, m.info, em_x, em, m.tt_full, pair, sw, dy, M)
    wx, wem, wtwo = ops.tabulate_fusion_se_a_grad(m.table, m.info, em_x, em, dy, M, two_embed=two, is_sorted=True)
    wq = (wtwo * m.tt_full[pair.reshape(-1).long()]).sum(1).reshape(nloc, nnei)
    assert ((gx - wx).abs().max() / wx.abs().max()).item() <= tol
    assert ((gem - wem).abs().max() / wem.abs().max()).item() <= tol
    assert ((gq - wq)[:, :100].abs().max() / wq.abs().max()).item() <= 10 * tol
    # em_x gradient folded into component 0 (dy_dem_x = NULL), with and without the compressed coefficient table
    want = wem.clone()
    want[:, :, 0] += wx.reshape(nloc, nnei)
    flag_sets = [0]
    if dtype == torch.float64:
        fl = ops.compressed_coef_flags(m.table64, m.info)
        assert fl == m.coef_flags and fl != 0
        flag_sets.append(fl)
    for fl in flag_sets:
        none_x, gem2, gq2 = ops.tabulate_fusion_se_atten_gate_grad(m.table, m.info, em_x, em, m.tt_full, pair, sw, dy, M,
                                                                   fuse_x=True, flags=fl)
        t2 = tol if fl == 0 else 1e-11
        assert none_x is None
        a2 = ops.tabulate_fusion_se_atten_gate(m.table, m.info, em_x, em, m.tt_full, pair, sw, M, flags=fl)
        assert ((a2 - b).abs().max() / b.abs().max()).item() <= t2
        assert ((gem2 - want).abs().max() / want.abs().max()).item() <= t2
        assert ((gq2 - wq)[:, :100].abs().max() / wq.abs().max()).item() <= 10 * t2
        assert float(gq2[:, 101:].abs().max()) == 0.0  # behind the folded padding entry


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gate_scalars_and_pair_force_match_torch(dtype):
    """dpb200_se_atten_gate_scalars against the torch expressions of SeAttenModel.gate_scalars (se_atten.py:916-926,
    switcher.h:61-84), and dpb200_prod_force_virial_a_pair against prod_force_virial_a + the explicit scatter of
    -(q w) r_ij (force on the neighbour, opposite on the centre, virial and atomic virial)."""
    pkg = g.load_package()
    ops = pkg.ops
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel, switch_and_derivative

    dev = "cuda:0"
    torch.manual_seed(11)
    cfg = SeAttenConfig()
    nloc, nall, nnei, nt = 257, 400, cfg.nnei, cfg.ntypes
    nlist = torch.randint(0, nall, (nloc, nnei), dtype=torch.int32, device=dev)
    nlist[:, 90:] = -1
    nlist[5] = -1
    ext_type = torch.randint(0, nt, (nall,), dtype=torch.int32, device=dev)
    ext_type[7] = -1  # virtual atom: padding type
    rij = torch.randn(nloc, nnei, 3, dtype=dtype, device=dev) * 2.5
    rij[3, 0] = 0.0   # r = 0: inside rcut_smth, no switch force
    rij[4, 1] = torch.tensor([cfg.rcut + 1.0, 0.0, 0.0], dtype=dtype, device=dev)  # beyond rcut
    pair, sw, dswr = ops.se_atten_gate_scalars(nlist, ext_type, rij.reshape(nloc, -1), nloc, nnei, nt, cfg.rcut_smth, cfg.rcut)
    m = SeAttenModel.__new__(SeAttenModel)
    m.cfg = cfg
    wpair, wsw, wdsw, wr = SeAttenModel.gate_scalars(m, ext_type, nlist, rij.reshape(nloc, -1), 0, nloc)
    assert torch.equal(pair.long(), wpair)
    tol = 1e-13 if dtype == torch.float64 else 2e-6
    assert (sw - wsw).abs().max().item() <= tol
    wdswr = torch.where((wr > 0) & (nlist >= 0), wdsw / wr.clamp_min(1e-30), torch.zeros_like(wr))
    assert (dswr - wdswr).abs().max().item() <= tol * max(1.0, wdswr.abs().max().item())
    # pair force
    nd = torch.randn(nloc, nnei * 4, dtype=dtype, device=dev)
    dv = torch.randn(nloc, nnei * 12, dtype=dtype, device=dev)
    q = torch.randn(nloc, nnei, dtype=dtype, device=dev)
    f0, v0, a0 = ops.prod_force_virial_a(nd, dv, rij.reshape(nloc, -1), nlist, nloc, nall, nnei, atom_virial=True)
    f1, v1, a1 = ops.prod_force_virial_a_pair(nd, dv, rij.reshape(nloc, -1), nlist, q, dswr, nloc, nall, nnei,
                                              atom_virial=True)
    vec = (q * dswr).unsqueeze(-1) * rij  # dE/dr_j of the switch path
    vec = torch.where((nlist >= 0).unsqueeze(-1), vec, torch.zeros_like(vec))
    idx = nlist.clamp_min(0).long().reshape(-1)
    want_f = f0.reshape(-1, 3).clone()
    want_f.index_add_(0, idx, -vec.reshape(-1, 3))
    want_f[:nloc] += vec.sum(1)
    want_v = v0 - torch.einsum("pi,pj->ij", vec.reshape(-1, 3), rij.reshape(-1, 3)).reshape(9)
    want_a = a0.reshape(-1, 9).clone()
    want_a.index_add_(0, idx, -(vec.reshape(-1, 3, 1) * rij.reshape(-1, 1, 3)).reshape(-1, 9))
    ftol = 1e-12 if dtype == torch.float64 else 3e-5
    assert ((f1.reshape(-1, 3) - want_f).abs().max() / want_f.abs().max()).item() <= ftol
    assert ((v1 - want_v).abs().max() / want_v.abs().max()).item() <= ftol
    assert ((a1.reshape(-1, 9) - want_a).abs().max() / want_a.abs().max()).item() <= ftol


def test_gate_desc_epilogue_and_slice_cols():
    """The gated forward with the descriptor epilogue: same table output as the plain gated forward; the int8 operand
    rows reproduce [D | tebd(centre) | 0] to 2^-44 of 2^row_exp with row_exp >= the requested lower bound."""
    pkg = g.load_package()
    ops = pkg.ops
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

    dev = "cuda:0"
    torch.manual_seed(5)
    dtype = torch.float64
    m = SeAttenModel(SeAttenConfig(), dtype, dev)
    cfg = m.cfg
    nloc, nnei, M = 130, cfg.nnei, m.M
    em = torch.randn(nloc, nnei, 4, dtype=dtype, device=dev) * 0.3
    em[:, :, 0] = torch.sort(torch.rand(nloc, nnei, dtype=dtype, device=dev) * 5 - 0.3, dim=1, descending=True)[0]
    em[:, 100:, :] = torch.tensor([-0.36, 0.0, 0.0, 0.0], dtype=dtype, device=dev)
    em[9] *= 1e-6  # a tiny row: its exponent is set by the lower bound
    em_x = em[:, :, 0].reshape(-1, 1).contiguous()
    pair = torch.randint(0, 9, (nloc, nnei), dtype=torch.int32, device=dev)
    sw = torch.rand(nloc, nnei, dtype=dtype, device=dev)
    sw[:, 100:] = 0
    want = ops.tabulate_fusion_se_atten_gate(m.table, m.info, em_x, em, m.tt_full, pair, sw, M)
    inv = 1.0 / nnei
    out, desc, ex = ops.tabulate_fusion_se_atten_gate_desc(m.table, m.info, em_x, em, m.tt_full, pair, sw, M,
                                                           cfg.axis_neuron, inv, m.dim_in, 6, m.tebd_exp, pad_rows=32)
    assert torch.equal(out, want)
    out_c, desc_c, ex_c = ops.tabulate_fusion_se_atten_gate_desc(m.table, m.info, em_x, em, m.tt_full, pair, sw, M,
                                                                 cfg.axis_neuron, inv, m.dim_in, 6, m.tebd_exp,
                                                                 pad_rows=32, flags=m.coef_flags)
    assert m.coef_flags != 0 and ((out_c - want).abs().max() / want.abs().max()).item() <= 1e-11  # (flag ignored)
    assert (ex_c - ex).abs().max().item() <= 1
    assert desc.shape == (nloc + 32, 6 * m.dim_in) and int(desc[nloc:].abs().sum()) == 0
    assert int(ex[:nloc].min()) >= m.tebd_exp
    ctype = torch.randint(0, cfg.ntypes + 1, (nloc,), dtype=torch.int32, device=dev)
    ops.fit_slice_cols(desc, m.dim_in, m.dim_d, 6, ex, m.tebd, idx=ctype)
    sl = desc[:nloc].reshape(nloc, 6, m.dim_in).to(torch.float64)
    assert int(sl.abs().max()) <= 128
    w = torch.tensor([2.0 ** (-7 - 8 * s) for s in range(6)], dtype=torch.float64, device=dev)
    scale = torch.ldexp(torch.ones(nloc, dtype=torch.float64, device=dev), ex[:nloc])
    rec = (sl * w[None, :, None]).sum(1) * scale[:, None]
    d = ops.se_a_descriptor(want, cfg.axis_neuron, inv)
    full = torch.zeros(nloc, m.dim_in, dtype=torch.float64, device=dev)
    full[:, :m.dim_d] = d
    full[:, m.dim_d:m.dim_d + cfg.tebd_dim] = m.tebd[ctype.long()]
    assert ((rec - full).abs() / scale[:, None]).max().item() <= 2.0 ** -44
