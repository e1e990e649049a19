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
        assert ((gem2 - want).abs().max() / want.abs().max()).item() <= t2
        assert ((gq2 - wq)[:, :100].abs().max() / wq.abs().max()).item() <= 10 * t2
        assert float(gq2[:, 101:].abs().max()) == 0.0  # behind the folded padding entry
