"""DPA-1 with attention layers (deepmd/pt/model/descriptor/se_atten.py:1058-1447).

Golden: tests/golden/dpa1_attn.npz, written by tests/golden/make_dpa1_attn.py from the reference's NumPy backend
(descriptor, attention output g2, atomic energies) and PyTorch backend (autograd energy / force / virial) for
se_atten_v2 with attn_layer 2, attn 128, attn_dotr, normalised q / k / v on the 192-atom water frame.

CPU (`not gpu`): the checker (oracle.pipeline_atten.evaluate_layers: the UNCOMPRESSED embedding net + a plain torch
restatement of the attention layers, autograd) against the golden.
GPU: the product path (table-compressed embedding + csrc/attn_layers.cu stages + library GEMMs, hand-derived backward)
against the golden at 1e-10, and every stage kernel against torch autograd of the same formula."""
import json
import os

import numpy as np
import pytest
import torch

import __graft_entry__ as g
from oracle import cpu as ocpu
from oracle import pipeline
from oracle import pipeline_atten

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.abs(np.asarray(a, np.float64).reshape(-1) - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(ROOT, "tests", "golden", "dpa1_attn.npz"))
    return {k: z[k] for k in z.files}


def _model(gold, device, dtype=torch.float64):
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

    c = json.loads(str(gold["config_json"]))
    cfg = SeAttenConfig(ntypes=c["ntypes"], nsel=c["sel"], rcut=c["rcut"], rcut_smth=c["rcut_smth"],
                        neuron=tuple(c["neuron"]), axis_neuron=c["axis_neuron"], tebd_dim=c["tebd_dim"],
                        stats=tuple(tuple(s) for s in c["stats"]), fitting_neuron=tuple(c["fitting_neuron"]),
                        fitting_resnet_dt=bool(c["fitting_resnet_dt"]), attn_layer=c["attn_layer"], attn=c["attn"],
                        attn_dotr=c["attn_dotr"], attn_normalize=c["normalize"], ln_eps=c["ln_eps"],
                        attnw_shift=c["attnw_shift"])
    assert abs((cfg.attn * cfg.scaling_factor) ** -0.5 - c["scaling"]) < 1e-15

    def net(prefix, n):
        return [dict(w=gold[f"w_{prefix}_{i}_w"], b=gold[f"w_{prefix}_{i}_b"]) for i in range(n)]

    # (a layer without timestep was stored as a NaN scalar: np.asarray(None, float64))
    fit_layers = [dict(w=gold[f"w_fit_{i}_w"], b=gold[f"w_fit_{i}_b"],
                       idt=None if gold[f"w_fit_{i}_idt"].ndim == 0 else gold[f"w_fit_{i}_idt"])
                  for i in range(c["n_fit_layers"])]
    weights = dict(embed=net("embed", 3), strip=net("strip", 3), tebd=gold["w_tebd"],
                   fit=dict(layers=fit_layers, head=dict(w=gold["w_fit_head_w"], b=gold["w_fit_head_b"])),
                   bias_atom_e=gold["w_bias_atom_e"],
                   attn=[dict(in_w=gold[f"w_attn_{i}_in_w"], in_b=gold[f"w_attn_{i}_in_b"], out_w=gold[f"w_attn_{i}_out_w"],
                              out_b=gold[f"w_attn_{i}_out_b"], ln_w=gold[f"w_attn_{i}_ln_w"], ln_b=gold[f"w_attn_{i}_ln_b"])
                         for i in range(c["attn_layer"])])
    return SeAttenModel(cfg, dtype, device, weights=weights)


def test_checker_matches_reference_backends(gold):
    """oracle restatement (uncompressed embedding + attention layers, torch autograd) == reference NumPy / PyTorch."""
    g.load_package()
    model = _model(gold, "cpu")
    coord, atype, box = g.water_box(1, 0.0)
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    lists = pipeline.build_lists(lib, coord, atype, box, model.cfg.rcut + 2.0)
    e, f, v, ex = pipeline_atten.evaluate_layers(lib, model, lists, return_g2=True)
    assert ((ex["nlist"] >= 0).sum(1) == gold["x_numneigh"]).all()
    assert rel(ex["g2"][[0, 100]], gold["x_gg_rows"]) <= 1e-12
    xs = torch.as_tensor(ex["xyz"]) / model.cfg.nnei
    d = torch.matmul(xs.permute(0, 2, 1), xs[:, :, :model.cfg.axis_neuron]).reshape(len(atype), -1)
    got = torch.cat([d, model.tebd[torch.as_tensor(atype.astype(np.int64))]], 1).numpy()
    assert rel(got[gold["x_rows"]], gold["x_descriptor"]) <= 1e-12
    assert abs(got.sum() - float(gold["x_total"])) <= 1e-12 * abs(float(gold["x_total"]))
    assert abs((got * got).sum() - float(gold["x_total_sq"])) <= 1e-12 * float(gold["x_total_sq"])
    assert rel(ex["atom_energy"], gold["x_atomic_energy"]) <= 1e-12
    assert abs(e - float(gold["x_pt_energy"])) <= 1e-12 * abs(float(gold["x_pt_energy"]))
    assert rel(f, gold["x_pt_force"]) <= 1e-10
    assert rel(v, gold["x_pt_virial"]) <= 1e-10


@pytest.mark.gpu
def test_product_matches_reference_backends_gpu(gold):
    """DeepPotB200 on the model with attention layers: E / F / V of the reference's PyTorch backend, 1e-10."""
    g.load_package()
    from deepmd_kit_b200._lib import lib
    from deepmd_kit_b200.model import DeepPotB200

    n0 = lib().launch_count()
    model = _model(gold, "cuda")
    coord, atype, box = g.water_box(1, 0.0)
    dp = DeepPotB200(model, skin=2.0)
    e, f, v, ex = dp.eval_device(torch.as_tensor(coord).cuda(), torch.as_tensor(atype).cuda(), box)
    assert rel(ex["atom_energy"].cpu().numpy(), gold["x_atomic_energy"]) <= 1e-10
    assert abs(float(e) - float(gold["x_pt_energy"])) <= 1e-10 * abs(float(gold["x_pt_energy"]))
    assert rel(f.cpu().numpy(), gold["x_pt_force"]) <= 1e-10
    assert rel(v.cpu().numpy(), gold["x_pt_virial"]) <= 1e-10
    # two slabs of centre atoms give the same answer as one
    model.attn_chunk = 100
    e2, f2, v2, _ = dp.eval_device(torch.as_tensor(coord).cuda(), torch.as_tensor(atype).cuda(), box)
    assert abs(float(e2) - float(e)) <= 1e-12 * abs(float(e))
    assert rel(f2.cpu().numpy(), f.cpu().numpy()) <= 1e-11
    # ... and so does the evaluation on all nnei slots (default: the trailing empty slots of a slab are folded into one)
    assert model.attn_compact
    model.attn_compact = False
    e3, f3, v3, _ = dp.eval_device(torch.as_tensor(coord).cuda(), torch.as_tensor(atype).cuda(), box)
    assert abs(float(e3) - float(e)) <= 1e-12 * abs(float(e))
    assert rel(f3.cpu().numpy(), f.cpu().numpy()) <= 1e-11
    assert rel(v3.cpu().numpy(), v.cpu().numpy()) <= 1e-11
    assert rel(f3.cpu().numpy(), gold["x_pt_force"]) <= 1e-10
    assert lib().launch_count() > n0


@pytest.mark.gpu
def test_product_fp32_model_gpu(gold):
    """The float32 model with attention layers end to end (fp32 table, stage kernels, library SGEMMs, four-slice
    fitting net) against the reference's fp64 PyTorch-backend values: single-precision agreement."""
    g.load_package()
    from deepmd_kit_b200.model import DeepPotB200

    model = _model(gold, "cuda", torch.float32)
    coord, atype, box = g.water_box(1, 0.0)
    dp = DeepPotB200(model, skin=2.0)
    e, f, v, ex = dp.eval_device(torch.as_tensor(coord.astype(np.float32)).cuda(), torch.as_tensor(atype).cuda(), box)
    assert f.dtype == torch.float32
    ee = abs(float(e) - float(gold["x_pt_energy"])) / abs(float(gold["x_pt_energy"]))
    fe = rel(f.double().cpu().numpy(), gold["x_pt_force"])
    ve = rel(v.double().cpu().numpy(), gold["x_pt_virial"])
    print(f"fp32 attention model vs fp64 reference: energy {ee:.2e} force {fe:.2e} virial {ve:.2e}")
    assert ee <= 1e-6 and fe <= 1e-5 and ve <= 1e-5  # measured on B200: 5.9e-08, 2.5e-06, 4.0e-06


@pytest.mark.gpu
def test_attention_g2_matches_reference_gpu(gold):
    """The attention output g2 of two atoms (forward only) against the reference's NumPy backend."""
    g.load_package()
    from deepmd_kit_b200 import ops
    from deepmd_kit_b200.model import DeepPotB200

    model = _model(gold, "cuda")
    coord, atype, box = g.water_box(1, 0.0)
    dp = DeepPotB200(model, skin=2.0)
    _, _, _, ex = dp.eval_device(torch.as_tensor(coord).cuda(), torch.as_tensor(atype).cuda(), box)
    st = dp.state
    cfg = model.cfg
    nloc = len(atype)
    em, dv, rij, nlist = ops.prod_env_mat_a(dp._last_ext_coord.reshape(-1), st.ext_type, st.numneigh, st.rows, model.davg,
                                            model.dstd, nloc, st.ext_type.numel(), cfg.rcut, cfg.rcut_smth, cfg.sec,
                                            f_type=torch.zeros_like(st.ext_type))
    em3 = em.reshape(nloc, cfg.nnei, 4)
    pair, sw, _ = ops.se_atten_gate_scalars(nlist, st.ext_type, rij, nloc, cfg.nnei, cfg.ntypes, cfg.rcut_smth, cfg.rcut)
    x0, _, _ = ops.se_atten_embed(model.table, model.info, em3, model.tt_full, pair, sw, model.M)
    rhat, _ = ops.se_atten_rhat(em3)
    x, _ = model.attention_forward(x0, sw, rhat, keep=False)
    assert rel(x[[0, 100]].cpu().numpy(), gold["x_gg_rows"]) <= 1e-10


def _rand_layer(M, h, dtype, dev, seed):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    lay = dict(in_w=r(M, 3 * h) / (M + 3 * h) ** 0.5, in_b=r(3 * h), out_w=r(h, M) / (M + h) ** 0.5, out_b=r(M),
               ln_w=1 + 0.2 * r(M), ln_b=0.1 * r(M))
    return {k: v.to(dev, dtype) for k, v in lay.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,n,M,h,dotr,normalize", [
    (torch.float64, 120, 100, 128, True, True),
    (torch.float64, 37, 24, 20, True, True),     # ragged widths: nnei, M, hidden not multiples of 32
    (torch.float64, 64, 32, 32, False, True),
    (torch.float64, 33, 16, 48, True, False),
    (torch.float32, 120, 100, 128, True, True),
])
def test_attention_stack_against_torch_autograd(dtype, n, M, h, dotr, normalize):
    """attention_forward / attention_backward (stage kernels + library GEMMs) on random slabs with empty slots
    (sw = 0, rhat = 0) against torch autograd of the plain restatement, including dE/d(sw) and dE/d(rhat)."""
    g.load_package()
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

    cfg = SeAttenConfig(nsel=n, neuron=(8, M), attn_layer=2, attn=h, attn_dotr=dotr,
                        attn_normalize=normalize, fitting_neuron=(16,))
    model = SeAttenModel(cfg, dtype, "cuda")
    model.attn_layers = []
    for li in range(2):
        lay = _rand_layer(M, h, dtype, "cuda", 11 + li)
        lay["in_wt"], lay["out_wt"] = lay["in_w"].t().contiguous(), lay["out_w"].t().contiguous()
        model.attn_layers.append(lay)
    gen = torch.Generator().manual_seed(5)
    B = 7
    x = torch.randn(B, n, M, generator=gen, dtype=torch.float64)
    sw = torch.rand(B, n, generator=gen, dtype=torch.float64)
    rh = torch.nn.functional.normalize(torch.randn(B, n, 3, generator=gen, dtype=torch.float64), dim=-1)
    npad = torch.randint(0, n // 2, (B,), generator=gen)
    for bi in range(B):
        if npad[bi] > 0:
            sw[bi, -int(npad[bi]):] = 0
            rh[bi, -int(npad[bi]):] = 0
    dout = torch.randn(B, n, M, generator=gen, dtype=torch.float64)
    xd, swd, rhd, doutd = (t.to("cuda", dtype).contiguous() for t in (x, sw, rh, dout))
    y, saved = model.attention_forward(xd.clone(), swd, rhd)
    d_sw = torch.zeros_like(swd)
    d_rh = torch.zeros_like(rhd)
    dx = model.attention_backward(doutd.clone(), saved, swd, rhd, d_sw, d_rh)
    # the checker: fp64 torch autograd
    xr, swr, rhr = (t.cuda().requires_grad_(True) for t in (x, sw, rh))
    yr = xr
    for lay in model.attn_layers:
        yr = pipeline_atten.gated_attention_layer(yr, swr, rhr, {k: v.double() for k, v in lay.items()}, h,
                                                  model.attn_scaling, cfg.attnw_shift, dotr, normalize, cfg.ln_eps)
    (yr * dout.cuda()).sum().backward()
    tol = 1e-11 if dtype == torch.float64 else 2e-4
    for got, want in ((y, yr), (dx, xr.grad), (d_sw, swr.grad)) + (((d_rh, rhr.grad),) if dotr else ()):
        want = want.detach()
        err = float((got.double() - want).abs().max() / want.abs().max())
        assert err <= tol, err


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_embed_and_rhat_stages(dtype):
    """se_atten_embed (+ grad) against the quintic of the table evaluated in torch, se_atten_rhat (+ grad) against
    torch.nn.functional.normalize, including zero rows (empty slots) and out-of-range s (both extrapolation sides)."""
    g.load_package()
    from deepmd_kit_b200 import ops
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

    model = SeAttenModel(SeAttenConfig(nsel=40), torch.float64, "cuda")
    cfg, M = model.cfg, model.M
    gen = torch.Generator().manual_seed(3)
    B, n = 9, cfg.nnei
    info = model.info.numpy()
    s = torch.empty(B, n, dtype=torch.float64).uniform_(float(info[0]) - 0.5, float(info[2]) + 0.5, generator=gen)
    em = torch.randn(B, n, 4, generator=gen, dtype=torch.float64)
    em[:, :, 0] = s
    em[:, -5:, 1:] = 0
    sw = torch.rand(B, n, generator=gen, dtype=torch.float64)
    pair = torch.randint(0, (cfg.ntypes + 1) ** 2, (B, n), generator=gen).to(torch.int32)
    emd, swd, paird = em.to("cuda", dtype), sw.to("cuda", dtype), pair.cuda()
    table = model.table64.to("cuda", dtype)
    tt = model.tt_full.to(dtype)
    x0, gs, dgs = ops.se_atten_embed(table, model.info, emd, tt, paird, swd, M)
    # checker: locate + quintic in torch fp64 (tabulate.cc:45-73)
    lower, upper, vmax, s0, s1 = (float(v) for v in info[:5])
    first = int((upper - lower) / s0)
    tab = model.table64.reshape(-1, M, 6)
    sf = s.reshape(-1)
    idx = torch.where(sf < lower, torch.zeros_like(sf),
                      torch.where(sf < upper, torch.floor((sf - lower) / s0),
                                  torch.where(sf < vmax, first + torch.floor((sf - upper) / s1),
                                              torch.full_like(sf, tab.shape[0] - 1))))
    base = torch.where(idx < first, idx * s0 + lower, (idx - first) * s1 + upper)  # (fp64: idx is a float64 tensor)
    idx = idx.to(torch.int64)
    xx = torch.where(sf < lower, torch.zeros_like(sf), torch.where(sf < vmax, sf - base, vmax - base))
    dl = torch.where(sf < lower, sf - lower, torch.where(sf < vmax, torch.zeros_like(sf), sf - vmax))
    a = tab[idx]
    xx_, dl_ = xx[:, None], dl[:, None]
    g1 = a[:, :, 1] + (2 * a[:, :, 2] + (3 * a[:, :, 3] + (4 * a[:, :, 4] + 5 * a[:, :, 5] * xx_) * xx_) * xx_) * xx_
    g0 = a[:, :, 0] + (a[:, :, 1] + (a[:, :, 2] + (a[:, :, 3] + (a[:, :, 4] + a[:, :, 5] * xx_) * xx_) * xx_) * xx_) * xx_ + g1 * dl_
    ttp = model.tt_full.double().cpu()[pair.reshape(-1).to(torch.int64)]
    want_x0 = g0 * (1 + ttp * sw.reshape(-1, 1))
    tol = 1e-12 if dtype == torch.float64 else 3e-5
    for got, want in ((gs, g0), (dgs, g1), (x0, want_x0)):
        err = float((got.double().cpu().reshape(-1, M) - want).abs().max() / want.abs().max())
        assert err <= tol, err
    dx0 = torch.randn(B, n, M, generator=gen, dtype=torch.float64)
    d_em = torch.zeros_like(emd)
    d_sw = torch.ones_like(swd)
    ops.se_atten_embed_grad(d_em, d_sw, dx0.to("cuda", dtype), gs, dgs, tt, paird, swd)
    want_ds = (dx0.reshape(-1, M) * (1 + ttp * sw.reshape(-1, 1)) * g1).sum(1)
    want_dsw = 1 + (dx0.reshape(-1, M) * g0 * ttp).sum(1)
    assert float((d_em[:, :, 0].double().cpu().reshape(-1) - want_ds).abs().max() / want_ds.abs().max()) <= tol * 10
    assert float((d_sw.double().cpu().reshape(-1) - want_dsw).abs().max() / want_dsw.abs().max()) <= tol * 10
    assert float(d_em[:, :, 1:].abs().max()) == 0.0
    # rhat
    rhat, rinv = ops.se_atten_rhat(emd)
    emr = em.clone().requires_grad_(True)
    want = torch.nn.functional.normalize(emr[:, :, 1:4], dim=-1)
    assert float((rhat.double().cpu() - want.detach()).abs().max()) <= tol
    d_rhat = torch.randn(B, n, 3, generator=gen, dtype=torch.float64)
    (want * d_rhat).sum().backward()
    d_em2 = torch.zeros_like(emd)
    ops.se_atten_rhat_grad(d_em2, d_rhat.to("cuda", dtype), rhat, rinv)
    ref = emr.grad[:, :, 1:]
    live = em[:, :, 1:].abs().sum(-1) > 0  # (zero rows: both sides give g / eps = 1e12 g, compared separately)
    assert float((d_em2[:, :, 1:].double().cpu() - ref)[live].abs().max() / ref[live].abs().max()) <= tol * 10
    assert float((d_em2[:, :, 1:].double().cpu() - ref)[~live].abs().max() / ref[~live].abs().max()) <= tol * 10
    assert float(d_em2[:, :, 0].abs().max()) == 0.0
