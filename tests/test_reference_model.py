"""The reference's own end-to-end test model (source/tests/infer/deeppot_sea.pth, expected values of
source/tests/infer/deeppot-testcase.yaml; fixture tests/golden/deeppot_sea.json written by
tests/golden/make_deeppot_sea.py): model-level conventions -- per-slot davg / dstd, type_one_side net indexing,
fitting resnet_dt / idt, bias_atom_e, the /nnei scaling -- checked against values produced by the reference itself.
The model is uncompressed there; here its embedding nets are tabulated (compress.py = `dp compress`), which the
stride-0.01 quintic table reproduces to ~1e-13."""
import json
import os

import numpy as np
import pytest
import torch

import __graft_entry__ as g
from oracle import cpu as ocpu
from oracle import pipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def data():
    with open(os.path.join(ROOT, "tests", "golden", "deeppot_sea.json")) as f:
        return json.load(f)


def _case(c):
    coord = np.array(c["coord"], np.float64).reshape(-1, 3)
    return coord, np.array(c["atype"], np.int32), np.array(c["box"], np.float64).reshape(3, 3)


def rel(a, b):
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.abs(np.asarray(a, np.float64).reshape(-1) - b).max() / np.abs(b).max())


def test_reference_model_through_the_cpu_pipeline(data):
    """CPU: reference CPU ops (oracle) + the loader + the table builder against the reference's expected values."""
    g.load_package()
    from deepmd_kit_b200.model import SeAModel

    model = SeAModel.from_reference(data, torch.float64, "cpu")
    assert model.cfg.sel == (46, 92) and model.M == 12 and model.cfg.axis_neuron == 2
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    assert len(data["cases"]) >= 2
    for c in data["cases"]:
        coord, atype, box = _case(c)
        L = np.diag(box)
        lists = pipeline.build_lists(lib, coord - np.floor(coord / L) * L, atype, box, model.cfg.rcut + 2.0)
        e, f, v, ex = pipeline.evaluate(lib, model, lists)
        assert abs(e - c["energy"]) <= 1e-12 * abs(c["energy"])
        assert rel(ex["atom_energy"], c["atomic_energy"]) <= 1e-12
        assert rel(f, c["force"]) <= 1e-10
        assert rel(v, c["virial"]) <= 1e-10


@pytest.mark.gpu
def test_reference_model_on_the_gpu(data):
    """GPU: DeepPotB200.eval of the loaded model against the reference's expected energy / force / virial, 1e-10."""
    g.load_package()
    from deepmd_kit_b200.model import DeepPotB200, SeAModel

    dp = DeepPotB200(SeAModel.from_reference(data, torch.float64, "cuda:0"), skin=2.0)
    for c in data["cases"]:
        coord, atype, box = _case(c)
        e, f, v, ae, av = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype, atomic=True)
        assert abs(e[0, 0] - c["energy"]) <= 1e-10 * abs(c["energy"])
        assert rel(ae[0], c["atomic_energy"]) <= 1e-10
        assert rel(f[0], c["force"]) <= 1e-10
        assert rel(v[0], c["virial"]) <= 1e-10


# ---------------------------------------------------------------------------------------------------------------
# se_atten (DPA-1 strip / smooth, attn_layer 0): descriptor rows produced by the reference's own NumPy backend
# (deepmd/dpmodel/descriptor/dpa1.py DescrptDPA1.call; fixture tests/golden/dpa1_strip.json written by
# tests/golden/make_dpa1_strip.py).  Pins the MODEL composition -- one type-agnostic section with per-type statistics,
# tebd_idx = centre * (ntypes + 1) + neighbour, strip net on [tebd(neighbour), tebd(centre)], gate
# gg_s * (1 + gg_t * sw), /nnei, GR^T GR[:, :axis], centre type embedding appended -- to the reference itself.  The
# reference evaluates the embedding net exactly; here it is tabulated (stride 0.01 quintic): measured 3.4e-15 (CPU).
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dpa1():
    with open(os.path.join(ROOT, "tests", "golden", "dpa1_strip.json")) as f:
        return json.load(f)


def _atten_model(dpa1, device):
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

    c = dpa1["config"]
    cfg = SeAttenConfig(ntypes=c["ntypes"], nsel=c["sel"], rcut=c["rcut"], rcut_smth=c["rcut_smth"],
                        neuron=tuple(c["neuron"]), axis_neuron=c["axis_neuron"], tebd_dim=c["tebd_dim"],
                        stats=tuple(tuple(s) for s in c["stats"]), fitting_neuron=tuple(c["fitting_neuron"]),
                        fitting_resnet_dt=bool(c["fitting_resnet_dt"]))
    return SeAttenModel(cfg, torch.float64, device, weights=dpa1["weights"])


def _descriptor_rows(model, xyz, ctype):
    """[D | tebd(centre)] from the gated table output xyz [nloc, 4, M] (what the fitting net is fed with)."""
    cfg = model.cfg
    xs = xyz / cfg.nnei
    d = torch.matmul(xs.permute(0, 2, 1), xs[:, :, :cfg.axis_neuron]).reshape(xyz.shape[0], -1)
    return torch.cat([d, model.tebd.to(d.device)[ctype]], 1)


def test_se_atten_composition_matches_reference_backend_cpu(dpa1):
    g.load_package()
    model = _atten_model(dpa1, "cpu")
    from oracle import pipeline_atten

    coord, atype, box = g.water_box(1, 0.0)
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    lists = pipeline.build_lists(lib, coord, atype, box, model.cfg.rcut + 2.0)
    e, f_cpu, v_cpu, ex = pipeline_atten.evaluate(lib, model, lists)
    exp = dpa1["expected"]
    assert ((ex["nlist"] >= 0).sum(1) == np.array(exp["numneigh"])).all()
    got = _descriptor_rows(model, torch.as_tensor(ex["xyz"]), torch.as_tensor(atype.astype(np.int64))).numpy()
    want = np.array(exp["descriptor"])
    assert got.shape[1] == want.shape[1] == 1608
    assert rel(got[exp["rows"]], want) <= 1e-12  # measured 3.4e-15
    assert abs(got.sum() - exp["total"]) <= 1e-12 * abs(exp["total"])
    assert abs((got * got).sum() - exp["total_sq"]) <= 1e-12 * exp["total_sq"]
    # ... and the energies of the reference's EnergyFittingNet on top (mixed types, resnet_dt / idt, bias_atom_e)
    assert rel(ex["atom_energy"], exp["atomic_energy"]) <= 1e-12
    assert abs(e - exp["energy"]) <= 1e-12 * abs(exp["energy"])
    # ... and energy / forces / virial of the reference's PyTorch backend (autograd through the uncompressed model)
    assert abs(e - exp["pt_energy"]) <= 1e-12 * abs(exp["pt_energy"])
    assert rel(f_cpu, exp["pt_force"]) <= 1e-10
    assert rel(v_cpu, exp["pt_virial"]) <= 1e-10


@pytest.mark.gpu
def test_se_atten_composition_matches_reference_backend_gpu(dpa1):
    """The product kernels (env-mat, gate scalars, gated table forward) on the same weights."""
    pkg = g.load_package()
    ops = pkg.ops
    from deepmd_kit_b200.model import DeepPotB200

    model = _atten_model(dpa1, "cuda:0")
    cfg = model.cfg
    coord, atype, box = g.water_box(1, 0.0)
    dp = DeepPotB200(model, skin=2.0, use_graph=False)
    c = torch.as_tensor(coord).to("cuda:0")
    t = torch.as_tensor(atype).to("cuda:0")
    st = dp.build_neighbors(c, t, box)
    ext_c = (c.reshape(-1, 3).index_select(0, st.map64) + st.shift).contiguous()
    nloc, nall, nnei = st.nloc, int(st.ext_type.numel()), cfg.nnei
    em, dv, rij, nlist = ops.prod_env_mat_a(ext_c.reshape(-1), st.ext_type, st.numneigh, st.rows, model.davg, model.dstd,
                                            nloc, nall, cfg.rcut, cfg.rcut_smth, cfg.sec,
                                            f_type=torch.zeros_like(st.ext_type))
    pair, sw, _ = ops.se_atten_gate_scalars(nlist, st.ext_type, rij, nloc, nnei, cfg.ntypes, cfg.rcut_smth, cfg.rcut)
    em3 = em.reshape(nloc, nnei, 4)
    em_x = em3[:, :, 0].reshape(-1, 1).contiguous()
    xyz = ops.tabulate_fusion_se_atten_gate(model.table, model.info, em_x, em3, model.tt_full, pair, sw, model.M)
    got = _descriptor_rows(model, xyz, t.long()).cpu().numpy()
    exp = dpa1["expected"]
    assert rel(got[exp["rows"]], np.array(exp["descriptor"])) <= 1e-10
    assert abs(got.sum() - exp["total"]) <= 1e-10 * abs(exp["total"])
    # the whole model through the public API (fitting net on the tcgen05 kernels): the reference's energies
    assert model.use_tc
    eg, fg, vg, ae, _ = DeepPotB200(model, skin=2.0).eval(coord.reshape(1, -1), box.reshape(1, 9), atype, atomic=True)
    assert rel(ae[0].reshape(-1), exp["atomic_energy"]) <= 1e-10
    assert abs(eg[0, 0] - exp["energy"]) <= 1e-10 * abs(exp["energy"])
    assert np.abs(fg[0].sum(0)).max() <= 1e-9 * np.abs(fg[0]).max()  # no net force
    assert rel(fg[0], exp["pt_force"]) <= 1e-10   # the reference's own autograd forces and virial
    assert rel(vg[0], exp["pt_virial"]) <= 1e-10


def test_dp_compress_restatement_matches_reference_table(dpa1):
    """compress.py against the table the reference's own DescrptDPA1.enable_compression builds for the same embedding
    net and statistics: identical table_info / row count (the upper boundary comes from the angular components of slot
    0: 15, not 9), the tabulated quintics equal where it matters (values at three points of every row, summed per row
    and per channel: 1e-12), and the type-pair table of the strip net (tt_full) element-wise."""
    g.load_package()
    model = _atten_model(dpa1, "cpu")
    cz = dpa1["compress"]
    assert model.cfg.min_nbor_dist == cz["min_nbor_dist"]
    info = model.info.numpy()
    assert np.array_equal(info, np.array(cz["table_info"]))
    tab = model.table64.numpy()
    assert tab.shape[0] == cz["nrow"]
    first = int((info[1] - info[0]) / info[3])
    h = np.where(np.arange(tab.shape[0]) < first, info[3], info[4])[:, None]
    a = tab.reshape(tab.shape[0], -1, 6)
    for srec in cz["sums"]:
        x = srec["frac"] * h
        v = a[:, :, 0] + (a[:, :, 1] + (a[:, :, 2] + (a[:, :, 3] + (a[:, :, 4] + a[:, :, 5] * x) * x) * x) * x) * x
        assert rel(v.sum(1), srec["per_row"]) <= 1e-12
        assert rel(v.sum(0), srec["per_channel"]) <= 1e-12
    assert rel(model.tt_full.numpy(), cz["tt_full"]) <= 1e-14


def test_se_a_compress_restatement_matches_reference_tables():
    """compress_se_a (type_one_side: one table per NEIGHBOUR type, range over all centre types with sel > 0) on the
    benchmark configuration against the reference's own DescrptSeA.enable_compression for the same embedding nets
    (fixture tests/golden/sea_compress.json, written by tests/golden/make_sea_compress.py)."""
    g.load_package()
    from deepmd_kit_b200.model import SeAConfig, SeAModel

    with open(os.path.join(ROOT, "tests", "golden", "sea_compress.json")) as f:
        d = json.load(f)
    c = d["config"]
    cfg = SeAConfig(ntypes=len(c["sel"]), sel=tuple(c["sel"]), rcut=c["rcut"], rcut_smth=c["rcut_smth"],
                    neuron=tuple(c["neuron"]), axis_neuron=c["axis_neuron"], min_nbor_dist=c["min_nbor_dist"])
    nnei = sum(c["sel"])
    davg = np.zeros((cfg.ntypes, nnei, 4))
    dstd = np.ones((cfg.ntypes, nnei, 4))
    for t, (a0, s0, s1) in enumerate(c["stats"]):
        davg[t, :, 0], dstd[t, :, 0], dstd[t, :, 1:] = a0, s0, s1
    model = SeAModel(cfg, torch.float64, "cpu", weights=dict(davg=davg, dstd=dstd, embed=d["embed"]))
    for t, want in enumerate(d["tables"]):
        info = model.infos[t].numpy()
        assert np.array_equal(info, np.array(want["table_info"]))
        tab = model.tables64[t].numpy()
        assert tab.shape[0] == want["nrow"]
        first = int((info[1] - info[0]) / info[3])
        h = np.where(np.arange(tab.shape[0]) < first, info[3], info[4])[:, None]
        a = tab.reshape(tab.shape[0], -1, 6)
        for srec in want["sums"]:
            x = srec["frac"] * h
            v = a[:, :, 0] + (a[:, :, 1] + (a[:, :, 2] + (a[:, :, 3] + (a[:, :, 4] + a[:, :, 5] * x) * x) * x) * x) * x
            assert rel(v.sum(0), srec["per_channel"]) <= 1e-12
            assert abs(v.sum() - srec["total"]) <= 1e-12 * abs(srec["total"])


def test_se_a_descriptor_matches_reference_backend():
    """Benchmark-size se_e2_a (type_one_side, sel [46, 92], M 100, axis 16): the compressed descriptor through the CPU
    checker pipeline against the reference's UNCOMPRESSED DescrptSeA.call on the 192-atom water frame."""
    g.load_package()
    from deepmd_kit_b200.model import SeAConfig, SeAModel

    with open(os.path.join(ROOT, "tests", "golden", "sea_compress.json")) as f:
        d = json.load(f)
    c = d["config"]
    cfg = SeAConfig(ntypes=len(c["sel"]), sel=tuple(c["sel"]), rcut=c["rcut"], rcut_smth=c["rcut_smth"],
                    neuron=tuple(c["neuron"]), axis_neuron=c["axis_neuron"], min_nbor_dist=c["min_nbor_dist"])
    nnei = sum(c["sel"])
    davg = np.zeros((cfg.ntypes, nnei, 4))
    dstd = np.ones((cfg.ntypes, nnei, 4))
    for t, (a0, s0, s1) in enumerate(c["stats"]):
        davg[t, :, 0], dstd[t, :, 0], dstd[t, :, 1:] = a0, s0, s1
    model = SeAModel(cfg, torch.float64, "cpu", weights=dict(davg=davg, dstd=dstd, embed=d["embed"]))
    coord, atype, box = g.water_box(1, 0.0)
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    lists = pipeline.build_lists(lib, coord, atype, box, cfg.rcut + 2.0)
    _, _, _, ex = pipeline.evaluate(lib, model, lists)
    exp = d["descriptor"]
    assert ((ex["nlist"] >= 0).sum(1) == np.array(exp["numneigh"])).all()
    xs = torch.as_tensor(ex["xyz"]) / cfg.nnei
    got = torch.matmul(xs.permute(0, 2, 1), xs[:, :, :cfg.axis_neuron]).reshape(len(atype), -1).numpy()
    assert rel(got[exp["rows"]], exp["values"]) <= 1e-12
    assert abs(got.sum() - exp["total"]) <= 1e-12 * abs(exp["total"])
    assert abs((got * got).sum() - exp["total_sq"]) <= 1e-12 * exp["total_sq"]


def _sea_benchmark_model(device):
    """Benchmark-size se_e2_a model with the fixture's embedding nets and the closed-form fitting nets of
    tests/golden/make_sea_compress.py (fit_weights is imported from the generator: the fixture does not store them)."""
    import importlib.util

    from deepmd_kit_b200.model import SeAConfig, SeAModel

    spec = importlib.util.spec_from_file_location("_make_sea_compress", os.path.join(ROOT, "tests", "golden",
                                                                                      "make_sea_compress.py"))
    with open(os.path.join(ROOT, "tests", "golden", "sea_compress.json")) as f:
        d = json.load(f)
    # only fit_weights / constants are needed: parse them without executing the generator's reference imports
    src = open(spec.origin).read()
    ns = {"np": np}
    start = src.index("FIT_NEURON = ")
    exec(src[start:src.index("def main():")], ns)
    c = d["config"]
    cfg = SeAConfig(ntypes=len(c["sel"]), sel=tuple(c["sel"]), rcut=c["rcut"], rcut_smth=c["rcut_smth"],
                    neuron=tuple(c["neuron"]), axis_neuron=c["axis_neuron"], min_nbor_dist=c["min_nbor_dist"],
                    fitting_neuron=tuple(ns["FIT_NEURON"]), fitting_resnet_dt=True)
    nnei = sum(c["sel"])
    davg = np.zeros((cfg.ntypes, nnei, 4))
    dstd = np.ones((cfg.ntypes, nnei, 4))
    for t, (a0, s0, s1) in enumerate(c["stats"]):
        davg[t, :, 0], dstd[t, :, 0], dstd[t, :, 1:] = a0, s0, s1
    fits = []
    for t in range(cfg.ntypes):
        layers, head = ns["fit_weights"](t, 1600)
        fits.append(dict(layers=[[w, b, idt] for w, b, idt in layers], head=[head[0], head[1]]))
    model = SeAModel(cfg, torch.float64, device, weights=dict(davg=davg, dstd=dstd, embed=d["embed"], fit=fits,
                                                             bias_atom_e=list(ns["BIAS_ATOM_E"])))
    return model, d["descriptor"]


def test_se_a_benchmark_model_matches_reference_pytorch_backend_cpu():
    """Energy, forces and virial of the reference's PyTorch backend (autograd, UNCOMPRESSED benchmark-size se_e2_a +
    per-type fitting nets) on the 192-atom water frame, reproduced by the compressed CPU checker pipeline."""
    g.load_package()
    model, exp = _sea_benchmark_model("cpu")
    coord, atype, box = g.water_box(1, 0.0)
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    lists = pipeline.build_lists(lib, coord, atype, box, model.cfg.rcut + 2.0)
    e, f, v, ex = pipeline.evaluate(lib, model, lists)
    assert rel(ex["atom_energy"], exp["atomic_energy"]) <= 1e-12
    assert abs(e - exp["pt_energy"]) <= 1e-12 * abs(exp["pt_energy"])
    assert rel(f, exp["pt_force"]) <= 1e-10
    assert rel(v, exp["pt_virial"]) <= 1e-10


@pytest.mark.gpu
def test_se_a_benchmark_model_matches_reference_pytorch_backend_gpu():
    g.load_package()
    from deepmd_kit_b200.model import DeepPotB200

    model, exp = _sea_benchmark_model("cuda:0")
    assert model.use_tc
    coord, atype, box = g.water_box(1, 0.0)
    e, f, v, ae, _ = DeepPotB200(model, skin=2.0).eval(coord.reshape(1, -1), box.reshape(1, 9), atype, atomic=True)
    assert rel(ae[0].reshape(-1), exp["atomic_energy"]) <= 1e-10
    assert abs(e[0, 0] - exp["pt_energy"]) <= 1e-10 * abs(exp["pt_energy"])
    assert rel(f[0], exp["pt_force"]) <= 1e-10
    assert rel(v[0], exp["pt_virial"]) <= 1e-10


def _copper_model(device):
    import importlib.util  # noqa: F401

    from deepmd_kit_b200.model import SeAConfig, SeAModel

    with open(os.path.join(ROOT, "tests", "golden", "copper_pt.json")) as f:
        d = json.load(f)
    src = open(os.path.join(ROOT, "tests", "golden", "make_sea_compress.py")).read()
    ns = {"np": np}
    exec(src[src.index("FIT_NEURON = "):src.index("def main():")], ns)
    c = d["config"]
    cfg = SeAConfig(ntypes=1, sel=tuple(c["sel"]), rcut=c["rcut"], rcut_smth=c["rcut_smth"], neuron=tuple(c["neuron"]),
                    axis_neuron=c["axis_neuron"], min_nbor_dist=c["min_nbor_dist"], stats=[tuple(c["stats"][0])],
                    fitting_neuron=tuple(ns["FIT_NEURON"]), fitting_resnet_dt=True)
    nnei = sum(c["sel"])
    a0, s0, s1 = c["stats"][0]
    davg = np.zeros((1, nnei, 4))
    dstd = np.ones((1, nnei, 4))
    davg[0, :, 0], dstd[0, :, 0], dstd[0, :, 1:] = a0, s0, s1
    layers, head = ns["fit_weights"](0, 1600)
    fits = [dict(layers=[[w, b, idt] for w, b, idt in layers], head=[head[0], head[1]])]
    model = SeAModel(cfg, torch.float64, device, weights=dict(davg=davg, dstd=dstd, embed=d["embed"], fit=fits,
                                                             bias_atom_e=[ns["BIAS_ATOM_E"][0]]))
    return model, d


def test_copper_model_matches_reference_pytorch_backend_cpu():
    """BASELINE config 3 model (one type, sel 512, rcut 8) on a 500-atom FCC box: forces of a near-perfect lattice are
    sums with ~1000-fold cancellation (max |F| = 1.4e-4 against per-neighbour terms of 1e-1).  Reference = its PyTorch
    backend (autograd, uncompressed)."""
    g.load_package()
    model, d = _copper_model("cpu")
    exp, c = d["expected"], d["config"]
    coord, atype, box = g.copper_box(c["ncell"], c["jitter"])
    lib = ocpu.CpuLib("reference" if ocpu.available("reference") else "port")
    lists = pipeline.build_lists(lib, coord, atype, box, model.cfg.rcut + 2.0)
    e, f, v, ex = pipeline.evaluate(lib, model, lists)
    assert ((ex["nlist"] >= 0).sum(1) == np.array(exp["numneigh"])).all()
    assert rel(ex["atom_energy"], exp["atomic_energy"]) <= 1e-12
    assert abs(e - exp["pt_energy"]) <= 1e-12 * abs(exp["pt_energy"])
    assert rel(f, exp["pt_force"]) <= 1e-10
    assert rel(v, exp["pt_virial"]) <= 1e-10


@pytest.mark.gpu
def test_copper_model_matches_reference_pytorch_backend_gpu():
    g.load_package()
    from deepmd_kit_b200.model import DeepPotB200

    model, d = _copper_model("cuda:0")
    exp, c = d["expected"], d["config"]
    coord, atype, box = g.copper_box(c["ncell"], c["jitter"])
    e, f, v, ae, _ = DeepPotB200(model, skin=2.0).eval(coord.reshape(1, -1), box.reshape(1, 9), atype, atomic=True)
    assert rel(ae[0].reshape(-1), exp["atomic_energy"]) <= 1e-10
    assert abs(e[0, 0] - exp["pt_energy"]) <= 1e-10 * abs(exp["pt_energy"])
    assert rel(f[0], exp["pt_force"]) <= 1e-10
    assert rel(v[0], exp["pt_virial"]) <= 1e-10
