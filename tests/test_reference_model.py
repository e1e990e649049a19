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
