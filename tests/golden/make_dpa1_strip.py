"""Golden vectors for the se_atten (DPA-1 strip / smooth, attn_layer 0) MODEL composition, produced by the reference's
own NumPy backend (deepmd/dpmodel/descriptor/dpa1.py: DescrptDPA1.call), run HERE where /root/reference exists:

    python tests/golden/make_dpa1_strip.py

The reference package is imported from the read-only tree.  Four things it imports at module level are absent from
this image (no network): `array_api_compat`, `h5py`, `wcmatch`, `lmdb`, and the compiled `deepmd.lib` package.  None of
them takes part in a CPU evaluation of the descriptor; tests/golden/_ref_shims holds import stand-ins (NumPy >= 2
provides the array-API functions the backend calls).  Nothing of the reference is copied: the script builds the model
of examples/water/se_atten_compressible/input.json (se_atten_v2 = strip + smooth type embedding; sel 120, rcut 6.0 /
0.5, neuron [25, 50, 100], axis 16, tebd_dim 8, two-side) with the reference's default initialisation (seed 1), gives it
the env-mat statistics of deepmd_kit_b200.atten.SeAttenConfig, evaluates the 192-atom water frame and writes
tests/golden/dpa1_strip.json:
  weights   : davg / dstd, geometric embedding net, two-side strip net, type-embedding table (padding row last)
  expected  : descriptor rows [D (100 x 16) | tebd(centre)] of eight atoms, sum and sum of squares of all rows,
              per-atom neighbour counts, and the atomic energies of the reference's EnergyFittingNet
              (deepmd/dpmodel/fitting/ener_fitting.py; mixed types, resnet_dt, a REDUCED width [16, 16, 16] so that
              the fixture stays small -- the conventions it pins do not depend on the width -- and a non-zero
              bias_atom_e) on that descriptor
"""
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("DEEPMD_SOURCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def import_reference():
    sys.path.insert(0, os.path.join(HERE, "_ref_shims"))
    sys.path.insert(0, REF)
    import deepmd

    lib = types.ModuleType("deepmd.lib")
    lib.__path__ = [os.path.join(HERE, "_ref_shims", "dplib")]
    sys.modules["deepmd.lib"] = lib
    deepmd.lib = lib
    from deepmd.dpmodel.descriptor.dpa1 import DescrptDPA1
    from deepmd.dpmodel.fitting.ener_fitting import EnergyFittingNet
    from deepmd.dpmodel.utils.nlist import extend_input_and_build_neighbor_list

    return DescrptDPA1, extend_input_and_build_neighbor_list, EnergyFittingNet


def pt_energy_force_virial(dp_descriptor, dp_fitting, coord, atype, box, e_atom_numpy, type_map=("O", "H")):
    """deepmd.pt EnergyModel with the given (NumPy-backend) descriptor and fitting net: E, F = -dE/dr, virial."""
    import torch
    from deepmd.pt.model.descriptor.base_descriptor import BaseDescriptor
    from deepmd.pt.model.model import get_model
    from deepmd.pt.model.task.base_fitting import BaseFitting

    d_ser, f_ser = dp_descriptor.serialize(), dp_fitting.serialize()
    params = {"type_map": list(type_map),
              "descriptor": {"type": "se_e2_a", "sel": [4] * len(type_map), "rcut": 6.0, "rcut_smth": 0.5, "neuron": [2, 4],
                             "axis_neuron": 2},
              "fitting_net": {"neuron": [4], "resnet_dt": True}}
    model = get_model(params).to("cpu")  # a shell; descriptor and fitting net are replaced below
    model.atomic_model.descriptor = BaseDescriptor.deserialize(d_ser).to("cpu")
    model.atomic_model.fitting_net = BaseFitting.deserialize(f_ser).to("cpu")
    model.atomic_model.rcut = dp_descriptor.get_rcut()
    model.atomic_model.sel = dp_descriptor.get_sel()
    c = torch.tensor(coord.reshape(1, -1), dtype=torch.float64, requires_grad=True)
    out = model(c, torch.tensor(atype.reshape(1, -1), dtype=torch.int64), box=torch.tensor(box.reshape(1, 9)))
    ea = out["atom_energy"].detach().numpy().reshape(-1)
    assert np.abs(ea - e_atom_numpy).max() <= 1e-12 * np.abs(e_atom_numpy).max(), "PyTorch and NumPy backends disagree"
    return dict(pt_energy=float(out["energy"].detach().numpy().reshape(-1)[0]),
                pt_force=out["force"].detach().numpy().reshape(-1, 3).tolist(),
                pt_virial=out["virial"].detach().numpy().reshape(9).tolist())


def net_weights(net):
    out = []
    for layer in net.layers:
        assert layer.idt is None and layer.activation_function == "tanh"
        out.append(dict(w=np.asarray(layer.w).tolist(), b=np.asarray(layer.b).tolist(), resnet=bool(layer.resnet)))
    return out


def main():
    DescrptDPA1, build, EnergyFittingNet = import_reference()
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    coord, atype, box = g.water_box(1, 0.0)
    # statistics as in deepmd_kit_b200.atten.SeAttenConfig.stats (type 0 / 1: davg0, dstd radial, dstd angular)
    stats = ((0.05033, 0.13984, 0.08580), (0.04810, 0.12388, 0.07672))
    rcut, rcut_smth, sel, ntypes, tebd_dim = 6.0, 0.5, 120, 2, 8
    dp = DescrptDPA1(rcut=rcut, rcut_smth=rcut_smth, sel=sel, ntypes=ntypes, neuron=[25, 50, 100], axis_neuron=16,
                     tebd_dim=tebd_dim, tebd_input_mode="strip", attn_layer=0, smooth_type_embedding=True,
                     type_one_side=False, concat_output_tebd=True, seed=1, precision="float64")
    blk = dp.se_atten
    davg = np.zeros((ntypes, sel, 4))
    dstd = np.ones((ntypes, sel, 4))
    for t, (a0, s0, s1) in enumerate(stats):
        davg[t, :, 0] = a0
        dstd[t, :, 0] = s0
        dstd[t, :, 1:] = s1
    blk.mean[...] = davg
    blk.stddev[...] = dstd
    c = coord.reshape(1, -1, 3)
    ext_c, ext_t, mapping, nlist = build(c, atype.reshape(1, -1).astype(np.int64), rcut, [sel], mixed_types=True,
                                        box=box.reshape(1, 3, 3))
    out = dp.call(ext_c, ext_t, nlist, mapping)
    desc = np.asarray(out[0])[0]  # [nloc, 100*16 + 8]
    assert desc.shape == (len(atype), 1608)
    tebd = np.asarray(dp.type_embedding.call())  # [ntypes + 1, tebd_dim], padding row last
    fit_neuron = [16, 16, 16]
    fit = EnergyFittingNet(ntypes=ntypes, dim_descrpt=desc.shape[1], neuron=fit_neuron, resnet_dt=True, mixed_types=True,
                           seed=1, precision="float64")
    fit.bias_atom_e[...] = np.array([[-1.5], [0.7]])
    e_atom = np.asarray(fit.call(desc[None], atype.reshape(1, -1).astype(np.int64))["energy"]).reshape(-1)
    net = fit.nets[()]
    layers = [dict(w=np.asarray(l.w).tolist(), b=np.asarray(l.b).tolist(),
                   idt=None if l.idt is None else np.asarray(l.idt).tolist()) for l in net.layers[:-1]]
    head = net.layers[-1]
    assert head.activation_function == "none" and head.idt is None and not head.resnet
    rows = [0, 1, 2, 63, 64, 65, 100, 191]
    data = dict(
        config=dict(rcut=rcut, rcut_smth=rcut_smth, sel=sel, ntypes=ntypes, neuron=[25, 50, 100], axis_neuron=16,
                    tebd_dim=tebd_dim, stats=stats, fitting_neuron=fit_neuron, fitting_resnet_dt=True),
        weights=dict(embed=net_weights(blk.embeddings[0]), strip=net_weights(blk.embeddings_strip[0]),
                     tebd=tebd.tolist(),
                     fit=dict(layers=layers, head=dict(w=np.asarray(head.w).tolist(), b=np.asarray(head.b).tolist())),
                     bias_atom_e=np.asarray(fit.bias_atom_e).reshape(-1).tolist()),
        expected=dict(rows=rows, descriptor=desc[rows].tolist(), total=float(desc.sum()),
                      total_sq=float((desc * desc).sum()),
                      numneigh=(np.asarray(nlist)[0] >= 0).sum(1).tolist(),
                      atomic_energy=e_atom.tolist(), energy=float(e_atom.sum())),
    )
    # Energy, forces and virial of the whole model from the reference's PyTorch backend (autograd): the NumPy objects
    # above are handed over through the reference's own cross-backend serialisation.
    data["expected"].update(pt_energy_force_virial(dp, fit, coord, atype, box, e_atom))
    # `dp compress` by the reference itself (deepmd/utils/tabulate.py + tabulate_math.py through
    # DescrptDPA1.enable_compression): table_info and the tabulated quintics.  The high-order coefficients come from
    # differences divided by stride^3..5 and carry amplified rounding noise (two builds of the reference differ in
    # a5 at the 1e-5 level), so the fixture stores what the table MEANS: the quintic evaluated at three points of
    # every row, summed over channels (per row) and over rows (per channel).
    dp.enable_compression(0.9, 5, 0.01, 0.1, -1)
    info = np.asarray(blk.compress_info[0], np.float64)
    tab = np.asarray(blk.compress_data[0], np.float64)
    nrow = tab.shape[0]
    first = int((info[1] - info[0]) / info[3])
    h = np.where(np.arange(nrow) < first, info[3], info[4])[:, None]
    a = tab.reshape(nrow, -1, 6)
    sums = []
    for frac in (0.0, 0.5, 0.999):
        x = frac * h
        v = a[:, :, 0] + (a[:, :, 1] + (a[:, :, 2] + (a[:, :, 3] + (a[:, :, 4] + a[:, :, 5] * x) * x) * x) * x) * x
        sums.append(dict(frac=frac, per_row=v.sum(1).tolist(), per_channel=v.sum(0).tolist()))
    data["compress"] = dict(min_nbor_dist=0.9, table_info=info.tolist(), nrow=int(nrow), sums=sums,
                            tt_full=np.asarray(blk.type_embd_data).tolist())
    path = os.path.join(HERE, "dpa1_strip.json")
    with open(path, "w") as f:
        json.dump(data, f)
    print("wrote", path, os.path.getsize(path), "bytes; |D| max", float(np.abs(desc).max()))


if __name__ == "__main__":
    main()
