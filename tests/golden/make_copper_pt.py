"""Energy, forces and virial of the COPPER benchmark model (BASELINE config 3: one type, sel 512, rcut 8.0 / 2.0,
neuron [25, 50, 100], axis 16) from the reference's PyTorch backend (autograd, uncompressed), run HERE:

    python tests/golden/make_copper_pt.py

FCC box of 5^3 conventional cells (500 atoms, a0 3.615 A, Gaussian jitter 0.05 A): a near-perfect lattice whose forces
are sums with ~1000-fold cancellation -- the case that set the operand precision of the tensor-core fitting net.
Embedding net: the reference's default initialisation (seed 1), stored in the fixture; fitting net: the closed-form
weights of make_sea_compress.fit_weights (not stored).  Writes tests/golden/copper_pt.json."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_dpa1_strip as base  # noqa: E402
from make_sea_compress import BIAS_ATOM_E, FIT_NEURON, fit_weights  # noqa: E402


def main():
    base.import_reference()
    from deepmd.dpmodel.descriptor.se_e2_a import DescrptSeA
    from deepmd.dpmodel.fitting.ener_fitting import EnergyFittingNet
    from deepmd.dpmodel.utils.nlist import extend_input_and_build_neighbor_list

    sys.path.insert(0, base.ROOT)
    import __graft_entry__ as g

    stats = (0.06, 0.12, 0.07)
    sel, rcut, rcut_smth = [512], 8.0, 2.0
    dp = DescrptSeA(rcut=rcut, rcut_smth=rcut_smth, sel=sel, neuron=[25, 50, 100], axis_neuron=16, type_one_side=True,
                    seed=1, precision="float64")
    dp.davg[0, :, 0], dp.davg[0, :, 1:] = stats[0], 0.0
    dp.dstd[0, :, 0], dp.dstd[0, :, 1:] = stats[1], stats[2]
    net = dp.embeddings[(0,)]
    embed = [[[np.asarray(l.w).tolist() for l in net.layers], [np.asarray(l.b).tolist() for l in net.layers]]]
    coord, atype, box = g.copper_box(5, 0.05)
    ext_c, ext_t, mapping, nlist = extend_input_and_build_neighbor_list(
        coord.reshape(1, -1, 3), atype.reshape(1, -1).astype(np.int64), rcut, sel, mixed_types=False, box=box.reshape(1, 3, 3))
    desc = np.asarray(dp.call(ext_c, ext_t, nlist, mapping)[0])[0]
    fit = EnergyFittingNet(ntypes=1, dim_descrpt=1600, neuron=list(FIT_NEURON), resnet_dt=True, mixed_types=False, seed=1,
                           precision="float64")
    fit.bias_atom_e[...] = np.array([[BIAS_ATOM_E[0]]])
    layers, head = fit_weights(0, 1600)
    fnet = fit.nets[(0,)]
    for layer, (w, b, idt) in zip(fnet.layers[:-1], layers):
        layer.w[...] = w
        layer.b[...] = b
        if idt is not None:
            layer.idt[...] = idt
    fnet.layers[-1].w[...] = head[0]
    fnet.layers[-1].b[...] = head[1]
    e_atom = np.asarray(fit.call(desc[None], atype.reshape(1, -1).astype(np.int64))["energy"]).reshape(-1)
    # (the PyTorch shell model of base.pt_energy_force_virial has two types; a one-type model needs its own type map)
    efv = base.pt_energy_force_virial(dp, fit, coord, atype, box, e_atom, type_map=["Cu"])
    f = np.array(efv["pt_force"])
    data = dict(config=dict(ntypes=1, sel=sel, rcut=rcut, rcut_smth=rcut_smth, neuron=[25, 50, 100], axis_neuron=16,
                            stats=[stats], min_nbor_dist=2.0, ncell=5, jitter=0.05),
                embed=embed,
                expected=dict(atomic_energy=e_atom.tolist(), numneigh=(np.asarray(nlist)[0] >= 0).sum(1).tolist(), **efv))
    path = os.path.join(HERE, "copper_pt.json")
    with open(path, "w") as fh:
        json.dump(data, fh)
    print("wrote", path, os.path.getsize(path), "bytes; max |F|", float(np.abs(f).max()), "mean neighbours",
          float((np.asarray(nlist)[0] >= 0).sum(1).mean()))


if __name__ == "__main__":
    main()
