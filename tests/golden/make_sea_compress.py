"""`dp compress` of a se_e2_a descriptor by the reference's own NumPy backend (deepmd/dpmodel/descriptor/se_e2_a.py:
DescrptSeA.enable_compression -> deepmd/utils/tabulate.py + tabulate_math.py), run HERE where /root/reference exists:

    python tests/golden/make_sea_compress.py

Benchmark configuration (examples/water/se_e2_a: rcut 6.0 / 0.5, sel [46, 92], neuron [25, 50, 100], type_one_side,
default initialisation with seed 1) with the env-mat statistics of deepmd_kit_b200.model.WATER_STATS.  Writes
tests/golden/sea_compress.json: the embedding-net weights per NEIGHBOUR type (the table of net `filter_-1_net_<t>`),
table_info and row count of every table, the tabulated quintics evaluated at three points of every row, summed
over rows (per channel), and rows of the UNCOMPRESSED descriptor of the 192-atom water frame (DescrptSeA.call) -- see make_dpa1_strip.py for why values and not coefficients.  Import stand-ins:
tests/golden/_ref_shims (same as make_dpa1_strip.py)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_dpa1_strip import REF  # noqa: E402,F401
import make_dpa1_strip as base  # noqa: E402


FIT_NEURON = (16, 16, 16)
BIAS_ATOM_E = (-1.5, 0.7)


def fit_weights(t, dim_in, neuron=FIT_NEURON):
    """Closed-form weights of the fitting net of centre type t: [(w, b, idt | None)] per hidden layer and (w, b) of
    the head.  Used by this generator AND by tests/test_reference_model.py, so the fixture does not carry them."""
    layers = []
    n_in = dim_in
    for k, n_out in enumerate(neuron):
        i = np.arange(n_in, dtype=np.float64)[:, None]
        j = np.arange(n_out, dtype=np.float64)[None, :]
        w = np.sin(0.37 * i + 1.3 * j + 0.71 * t + k) / np.sqrt(n_in + n_out)
        b = 0.1 * np.cos(j[0] + t + k)
        idt = (0.1 + 0.01 * np.sin(j[0] + k)) if n_in == n_out else None
        layers.append((w, b, idt))
        n_in = n_out
    i = np.arange(n_in, dtype=np.float64)[:, None]
    return layers, (0.3 * np.cos(0.9 * i + t), np.array([0.05 * t]))


def main():
    base.import_reference()
    from deepmd.dpmodel.descriptor.se_e2_a import DescrptSeA

    stats = ((0.05033, 0.13984, 0.08580), (0.04810, 0.12388, 0.07672))
    sel = [46, 92]
    dp = DescrptSeA(rcut=6.0, rcut_smth=0.5, sel=sel, neuron=[25, 50, 100], axis_neuron=16, type_one_side=True, seed=1,
                    precision="float64")
    for t, (a0, s0, s1) in enumerate(stats):
        dp.davg[t, :, 0] = a0
        dp.davg[t, :, 1:] = 0.0
        dp.dstd[t, :, 0] = s0
        dp.dstd[t, :, 1:] = s1
    embed = []
    for t in range(len(sel)):
        net = dp.embeddings[(t,)]
        for layer in net.layers:
            assert layer.idt is None and layer.activation_function == "tanh"
        embed.append([[np.asarray(l.w).tolist() for l in net.layers], [np.asarray(l.b).tolist() for l in net.layers]])
    # the UNCOMPRESSED descriptor of the 192-atom water frame (DescrptSeA.call: exact embedding nets)
    from deepmd.dpmodel.utils.nlist import extend_input_and_build_neighbor_list

    sys.path.insert(0, base.ROOT)
    import __graft_entry__ as g

    coord, atype, box = g.water_box(1, 0.0)
    ext_c, ext_t, mapping, nlist = extend_input_and_build_neighbor_list(
        coord.reshape(1, -1, 3), atype.reshape(1, -1).astype(np.int64), 6.0, sel, mixed_types=False, box=box.reshape(1, 3, 3))
    desc = np.asarray(dp.call(ext_c, ext_t, nlist, mapping)[0])[0]
    assert desc.shape == (len(atype), 1600)
    rows = [0, 1, 63, 64, 130, 191]
    descriptor = dict(rows=rows, values=desc[rows].tolist(), total=float(desc.sum()), total_sq=float((desc * desc).sum()),
                      numneigh=(np.asarray(nlist)[0] >= 0).sum(1).tolist())
    # whole model: per-type energy fitting nets with CLOSED-FORM weights (fit_weights below: the test rebuilds them, the
    # fixture need not store them), energy / forces / virial from the reference's PyTorch backend (autograd)
    from deepmd.dpmodel.fitting.ener_fitting import EnergyFittingNet

    fit = EnergyFittingNet(ntypes=len(sel), dim_descrpt=1600, neuron=list(FIT_NEURON), resnet_dt=True, mixed_types=False,
                           seed=1, precision="float64")
    fit.bias_atom_e[...] = np.array(BIAS_ATOM_E).reshape(-1, 1)
    for t in range(len(sel)):
        net = fit.nets[(t,)]
        layers, head = fit_weights(t, 1600)
        for layer, (w, b, idt) in zip(net.layers[:-1], layers):
            assert layer.w.shape == w.shape and (layer.idt is None) == (idt is None)
            layer.w[...] = w
            layer.b[...] = b
            if idt is not None:
                layer.idt[...] = idt
        net.layers[-1].w[...] = head[0]
        net.layers[-1].b[...] = head[1]
    e_atom = np.asarray(fit.call(desc[None], atype.reshape(1, -1).astype(np.int64))["energy"]).reshape(-1)
    efv = base.pt_energy_force_virial(dp, fit, coord, atype, box, e_atom)
    descriptor.update(atomic_energy=e_atom.tolist(), energy=float(e_atom.sum()), **efv)
    dp.enable_compression(0.9, 5, 0.01, 0.1, -1)
    tables = []
    for t in range(len(sel)):
        info = np.asarray(dp.compress_info[t], np.float64)
        tab = np.asarray(dp.compress_data[t], np.float64)
        nrow = tab.shape[0]
        first = int((info[1] - info[0]) / info[3])
        h = np.where(np.arange(nrow) < first, info[3], info[4])[:, None]
        a = tab.reshape(nrow, -1, 6)
        sums = []
        for frac in (0.0, 0.5, 0.999):
            x = frac * h
            v = a[:, :, 0] + (a[:, :, 1] + (a[:, :, 2] + (a[:, :, 3] + (a[:, :, 4] + a[:, :, 5] * x) * x) * x) * x) * x
            sums.append(dict(frac=frac, per_channel=v.sum(0).tolist(), total=float(v.sum())))
        tables.append(dict(table_info=info.tolist(), nrow=int(nrow), sums=sums))
    data = dict(config=dict(rcut=6.0, rcut_smth=0.5, sel=sel, neuron=[25, 50, 100], axis_neuron=16, stats=stats,
                            min_nbor_dist=0.9),
                embed=embed, tables=tables, descriptor=descriptor)
    path = os.path.join(HERE, "sea_compress.json")
    with open(path, "w") as f:
        json.dump(data, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
