"""Golden vectors for DPA-1 WITH attention layers (se_atten_v2 = strip + smooth, attn_layer 2,
attn 128, attn_dotr, normalised q / k / v), produced by the reference's NumPy backend
(deepmd/dpmodel/descriptor/dpa1.py: DescrptDPA1.call -> NeighborGatedAttention) and its PyTorch backend (autograd
forces / virial; deepmd/pt/model/descriptor/se_atten.py:1058-1447), run HERE where /root/reference exists:

    python tests/golden/make_dpa1_attn.py

Same frame, statistics, import stand-ins and fitting net as make_dpa1_strip.py (which this script imports its helpers
from); the layer norms get non-trivial scale / shift so that their conventions are pinned too.  Writes
tests/golden/dpa1_attn.npz (binary: ~100 k weights would be 2.5 MB as JSON):
  config_json                         : the model hyper-parameters
  w_*                                 : embedding / strip / type-embedding / attention / fitting weights
  x_descriptor [8, 1608], x_rows, x_total, x_total_sq, x_numneigh, x_atomic_energy, x_energy   (NumPy backend)
  x_gg_rows [2, 120, 100]             : the attention output g2 of atoms 0 and 100
  x_pt_energy, x_pt_force [192, 3], x_pt_virial [9]                                           (PyTorch backend)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_dpa1_strip as base  # noqa: E402


def main():
    DescrptDPA1, build, EnergyFittingNet = base.import_reference()
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    coord, atype, box = g.water_box(1, 0.0)
    stats = ((0.05033, 0.13984, 0.08580), (0.04810, 0.12388, 0.07672))
    rcut, rcut_smth, sel, ntypes, tebd_dim, attn, nlayer = 6.0, 0.5, 120, 2, 8, 128, 2
    dp = DescrptDPA1(rcut=rcut, rcut_smth=rcut_smth, sel=sel, ntypes=ntypes, neuron=[25, 50, 100], axis_neuron=16,
                     tebd_dim=tebd_dim, tebd_input_mode="strip", attn=attn, attn_layer=nlayer, attn_dotr=True,
                     attn_mask=False, normalize=True, smooth_type_embedding=True, type_one_side=False,
                     concat_output_tebd=True, seed=1, precision="float64")
    blk = dp.se_atten
    davg = np.zeros((ntypes, sel, 4))
    dstd = np.ones((ntypes, sel, 4))
    for t, (a0, s0, s1) in enumerate(stats):
        davg[t, :, 0] = a0
        dstd[t, :, 0] = s0
        dstd[t, :, 1:] = s1
    blk.mean[...] = davg
    blk.stddev[...] = dstd
    rng = np.random.default_rng(7)
    layers = blk.dpa1_attention.attention_layers
    assert len(layers) == nlayer
    for lay in layers:
        ln = lay.attn_layer_norm
        ln.w = 1.0 + 0.2 * rng.standard_normal(np.asarray(ln.w).shape)
        ln.b = 0.1 * rng.standard_normal(np.asarray(ln.b).shape)
    c = coord.reshape(1, -1, 3)
    ext_c, ext_t, mapping, nlist = build(c, atype.reshape(1, -1).astype(np.int64), rcut, [sel], mixed_types=True,
                                        box=box.reshape(1, 3, 3))
    out = dp.call(ext_c, ext_t, nlist, mapping)
    desc = np.asarray(out[0])[0]
    te = np.asarray(dp.type_embedding.call())
    _, gg, _, _, _ = blk(nlist, ext_c, ext_t, te[np.asarray(ext_t).reshape(-1)].reshape(1, -1, tebd_dim), mapping=None,
                         type_embedding=te)
    gg = np.asarray(gg).reshape(len(atype), sel, 100)  # g2 = the attention output [nloc, nnei, ng]
    assert desc.shape == (len(atype), 1608) and gg.shape == (len(atype), sel, 100)
    tebd = np.asarray(dp.type_embedding.call())
    fit_neuron = [16, 16, 16]
    fit = EnergyFittingNet(ntypes=ntypes, dim_descrpt=desc.shape[1], neuron=fit_neuron, resnet_dt=True, mixed_types=True,
                           seed=1, precision="float64")
    fit.bias_atom_e[...] = np.array([[-1.5], [0.7]])
    e_atom = np.asarray(fit.call(desc[None], atype.reshape(1, -1).astype(np.int64))["energy"]).reshape(-1)
    pt = base.pt_energy_force_virial(dp, fit, coord, atype, box, e_atom)
    rows = [0, 1, 2, 63, 64, 65, 100, 191]
    arrays = {}

    def put_net(prefix, net):
        for li, l in enumerate(net.layers):
            assert l.idt is None and l.activation_function == "tanh"
            arrays[f"w_{prefix}_{li}_w"] = np.asarray(l.w, np.float64)
            arrays[f"w_{prefix}_{li}_b"] = np.asarray(l.b, np.float64)

    put_net("embed", blk.embeddings[0])
    put_net("strip", blk.embeddings_strip[0])
    arrays["w_tebd"] = tebd
    for li, lay in enumerate(layers):
        ga = lay.attention_layer
        assert ga.num_heads == 1 and ga.normalize and ga.dotr and ga.smooth and not ga.do_mask
        for name, mlp in (("in", ga.in_proj), ("out", ga.out_proj)):
            assert mlp.idt is None and not mlp.resnet and mlp.activation_function in (None, "none", "linear")
            arrays[f"w_attn_{li}_{name}_w"] = np.asarray(mlp.w, np.float64)
            arrays[f"w_attn_{li}_{name}_b"] = np.asarray(mlp.b, np.float64)
        arrays[f"w_attn_{li}_ln_w"] = np.asarray(lay.attn_layer_norm.w, np.float64).reshape(-1)
        arrays[f"w_attn_{li}_ln_b"] = np.asarray(lay.attn_layer_norm.b, np.float64).reshape(-1)
        scaling, ln_eps = float(ga.scaling), float(lay.attn_layer_norm.eps)
    net = fit.nets[()]
    for li, l in enumerate(net.layers[:-1]):
        arrays[f"w_fit_{li}_w"] = np.asarray(l.w, np.float64)
        arrays[f"w_fit_{li}_b"] = np.asarray(l.b, np.float64)
        arrays[f"w_fit_{li}_idt"] = np.asarray(l.idt, np.float64)
    head = net.layers[-1]
    arrays["w_fit_head_w"] = np.asarray(head.w, np.float64)
    arrays["w_fit_head_b"] = np.asarray(head.b, np.float64)
    arrays["w_bias_atom_e"] = np.asarray(fit.bias_atom_e, np.float64).reshape(-1)
    cfg = dict(rcut=rcut, rcut_smth=rcut_smth, sel=sel, ntypes=ntypes, neuron=[25, 50, 100], axis_neuron=16,
               tebd_dim=tebd_dim, stats=stats, fitting_neuron=fit_neuron, fitting_resnet_dt=True, attn=attn,
               attn_layer=nlayer, attn_dotr=True, normalize=True, scaling=scaling, ln_eps=ln_eps, attnw_shift=20.0,
               n_fit_layers=len(net.layers) - 1, n_embed_layers=3)
    arrays.update(
        config_json=np.array(json.dumps(cfg)),
        x_rows=np.array(rows), x_descriptor=desc[rows], x_total=np.array(desc.sum()),
        x_total_sq=np.array((desc * desc).sum()), x_numneigh=(np.asarray(nlist)[0] >= 0).sum(1),
        x_atomic_energy=e_atom, x_energy=np.array(e_atom.sum()), x_gg_rows=gg[[0, 100]],
        x_pt_energy=np.array(pt["pt_energy"]), x_pt_force=np.array(pt["pt_force"]), x_pt_virial=np.array(pt["pt_virial"]))
    path = os.path.join(HERE, "dpa1_attn.npz")
    np.savez(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes; |D| max", float(np.abs(desc).max()), "scaling", scaling,
          "ln_eps", ln_eps)


if __name__ == "__main__":
    main()
