"""Extract the reference's own end-to-end test model into a fixture (run HERE, where /root/reference exists):

    python tests/golden/make_deeppot_sea.py

Reads source/tests/infer/deeppot_sea.pth (TorchScript se_e2_a + energy model, type_one_side, sel [46, 92], neuron
[3, 6, 12], axis 2, fitting [10, 10, 10] with resnet_dt) and the expected atomic energies / forces / atomic virials
of source/tests/infer/deeppot-testcase.yaml (periodic blocks only), and writes tests/golden/deeppot_sea.json:
  descriptor / fitting_net : the model_def_script entries the loader needs
  weights                  : davg, dstd, embedding nets per neighbour type, fitting nets per centre type, bias_atom_e
  cases                    : coord, atype, box, energy, force, virial (sums of the atomic values)
The weights are the reference's test data (category: golden vectors), not code."""
import json
import os
import sys

import numpy as np
import torch
import yaml

REF = os.environ.get("DEEPMD_SOURCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    m = torch.jit.load(os.path.join(REF, "source/tests/infer/deeppot_sea.pth"), map_location="cpu")
    sd = {k: v.detach().double().numpy() for k, v in m.state_dict().items()}
    with open(os.path.join(REF, "source/tests/infer/deeppot-testcase.yaml")) as f:
        tc = yaml.safe_load(f)
    script = tc["model_def_script"]
    desc, fit = script["descriptor"], script["fitting_net"]
    ntypes = len(desc["sel"])
    pre_d = "atomic_model.descriptor.sea."
    pre_f = "atomic_model.fitting_net."
    embed = []
    for t in range(ntypes):
        ws = [sd[f"{pre_d}filter_layers.networks.{t}.layers.{k}.matrix"].tolist() for k in range(len(desc["neuron"]))]
        bs = [sd[f"{pre_d}filter_layers.networks.{t}.layers.{k}.bias"].tolist() for k in range(len(desc["neuron"]))]
        embed.append([ws, bs])
    fits = []
    nl = len(fit["neuron"])
    for t in range(ntypes):
        layers = []
        for k in range(nl):
            key = f"{pre_f}filter_layers.networks.{t}.layers.{k}."
            idt = sd.get(key + "idt")
            layers.append([sd[key + "matrix"].tolist(), sd[key + "bias"].tolist(), None if idt is None else idt.tolist()])
        key = f"{pre_f}filter_layers.networks.{t}.layers.{nl}."
        fits.append(dict(layers=layers, head=[sd[key + "matrix"].tolist(), sd[key + "bias"].tolist()]))
    weights = dict(davg=sd[pre_d + "mean"].tolist(), dstd=sd[pre_d + "stddev"].tolist(), embed=embed, fit=fits,
                   bias_atom_e=sd[pre_f + "bias_atom_e"].reshape(-1).tolist())
    assert float(np.abs(sd["atomic_model.out_bias"]).max()) == 0.0 and float(np.abs(sd["atomic_model.out_std"] - 1).max()) == 0.0
    cases = []
    for blk in tc["results"]:
        if blk.get("box") is None:
            continue  # the non-periodic block: DeepPotB200 evaluates periodic cells
        nat = len(blk["atype"])
        av = np.array(blk["atomic_virial"], np.float64).reshape(nat, 9)
        cases.append(dict(coord=blk["coord"], atype=blk["atype"], box=blk["box"],
                          energy=float(np.sum(blk["atomic_energy"])), atomic_energy=blk["atomic_energy"],
                          force=blk["force"], virial=av.sum(0).tolist()))
    out = dict(source="source/tests/infer/deeppot_sea.pth + deeppot-testcase.yaml",
               descriptor=dict(sel=desc["sel"], rcut=desc["rcut"], rcut_smth=desc["rcut_smth"], neuron=desc["neuron"],
                               axis_neuron=desc["axis_neuron"], resnet_dt=desc["resnet_dt"]),
               fitting_net=dict(neuron=fit["neuron"], resnet_dt=fit["resnet_dt"]), weights=weights, cases=cases)
    with open(os.path.join(HERE, "deeppot_sea.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "deeppot_sea.json"), len(cases), "cases")


if __name__ == "__main__":
    sys.exit(main())
