"""One launch of each hot operator on a mid-size water box, for `ncu` (see tools/README in DESIGN.md §5).
usage: python tools/prof_ops.py [ncopy=8] [dtype=f64]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.load_package()
from deepmd_kit_b200 import ops
from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

ncopy = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dtype = torch.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else torch.float32
dev = torch.device("cuda:0")
cfg = SeAConfig()
model = SeAModel(cfg, dtype, dev)
coord, atype, box = g.water_box(ncopy, 0.01)
c = torch.as_tensor(coord).to(dev, dtype); t = torch.as_tensor(atype).to(dev)
dp = DeepPotB200(model, use_graph=False)
st = dp.build_neighbors(c, t, box)
nloc = st.nloc; nall = st.ext_type.numel()
ext_c = (c.reshape(-1, 3).index_select(0, st.map64) + st.shift).contiguous()
nnei, M = cfg.nnei, model.M
for rep in range(2):
    em, dv, rij, nlist = ops.prod_env_mat_a(ext_c.reshape(-1), st.ext_type, st.numneigh, st.rows, model.davg, model.dstd,
                                            nloc, nall, cfg.rcut, cfg.rcut_smth, cfg.sec)
    xyz = ops.tabulate_sections_fwd(model.tables, model.infos, em, cfg.sec, M)
    if model.use_split:
        ops.tabulate_sections_desc(model.tables, model.infos, em, cfg.sec, M, cfg.axis_neuron, 1.0 / nnei,
                                   desc_row=st.type_inv, mode=2, nslice=model.nslice, pad_rows=32,
                                   flags=model.coef_flags)
    dy = torch.randn_like(xyz)
    nd = ops.tabulate_sections_grad(model.tables, model.infos, em, dy, cfg.sec, M)  # full table
    nd = ops.tabulate_sections_grad(model.tables, model.infos, em, dy, cfg.sec, M, flags=model.coef_flags)
    gd = torch.randn(nloc, M * cfg.axis_neuron, dtype=dtype, device=dev)
    ops.se_a_descriptor_grad(gd, xyz, cfg.axis_neuron, 1.0 / nnei, rows=st.type_perm.to(torch.int32), out=torch.empty_like(xyz))
    nl2 = nlist.clone(); ops.use_nlist_map(nl2, st.mapping)
    ops.prod_force_virial_a(nd, dv, rij, nl2, nloc, nloc, nnei)
torch.cuda.synchronize()
print("done", nloc)
