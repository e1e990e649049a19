"""BASELINE config 5 at the operator level: tabulate_fusion_se_atten (+ grad) on a 526 848-atom water box
(14^3 replicas), sel 120 type-mixed, M = 100, two_embed [nloc*nnei, M] materialised as the reference op requires.
usage: python tools/atten_bench.py [ncopy=14] [dtype=f64]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.load_package()
from deepmd_kit_b200 import ops
from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

ncopy = int(sys.argv[1]) if len(sys.argv) > 1 else 14
dtype = torch.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else torch.float32
dev = torch.device("cuda:0")
cfg = SeAConfig(ntypes=1, sel=(120,), stats=[(0.05, 0.13, 0.08)])
model = SeAModel(cfg, dtype, dev)
coord, atype, box = g.water_box(ncopy, 0.01)
atype[:] = 0  # type-mixed neighbour list of se_atten: one section, ordered by distance
c = torch.as_tensor(coord).to(dev, dtype); t = torch.as_tensor(atype).to(dev)
dp = DeepPotB200(model, use_graph=False)
st = dp.build_neighbors(c, t, box)
nloc = st.nloc; nall = st.ext_type.numel()
ext_c = (c.reshape(-1, 3).index_select(0, st.map64) + st.shift).contiguous()
em, dv, rij, nlist = ops.prod_env_mat_a(ext_c.reshape(-1), st.ext_type, st.numneigh, st.rows, model.davg, model.dstd,
                                        nloc, nall, cfg.rcut, cfg.rcut_smth, cfg.sec)
del dv, rij
nnei, M = cfg.nnei, model.M
nreal = float((nlist >= 0).sum().item()) / nloc
em3 = em.reshape(nloc, nnei, 4)
em_x = em3[:, :, 0].reshape(-1, 1).contiguous()
two = torch.randn(nloc * nnei, M, dtype=dtype, device=dev) * 0.1
F = 8 if dtype == torch.float64 else 4
print(f"natoms {nloc} nnei {nnei} real nbrs {nreal:.1f} two_embed {two.numel() * F / 1e9:.1f} GB dtype {dtype}")


def timeit(name, fn, nbytes):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); med = ts[len(ts) // 2]
    print(f"{name:34s} median {med:8.3f} ms ({med * 1e6 / nloc:7.2f} ns/atom)  {nbytes / med / 1e6:8.1f} GB/s algorithmic", flush=True)


tab, info = model.tables[0], model.infos[0]
# algorithmic bytes: the op must read the materialised two_embed of the REAL neighbours at least
rd = nloc * (nreal * M * F + nnei * 5 * F)
timeit("tabulate_fusion_se_atten fwd", lambda: ops.tabulate_fusion_se_a(tab, info, em_x, em3, M, two_embed=two), rd + nloc * 4 * M * F)
dy = torch.randn(nloc, 4, M, dtype=dtype, device=dev)
timeit("tabulate_fusion_se_atten grad", lambda: ops.tabulate_fusion_se_a_grad(tab, info, em_x, em3, dy, M, two_embed=two),
       2 * rd + nloc * (4 * M * F + nnei * 5 * F))
timeit("tabulate_fusion_se_a fwd (no gate)", lambda: ops.tabulate_fusion_se_a(tab, info, em_x, em3, M), nloc * nnei * 5 * F)
