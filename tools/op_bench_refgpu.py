"""Per-operator timing of the dpb200 kernels beside the UNMODIFIED reference CUDA kernels compiled for sm_100 on the
same box (baseline/_ref_gpu/libdeepmd_ref_gpu.so, recipe baseline/Makefile): prod_env_mat_a, tabulate_fusion_se_a
(+ grad) per type section, prod_force_a, prod_virial_a on the benchmark water box.  The reference kernels are the
comparator only: they are never on the product path and are not the parity oracle.
usage: python tools/op_bench_refgpu.py [ncopy=12] [reps=7]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

g.load_package()
from deepmd_kit_b200 import ops
from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = C.CDLL(os.path.join(ROOT, "baseline", "_ref_gpu", "libdeepmd_ref_gpu.so"))
ref.refgpu_last_error.restype = C.c_char_p
ncopy = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = torch.device("cuda:0")
dtype = torch.float64
cfg = SeAConfig()
model = SeAModel(cfg, dtype, dev)
coord, atype, box = g.water_box(ncopy, 0.01)
c = torch.as_tensor(coord).to(dev, dtype)
t = torch.as_tensor(atype).to(dev)
dp = DeepPotB200(model, use_graph=False)
st = dp.build_neighbors(c, t, box)
nloc, nall = st.nloc, st.ext_type.numel()
ext_c = (c.reshape(-1, 3).index_select(0, st.map64) + st.shift).contiguous()
nnei, M, sec = cfg.nnei, model.M, cfg.sec
P = lambda x: C.c_void_p(x.data_ptr())


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def chk(rc):
    if rc != 0:
        raise RuntimeError(ref.refgpu_last_error().decode())


rows = {}
# ---- prod_env_mat_a ---------------------------------------------------------------------------------------------
maxn = int(st.numneigh.max().item())
max_nbor = next(m for m in (256, 512, 1024, 2048, 4096) if m >= maxn)  # prod_env_mat.cc:349-359
cap = st.rows.shape[1]
first = (torch.arange(nloc, dtype=torch.int64, device=dev) * (cap * 4) + st.rows.data_ptr()).contiguous()
ilist = torch.arange(nloc, dtype=torch.int32, device=dev)
numneigh = st.numneigh.to(torch.int32).contiguous()
r_em = torch.empty((nloc, nnei * 4), dtype=dtype, device=dev)
r_dv = torch.empty((nloc, nnei * 12), dtype=dtype, device=dev)
r_rij = torch.empty((nloc, nnei * 3), dtype=dtype, device=dev)
r_nl = torch.empty((nloc, nnei), dtype=torch.int32, device=dev)
a_int = torch.empty(len(sec) + nloc * len(sec) + nloc, dtype=torch.int32, device=dev)
a_ll = torch.empty(nloc * max_nbor * 2, dtype=torch.int64, device=dev)
sec_c = (C.c_int * len(sec))(*sec)
ext_flat = ext_c.reshape(-1)


def ref_env():
    chk(ref.refgpu_prod_env_mat_a_f64(P(r_em), P(r_dv), P(r_rij), P(r_nl), P(ext_flat), P(st.ext_type), P(ilist),
                                      P(numneigh), P(first), P(a_int), P(a_ll), max_nbor, P(model.davg), P(model.dstd),
                                      nloc, nall, C.c_float(cfg.rcut), C.c_float(cfg.rcut_smth), sec_c, len(sec)))


def our_env():
    return ops.prod_env_mat_a(ext_flat, st.ext_type, st.numneigh, st.rows, model.davg, model.dstd, nloc, nall, cfg.rcut,
                              cfg.rcut_smth, cfg.sec)


rows["prod_env_mat_a"] = (timeit(our_env), timeit(ref_env))
em, dv, rij, nlist = our_env()
ref_env()
same_nl = float((r_nl == nlist).float().mean().item())
print(f"natoms {nloc} nall {nall} max raw neighbours {maxn} (reference max_nbor_size {max_nbor}); "
      f"formatted lists identical in {100 * same_nl:.4f} % of the slots; em max |diff| {float((r_em - em).abs().max()):.2e}")
# ---- tabulate forward / backward, per type section as deepmd/pt/model/descriptor/se_a.py:810-841 ------------------
em3 = em.reshape(nloc, nnei, 4)
secs = [(em3[:, sec[k]:sec[k + 1], :].contiguous(), em3[:, sec[k]:sec[k + 1], 0].contiguous()) for k in range(cfg.ntypes)]
infos = [i.to(torch.float64).contiguous() for i in model.infos]
outs = [torch.empty((nloc, 4, M), dtype=dtype, device=dev) for _ in range(cfg.ntypes)]


def ref_fwd():
    for k in range(cfg.ntypes):
        e4, ex = secs[k]
        chk(ref.refgpu_tabulate_fusion_se_a_f64(P(outs[k]), P(model.tables[k]), C.c_void_p(infos[k].data_ptr()), P(ex), P(e4),
                                                nloc, e4.shape[1], M))


rows["tabulate_fusion_se_a (2 sections)"] = (
    timeit(lambda: ops.tabulate_sections_fwd(model.tables, model.infos, em, cfg.sec, M)), timeit(ref_fwd))
xyz = ops.tabulate_sections_fwd(model.tables, model.infos, em, cfg.sec, M)
ref_fwd()
print(f"tabulate fwd max |diff| / max {float(((outs[0] + outs[1]) - xyz).abs().max() / xyz.abs().max()):.2e}")
dy = torch.randn_like(xyz)
gxs = [torch.empty_like(secs[k][1]) for k in range(cfg.ntypes)]
gems = [torch.empty_like(secs[k][0]) for k in range(cfg.ntypes)]


def ref_bwd():
    for k in range(cfg.ntypes):
        e4, ex = secs[k]
        chk(ref.refgpu_tabulate_fusion_se_a_grad_f64(P(gxs[k]), P(gems[k]), P(model.tables[k]),
                                                     C.c_void_p(infos[k].data_ptr()), P(ex), P(e4), P(dy), nloc,
                                                     e4.shape[1], M))


rows["tabulate_fusion_se_a_grad (2 sections)"] = (
    timeit(lambda: ops.tabulate_sections_grad(model.tables, model.infos, em, dy, cfg.sec, M)), timeit(ref_bwd))
# ---- force / virial scatter ------------------------------------------------------------------------------------
nd = torch.randn(nloc, nnei * 4, dtype=dtype, device=dev)
nl_own = nlist.clone()
ops.use_nlist_map(nl_own, st.mapping)
r_f = torch.empty(nloc * 3, dtype=dtype, device=dev)
r_v = torch.empty(9, dtype=dtype, device=dev)
r_av = torch.empty(nloc * 9, dtype=dtype, device=dev)
rows["prod_force_a"] = (timeit(lambda: ops.prod_force_a(nd, dv, nl_own, nloc, nloc, nnei)),
                        timeit(lambda: chk(ref.refgpu_prod_force_a_f64(P(r_f), P(nd), P(dv), P(nl_own), nloc, nloc, nnei))))
rows["prod_virial_a"] = (timeit(lambda: ops.prod_virial_a(nd, dv, rij, nl_own, nloc, nloc, nnei)),
                         timeit(lambda: chk(ref.refgpu_prod_virial_a_f64(P(r_v), P(r_av), P(nd), P(dv), P(rij), P(nl_own), nloc,
                                                                         nloc, nnei))))
f_ours = ops.prod_force_a(nd, dv, nl_own, nloc, nloc, nnei)
print(f"prod_force_a max |diff| / max {float((r_f - f_ours.reshape(-1)).abs().max() / f_ours.abs().max()):.2e}")
t_fused = timeit(lambda: ops.prod_force_virial_a(nd, dv, rij, nl_own, nloc, nloc, nnei))
rows["prod_force_a + prod_virial_a (ours: one fused kernel)"] = (t_fused, rows["prod_force_a"][1] + rows["prod_virial_a"][1])
print(f"{'operator':52s} {'dpb200 ms':>10s} {'reference GPU kernel ms':>24s} {'ratio':>7s}")
for k, (a, b) in rows.items():
    print(f"{k:52s} {a:10.3f} {b:24.3f} {b / a:7.2f}")
print("JSON " + json.dumps({"natoms": nloc, "dtype": "f64", "rows": {k: {"dpb200_ms": a, "reference_gpu_ms": b, "ratio": b / a}
                                                                    for k, (a, b) in rows.items()}}))
