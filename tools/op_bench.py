"""Stand-alone timing of every dpb200 operator on the benchmark workload (CUDA events, median of N).
usage: python tools/op_bench.py [ncopy=12] [dtype=f64] [reps=15]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.load_package()
from deepmd_kit_b200 import ops
from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

if os.environ.get("OPB_BLASLT"):
    torch.backends.cuda.preferred_blas_library("cublaslt")
ncopy = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dtype = torch.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else torch.float32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 15
dev = torch.device("cuda:0")
cfg = SeAConfig()
model = SeAModel(cfg, dtype, dev)
coord, atype, box = g.water_box(ncopy, 0.01)
c = torch.as_tensor(coord).to(dev, dtype); t = torch.as_tensor(atype).to(dev)
dp = DeepPotB200(model, use_graph=False)
st = dp.build_neighbors(c, t, box)
nloc = st.nloc; nall = st.ext_type.numel()
ext_c = (c.reshape(-1, 3).index_select(0, st.map64) + st.shift).contiguous()

def timeit(name, fn, bytes_per_atom=None, flops_per_atom=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); med = ts[len(ts) // 2]
    extra = ""
    if bytes_per_atom: extra += f"  {bytes_per_atom * nloc / med / 1e6:8.1f} GB/s"
    if flops_per_atom: extra += f"  {flops_per_atom * nloc / med / 1e9:8.2f} TFLOP/s"
    print(f"{name:28s} median {med:8.3f} ms  min {ts[0]:8.3f}  ({med * 1e6 / nloc:7.2f} ns/atom){extra}", flush=True)
    return med

F = 8 if dtype == torch.float64 else 4
nnei, M = cfg.nnei, model.M
em, dv, rij, nlist = ops.prod_env_mat_a(ext_c.reshape(-1), st.ext_type, st.numneigh, st.rows, model.davg, model.dstd,
                                        nloc, nall, cfg.rcut, cfg.rcut_smth, cfg.sec)
nreal = float((nlist >= 0).sum().item()) / nloc
raw = float(st.numneigh.sum().item()) / nloc
print(f"natoms {nloc} nall {nall} real nbrs {nreal:.1f} raw {raw:.1f} dtype {dtype}")
timeit("format_nlist", lambda: ops.format_nlist(ext_c.reshape(-1), st.ext_type, st.numneigh, st.rows, nloc, nall, cfg.rcut, cfg.sec))
timeit("prod_env_mat_a", lambda: ops.prod_env_mat_a(ext_c.reshape(-1), st.ext_type, st.numneigh, st.rows, model.davg, model.dstd,
                                                  nloc, nall, cfg.rcut, cfg.rcut_smth, cfg.sec),
       bytes_per_atom=19 * nnei * F + 4 * nnei + 4 * raw + 3 * F * (1 + nall / nloc))
xyz = ops.tabulate_sections_fwd(model.tables, model.infos, em, cfg.sec, M)
timeit("tabulate_sections_fwd", lambda: ops.tabulate_sections_fwd(model.tables, model.infos, em, cfg.sec, M),
       flops_per_atom=18 * (nreal + 2) * M)
dy = torch.randn_like(xyz)
timeit("tabulate_sections_grad", lambda: ops.tabulate_sections_grad(model.tables, model.infos, em, dy, cfg.sec, M),
       flops_per_atom=36 * (nreal + 2) * M)
timeit("tabulate_sections_desc(split)", lambda: ops.tabulate_sections_desc(
    model.tables, model.infos, em, cfg.sec, M, cfg.axis_neuron, 1.0 / nnei, mode=2, nslice=model.nslice, pad_rows=32,
    flags=model.coef_flags), flops_per_atom=18 * (nreal + 2) * M)
if getattr(model, "coef_flags", None):
    timeit("tabulate_sections_grad(cm)", lambda: ops.tabulate_sections_grad(
        model.tables, model.infos, em, dy, cfg.sec, M, flags=model.coef_flags), flops_per_atom=36 * (nreal + 2) * M)
if os.environ.get("OPB_ONLY") == "tab":
    sys.exit(0)
nd = ops.tabulate_sections_grad(model.tables, model.infos, em, dy, cfg.sec, M)
nl2 = nlist.clone(); ops.use_nlist_map(nl2, st.mapping)
timeit("prod_force_virial_a", lambda: ops.prod_force_virial_a(nd, dv, rij, nl2, nloc, nloc, nnei),
       bytes_per_atom=19 * nnei * F + 4 * nnei + 3 * F)
timeit("prod_force_a", lambda: ops.prod_force_a(nd, dv, nl2, nloc, nloc, nnei), bytes_per_atom=16 * nnei * F + 4 * nnei + 3 * F)
timeit("prod_virial_a(+atom)", lambda: ops.prod_virial_a(nd, dv, rij, nl2, nloc, nloc, nnei),
       bytes_per_atom=19 * nnei * F + 4 * nnei + 9 * F)
perm32 = st.type_perm.to(torch.int32)
timeit("se_a_descriptor", lambda: ops.se_a_descriptor(xyz, cfg.axis_neuron, 1.0 / nnei, rows=perm32[: min(nloc, 1 << 17)]))
nn = min(nloc, 1 << 17)
gd = torch.randn(nn, M * cfg.axis_neuron, dtype=dtype, device=dev)
out = torch.empty_like(xyz)
timeit("se_a_descriptor_grad", lambda: ops.se_a_descriptor_grad(gd, xyz, cfg.axis_neuron, 1.0 / nnei, rows=perm32[:nn], out=out))
timeit("energy_and_dy (plain GEMMs)", lambda: model.energy_and_dy(xyz, st.type_perm, st.type_ranges))
if model.use_split:
    inv = 1.0 / nnei
    timeit("tabulate_sections_desc(split)", lambda: ops.tabulate_sections_desc(
        model.tables, model.infos, em, cfg.sec, M, cfg.axis_neuron, inv, desc_row=st.type_inv, mode=2,
        nslice=model.nslice, pad_rows=32, flags=model.coef_flags), flops_per_atom=18 * (nreal + 2) * M)
    _, desc, rexp = ops.tabulate_sections_desc(model.tables, model.infos, em, cfg.sec, M, cfg.axis_neuron, inv,
                                               desc_row=st.type_inv, mode=2, nslice=model.nslice, pad_rows=32)
    timeit("fit fwd+bwd split, 1 chunk", lambda: model.fit[0].forward_backward_split(
        desc, None if rexp is None else rexp[:nn], nn))
    dd = ops.se_a_descriptor(xyz, cfg.axis_neuron, inv, rows=perm32[:nn])
    timeit("fit fwd+bwd plain, 1 chunk", lambda: model.fit[0].forward_backward(dd))
    timeit("energy_and_dy_split", lambda: model.energy_and_dy_split(xyz, desc, rexp, st.type_perm, st.type_ranges))
    e0, _, dy0 = model.energy_and_dy(xyz, st.type_perm, st.type_ranges)
    e1, _, dy1 = model.energy_and_dy_split(xyz, desc, rexp, st.type_perm, st.type_ranges)
    print(f"split vs plain: dE/E = {abs((e1 - e0) / e0).item():.3e}  max|d dy|/max|dy| = "
          f"{((dy1 - dy0).abs().max() / dy0.abs().max()).item():.3e}", flush=True)
timeit("build_neighbors", lambda: dp.build_neighbors(c, t, box))
