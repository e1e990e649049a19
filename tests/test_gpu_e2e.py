"""End-to-end parity on the GPU: DeepPotB200.eval (public API, host arrays in/out) against the CPU
pipeline of the reference library on the same boxes, plus size-independent physical properties."""
import numpy as np
import pytest
import torch

import __graft_entry__ as g
from oracle import cpu as ocpu
from oracle import pipeline

pytestmark = pytest.mark.gpu
TOL = {torch.float64: 1e-10, torch.float32: 1e-5}


@pytest.fixture(scope="module")
def pkg():
    return g.load_package()


def _cpu_lib():
    return ocpu.CpuLib("reference" if ocpu.available("reference") else "port")


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("ncopy,jitter", [(1, 0.0), (2, 0.0), (2, 0.01)])
def test_water_e2e_matches_reference_cpu(pkg, dtype, ncopy, jitter):
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    coord, atype, box = g.water_box(ncopy, jitter)
    cfg = SeAConfig()
    dp = DeepPotB200(SeAModel(cfg, dtype, "cuda:0"), skin=2.0)
    e, f, v = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    lib = _cpu_lib()
    np_dt = np.float64 if dtype == torch.float64 else np.float32
    lists = pipeline.build_lists(lib, coord.astype(np_dt), atype, box.astype(np_dt), cfg.rcut + 2.0)
    we, wf, wv, _ = pipeline.evaluate(lib, SeAModel(cfg, dtype, "cpu"), lists)
    tol = TOL[dtype]
    if dtype == torch.float64:
        assert abs(e[0, 0] - we) <= tol * abs(we)
        assert rel(f[0], wf) <= tol
        assert rel(v[0], wv) <= tol
    else:
        # two fp32 evaluations with different summation orders can only agree to fp32 round-off of the
        # O(100)-term force sums; measure both against the fp64 reference pipeline and require the CUDA
        # path to be as accurate as the reference's own fp32 CPU path (and within 1e-5 where that is
        # reachable, i.e. for the energy).
        lists64 = pipeline.build_lists(lib, coord, atype, box, cfg.rcut + 2.0)
        xe, xf, xv, _ = pipeline.evaluate(lib, SeAModel(cfg, torch.float64, "cpu"), lists64)
        msg = (f"energy {abs(e[0, 0] - xe) / abs(xe):.2e} (cpu32 {abs(we - xe) / abs(xe):.2e}) "
               f"force {rel(f[0], xf):.2e} (cpu32 {rel(wf, xf):.2e}) virial {rel(v[0], xv):.2e} (cpu32 {rel(wv, xv):.2e})")
        assert abs(e[0, 0] - xe) <= max(2 * abs(we - xe), 1e-5 * abs(xe)), msg
        assert rel(f[0], xf) <= max(2 * rel(wf, xf), 1e-5), msg
        assert rel(v[0], xv) <= max(2 * rel(wv, xv), 1e-5), msg
    # second call reuses the raw list (ago > 0) and must give the same answer
    e2, f2, v2 = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    assert rel(f2[0], f[0]) <= tol
    # unfused prod_force_a + prod_virial_a path and atomic outputs
    ea, fa, va, ae, av = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype, atomic=True)
    assert rel(fa[0], f[0]) <= tol * 10  # same path, different atomic order
    assert abs(ae.sum() - e[0, 0]) <= 10 * tol * abs(we)
    assert rel(av[0].sum(0), v[0]) <= tol * 100


def test_properties_at_scale(pkg):
    """4x4x4 replica (12288 atoms, exact replication => ties): net force zero, symmetric virial,
    replicas equivalent, energy extensive."""
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    cfg = SeAConfig()
    model = SeAModel(cfg, torch.float64, "cuda:0")
    c1, t1, b1 = g.water_box(1)
    e1, f1, v1 = DeepPotB200(model).eval(c1.reshape(1, -1), b1.reshape(1, 9), t1)
    c4, t4, b4 = g.water_box(4)
    e4, f4, v4 = DeepPotB200(model).eval(c4.reshape(1, -1), b4.reshape(1, 9), t4)
    assert abs(e4[0, 0] - 64 * e1[0, 0]) < 1e-9 * abs(e4[0, 0])
    assert np.abs(f4[0].sum(0)).max() < 1e-9
    assert np.abs(f4[0].reshape(64, 192, 3) - f1[0][None]).max() < 1e-10
    assert np.abs(v4[0] - 64 * v1[0]).max() < 1e-9 * np.abs(v4[0]).max()
    vm = v4[0].reshape(3, 3)
    assert np.abs(vm - vm.T).max() < 1e-9 * np.abs(vm).max()


def test_force_is_energy_gradient(pkg):
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    model = SeAModel(SeAConfig(), torch.float64, "cuda:0")
    coord, atype, box = g.water_box(1)
    dp = DeepPotB200(model, nlist_every=1)
    e0, f0, _ = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    h = 1e-5
    for i, d in ((3, 0), (150, 1)):
        es = []
        for sgn in (1, -1):
            c = coord.copy()
            c[i, d] += sgn * h
            es.append(dp.eval(c.reshape(1, -1), box.reshape(1, 9), atype)[0][0, 0])
        assert abs(-(es[0] - es[1]) / (2 * h) - f0[0, i, d]) < 1e-7


def fcc_box(ncell=6, a0=3.615, jitter=0.05, seed=20260102):
    """BASELINE config 3 in miniature: FCC copper, single type."""
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a0
    cells = np.stack(np.meshgrid(*[np.arange(ncell)] * 3, indexing="ij"), -1).reshape(-1, 3) * a0
    coord = (cells[:, None, :] + base[None]).reshape(-1, 3)
    coord = coord + np.random.default_rng(seed).normal(scale=jitter, size=coord.shape)
    return coord, np.zeros(len(coord), np.int32), np.eye(3) * ncell * a0


@pytest.mark.parametrize("dtype", [torch.float64])
def test_copper_large_sel_e2e(pkg, dtype):
    """Single type, sel [512] (35 % occupancy), rcut 8: the 'large sel, metallic density' case."""
    from deepmd_kit_b200.model import COPPER_CONFIG, DeepPotB200, SeAConfig, SeAModel

    cfg = SeAConfig(**COPPER_CONFIG)
    coord, atype, box = fcc_box()
    dp = DeepPotB200(SeAModel(cfg, dtype, "cuda:0"), skin=2.0)
    e, f, v = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    lib = _cpu_lib()
    lists = pipeline.build_lists(lib, coord, atype, box, cfg.rcut + 2.0)
    we, wf, wv, ex = pipeline.evaluate(lib, SeAModel(cfg, dtype, "cpu"), lists)
    assert (ex["nlist"] >= 0).sum(1).mean() > 150
    assert abs(e[0, 0] - we) <= 1e-10 * abs(we)
    assert rel(f[0], wf) <= 1e-10
    assert rel(v[0], wv) <= 1e-10


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_atom_chunked_evaluation_matches_one_slab(pkg, dtype):
    """Slab-wise evaluation (bounded memory, BASELINE config 3) == the one-slab evaluation."""
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    coord, atype, box = g.water_box(2, 0.01)
    model = SeAModel(SeAConfig(), dtype, "cuda:0")
    ref = DeepPotB200(model, atom_chunk=None, use_graph=False).eval(coord.reshape(1, -1), box.reshape(1, 9), atype,
                                                                    atomic=True)
    for chunk, graph in ((500, False), (1024, True)):
        dp = DeepPotB200(model, atom_chunk=chunk, use_graph=graph)
        got = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype, atomic=True)
        assert dp.state.chunks is not None and len(dp.state.chunks) == -(-len(atype) // chunk)
        tol = 1e-12 if dtype == torch.float64 else 2e-5
        for a, b in zip(got, ref):
            assert rel(a, b) <= tol
        got2 = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)  # graph replay / list reuse
        assert rel(got2[1], ref[1]) <= tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_properties_at_full_benchmark_size(pkg, dtype):
    """BASELINE config 2 at its full size (20^3 replicas = 1 536 000 atoms, variant A: exact replication, so every
    distance tie of the formatter is exercised 8000 times): energy extensive, every replica carries the forces of the
    192-atom cell, net force zero, virial extensive and symmetric.  Runs the production path (CUDA graph, fused
    descriptor epilogue, split-operand fitting net, tensor-core backward)."""
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~80 GB of free device memory")
    tol = 1e-10 if dtype == torch.float64 else 2e-5
    model = SeAModel(SeAConfig(), dtype, "cuda:0")
    c1, t1, b1 = g.water_box(1)
    e1, f1, v1 = DeepPotB200(model).eval(c1.reshape(1, -1), b1.reshape(1, 9), t1)
    n = 20
    c, t, b = g.water_box(n)
    dp = DeepPotB200(model)
    e, f, v = dp.eval(c.reshape(1, -1), b.reshape(1, 9), t)
    nrep = n ** 3
    assert abs(e[0, 0] - nrep * e1[0, 0]) <= 10 * tol * abs(e[0, 0])
    fmax = np.abs(f1[0]).max()
    # fp32: coordinates up to 249 A carry 1.5e-5 A of representation error (the reference forms rij in FPTYPE too),
    # which the stiff O-H bonds turn into ~1e-3 eV/A force differences between replicas
    ftol = tol if dtype == torch.float64 else 2e-3
    assert np.abs(f[0].reshape(nrep, 192, 3) - f1[0][None]).max() <= ftol * fmax
    assert np.abs(f[0].astype(np.float64).sum(0)).max() <= ftol * fmax * nrep ** 0.5 * 10
    vtol = 10 * tol if dtype == torch.float64 else 2e-3
    assert np.abs(v[0] - nrep * v1[0]).max() <= vtol * np.abs(v[0]).max()
    vm = v[0].reshape(3, 3)
    assert np.abs(vm - vm.T).max() <= vtol * np.abs(vm).max()
    # second evaluation replays the graph on the reused list
    e2, f2, _ = dp.eval(c.reshape(1, -1), b.reshape(1, 9), t)
    assert np.abs(f2[0] - f[0]).max() <= ftol * fmax
    del dp
    torch.cuda.empty_cache()


def test_eval_frames_are_independent(pkg):
    """DeepPot.eval semantics: every frame has its own cell and coordinates.  Frame 1 rescales the cell (NPT step),
    frame 2 moves atoms by more than skin / 2, frame 3 by a few hundredths of an Angstrom (the raw list and the
    ghost shifts may be reused there).  Each frame must match the CPU pipeline on that frame alone."""
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    cfg = SeAConfig()
    c0, atype, b0 = g.water_box(2, 0.01)
    rng = np.random.default_rng(7)
    frames = [(c0, b0), (c0 * 1.02, b0 * 1.02), (c0 + rng.normal(scale=0.6, size=c0.shape), b0)]
    frames.append((frames[2][0] + rng.normal(scale=0.02, size=c0.shape), b0))
    dp = DeepPotB200(SeAModel(cfg, torch.float64, "cuda:0"), skin=2.0)
    coords = np.stack([c.reshape(-1) for c, _ in frames])
    cells = np.stack([b.reshape(9) for _, b in frames])
    e, f, v = dp.eval(coords, cells, atype)
    lib = _cpu_lib()
    cpu_model = SeAModel(cfg, torch.float64, "cpu")
    for k, (c, b) in enumerate(frames):
        L = np.diag(b)
        cw = c - np.floor(c / L) * L
        lists = pipeline.build_lists(lib, cw, atype, b, cfg.rcut + 2.0)
        we, wf, wv, _ = pipeline.evaluate(lib, cpu_model, lists)
        assert abs(e[k, 0] - we) <= 1e-10 * abs(we), k
        assert rel(f[k], wf) <= 1e-10, k
        assert rel(v[k], wv) <= 1e-10, k
    assert dp.state.ago == 2  # frame 3 reused frame 2's list


def test_water_12288_matches_reference_cpu(pkg):
    """4^3 replicas of the 192-atom frame (12 288 atoms, the size of the CPU baseline sample) against the reference
    CPU library end to end, 1e-10."""
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    cfg = SeAConfig()
    coord, atype, box = g.water_box(4, 0.01)
    dp = DeepPotB200(SeAModel(cfg, torch.float64, "cuda:0"), skin=2.0)
    e, f, v = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    lib = _cpu_lib()
    lists = pipeline.build_lists(lib, coord, atype, box, cfg.rcut + 2.0)
    we, wf, wv, _ = pipeline.evaluate(lib, SeAModel(cfg, torch.float64, "cpu"), lists)
    assert abs(e[0, 0] - we) <= 1e-10 * abs(we)
    assert rel(f[0], wf) <= 1e-10
    assert rel(v[0], wv) <= 1e-10


def test_copper_10976_matches_reference_cpu(pkg):
    """14^3 FCC cells (10 976 atoms, sel 512, rcut 8) against the reference CPU library end to end, 1e-10: forces of
    a near-perfect lattice are sums with heavy cancellation, the hardest case for the split-integer fitting GEMMs."""
    from deepmd_kit_b200.model import COPPER_CONFIG, DeepPotB200, SeAConfig, SeAModel

    cfg = SeAConfig(**COPPER_CONFIG)
    coord, atype, box = fcc_box(ncell=14)
    dp = DeepPotB200(SeAModel(cfg, torch.float64, "cuda:0"), skin=2.0)
    e, f, v = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    lib = _cpu_lib()
    lists = pipeline.build_lists(lib, coord, atype, box, cfg.rcut + 2.0)
    we, wf, wv, _ = pipeline.evaluate(lib, SeAModel(cfg, torch.float64, "cpu"), lists)
    assert abs(e[0, 0] - we) <= 1e-10 * abs(we)
    assert rel(f[0], wf) <= 1e-10
    assert rel(v[0], wv) <= 1e-10


def test_domain_path_on_one_gpu_matches_plain(pkg):
    """DomainDeepPot with a 1 x 1 x 1 grid: every halo direction is a periodic self-image, so halo_pack, the local
    exchange and halo_unpack_add (the multi-GPU data path) run on a single GPU and must reproduce DeepPotB200."""
    from deepmd_kit_b200.domain import DomainDeepPot
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    cfg = SeAConfig()
    coord, atype, box = g.water_box(2, 0.01)
    model = SeAModel(cfg, torch.float64, "cuda:0")
    e0, f0, v0 = DeepPotB200(model, skin=2.0).eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    dd = DomainDeepPot(model, (1, 1, 1), skin=2.0)
    c = torch.as_tensor(coord, device="cuda:0")
    t = torch.as_tensor(atype, device="cuda:0")
    e, f, v, ex = dd.eval_device(c, t, box, atom_virial=True)
    assert dd.plan.nghost > len(atype)  # (rc + skin = 8 A shell of a 24.9 A box: more ghosts than atoms)
    assert abs(float(e) - e0[0, 0]) <= 1e-10 * abs(e0[0, 0])
    assert rel(f.cpu().numpy(), f0[0]) <= 1e-10
    assert rel(v.cpu().numpy(), v0[0]) <= 1e-10
    assert rel(ex["atom_virial"].reshape(-1, 9).sum(0).cpu().numpy(), v0[0]) <= 1e-9
    # second step: small displacement, the plan and the list are reused
    c2 = c + 0.01 * torch.randn_like(c)
    e2, f2, v2, _ = dd.eval_device(c2, t, box)
    e3, f3, v3 = DeepPotB200(model, skin=2.0).eval(c2.cpu().numpy().reshape(1, -1), box.reshape(1, 9), atype)
    assert dd.state.ago == 2
    assert rel(f2.cpu().numpy(), f3[0]) <= 1e-10


@pytest.mark.parametrize("form", ["dense", "pointers"])
def test_compute_with_a_device_neighbour_list(pkg, form):
    """DeepPotB200.compute_nlist consumes an MD code's own list, device tensors in and out (the
    counterpart of DeepPotPT::compute(..., InputNlist, ago)): forces on local AND ghost atoms, folded here with the
    ghost -> owner map as LAMMPS' reverse communication does, must equal the plain evaluation."""
    from deepmd_kit_b200 import ops
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    cfg = SeAConfig()
    coord, atype, box = g.water_box(2, 0.01)
    model = SeAModel(cfg, torch.float64, "cuda:0")
    e0, f0, v0 = DeepPotB200(model, skin=2.0).eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    dev = torch.device("cuda:0")
    c = torch.as_tensor(coord, device=dev)
    t = torch.as_tensor(atype, device=dev)
    nloc = len(atype)
    rc = cfg.rcut + 1.0  # the "MD code" keeps a 1 A skin
    ext_c, ext_t, mapping = ops.copy_coord(ops.normalize_coord(c.clone(), box), t, box, rc)
    numneigh, rows = ops.build_nlist(ext_c, nloc, rc, ext_t)
    dp = DeepPotB200(model)
    if form == "dense":
        e, f, v, ex = dp.compute_nlist(ext_c, ext_t, nloc, numneigh, rows=rows, ago=0)
    else:  # CSR rows behind device row pointers
        nn = numneigh.to(torch.int64)
        keep = torch.arange(rows.shape[1], device=dev)[None, :] < nn[:, None]
        neigh = rows[keep].contiguous()
        first = ops.csr_row_pointers(neigh, numneigh)
        e, f, v, ex = dp.compute_nlist(ext_c, ext_t, nloc, numneigh, firstneigh=first, ago=0)
        e, f, v, ex = dp.compute_nlist(ext_c, ext_t, nloc, numneigh, firstneigh=first, ago=1)  # cached partition
    assert f.shape == (ext_t.numel(), 3) and f.is_cuda
    folded = torch.zeros((nloc, 3), dtype=f.dtype, device=dev).index_add_(0, mapping.long(), f)
    assert abs(float(e) - e0[0, 0]) <= 1e-10 * abs(e0[0, 0])
    assert rel(folded.cpu().numpy(), f0[0]) <= 1e-10
    assert rel(v.cpu().numpy(), v0[0]) <= 1e-10
