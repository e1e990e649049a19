"""GPU parity tests: the CUDA path (through the C ABI / torch ops) against the CPU oracle and the
reference's golden vectors.  Integer outputs must be bit-exact; floating point within the
north_star tolerances: 1e-10 relative (fp64), 1e-5 (fp32), relative to the largest reference
magnitude of the compared array (sums with cancellation have no meaningful per-element ratio).
"""
import numpy as np
import pytest
import torch

from oracle import cpu as ocpu
from _systems import extended_system, golden, random_table, six_atom_system, water_like_box

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-10, np.float32: 1e-5}
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as g

    return g.load_package().ops


def T(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def close(got, want, dtype, scale=None, fac=1.0):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    assert got.shape == want.shape
    s = scale if scale is not None else max(np.abs(want).max() if want.size else 0.0, 1e-300)
    # fp64: the north_star tolerance itself, never loosened.  fp32: `fac` covers the different summation ORDER of
    # O(100)-term single-precision sums (the reference's own CPU and GPU paths differ by the same amount); DESIGN 4.
    tol = TOL[dtype] * (fac if dtype == np.float32 else 1.0)
    err = np.abs(got - want).max() if want.size else 0.0
    assert err <= tol * s, f"max abs err {err:.3e} > {tol:.1e} * {s:.3e}"


def gold(got, want):
    """Literal arrays of the reference lib tests carry ~8 digits; the lib tests use 1e-5
    (source/lib/tests/test_tabulate_se_a.cc:983-990)."""
    np.testing.assert_allclose(np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1),
                               rtol=1e-5, atol=1e-5)


def avg_std(ntypes, nnei, dtype, seed=3):
    rng = np.random.default_rng(seed)
    avg = rng.normal(scale=0.05, size=(ntypes, nnei * 4)).astype(dtype)
    avg[:, 1::4] = 0
    avg[:, 2::4] = 0
    avg[:, 3::4] = 0
    std = (0.08 + 0.1 * rng.random(size=(ntypes, nnei * 4))).astype(dtype)
    return avg, std


# ------------------------------------------------------------------ a5: format_nlist ---------
@pytest.mark.parametrize("cls", ["TestFormatNlist", "TestFormatNlistShortSel"])
def test_format_nlist_golden(ops, olib, cls):
    g = golden("fmt_nlist.json")[cls]
    s = six_atom_system(olib, rc=g["rc"])
    nl = ops.format_nlist(T(s["coord"]), T(s["atype"]), T(s["numneigh"]), T(s["rows"]), s["nloc"], len(s["atype"]),
                          g["rc"], g["sec_a"])
    assert N(nl).reshape(-1).tolist() == g["expect_nlist_cpy"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("jitter", [0.0, 0.05])
@pytest.mark.parametrize("sel", [(46, 92), (6, 11)])
def test_format_nlist_bit_exact(ops, olib, dtype, jitter, sel):
    """Replicated boxes (exact distance ties when jitter == 0) and overflowing selections."""
    coord, atype, box = water_like_box(ncopy=2, seed=1, jitter=jitter, dtype=dtype)
    s = extended_system(olib, coord, atype, box, 6.5, dtype=dtype)
    sec = [0, sel[0], sel[0] + sel[1]]
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    want, _ = olib.format_nlist(s["coord"], s["atype"], off, neigh, 6.0, sec)
    got = ops.format_nlist(T(s["coord"]), T(s["atype"]), T(s["numneigh"]), T(s["rows"]), s["nloc"], len(s["atype"]),
                           6.0, sec)
    assert np.array_equal(N(got), want)
    # CSR rows through device row pointers (InputNlist layout) + a shuffled raw order
    rng = np.random.default_rng(0)
    rows = s["rows"].copy()
    for i in range(rows.shape[0]):
        n = s["numneigh"][i]
        rows[i, :n] = rng.permutation(rows[i, :n])
    off2, neigh2 = ocpu.dense_to_csr(rows, s["numneigh"])
    neigh_t = T(neigh2)
    nn_t = T(s["numneigh"])
    fn = ops.csr_row_pointers(neigh_t, nn_t)
    got2 = ops.format_nlist(T(s["coord"]), T(s["atype"]), nn_t, None, s["nloc"], len(s["atype"]), 6.0, sec,
                            firstneigh=fn)
    assert np.array_equal(N(got2), want)


def test_format_nlist_virtual_atoms_and_empty(ops, olib):
    coord, atype, box = water_like_box(ncopy=1, seed=2, jitter=0.05)
    atype = atype.copy()
    atype[::7] = -1  # virtual atoms are never neighbours
    s = extended_system(olib, coord, atype, box, 6.0)
    sec = [0, 20, 60]
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    want, _ = olib.format_nlist(s["coord"], s["atype"], off, neigh, 6.0, sec)
    got = ops.format_nlist(T(s["coord"]), T(s["atype"]), T(s["numneigh"]), T(s["rows"]), s["nloc"], len(s["atype"]),
                           6.0, sec)
    assert np.array_equal(N(got), want)
    # rows with zero neighbours
    nn0 = np.zeros_like(s["numneigh"])
    got0 = ops.format_nlist(T(s["coord"]), T(s["atype"]), T(nn0), T(s["rows"]), s["nloc"], len(s["atype"]), 6.0, sec)
    assert (N(got0) == -1).all()
    # nloc == 0
    e = ops.format_nlist(T(s["coord"]), T(s["atype"]), T(nn0[:0]), T(s["rows"][:0]), 0, len(s["atype"]), 6.0, sec)
    assert e.shape == (0, 60)


# ------------------------------------------------------------------ a6/a7: env mat -----------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prod_env_mat_a_golden(ops, olib, dtype):
    g = golden("env_mat_a.json")["TestEnvMatA"]
    s = six_atom_system(olib, rc=g["rc"], dtype=dtype)
    sec = g["sec_a"]
    nnei = sec[-1]
    avg = np.zeros((2, nnei * 4), dtype)
    std = np.ones((2, nnei * 4), dtype)
    em, dv, rij, nl = ops.prod_env_mat_a(T(s["coord"]), T(s["atype"]), T(s["numneigh"]), T(s["rows"]), T(avg), T(std),
                                         s["nloc"], len(s["atype"]), g["rc"], g["rc_smth"], sec)
    np.testing.assert_allclose(N(em).reshape(-1), np.array(g["expected_env"]), atol=1e-5)
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    w_em, w_dv, w_rij, w_nl = olib.prod_env_mat_a(s["coord"], s["atype"], off, neigh, avg, std, s["nloc"], g["rc"],
                                                  g["rc_smth"], sec)
    assert np.array_equal(N(nl), w_nl)
    close(N(em), w_em, dtype)
    close(N(dv), w_dv, dtype)
    close(N(rij), w_rij, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("jitter", [0.0, 0.05])
def test_prod_env_mat_a_water(ops, olib, dtype, jitter):
    coord, atype, box = water_like_box(ncopy=2, seed=5, jitter=jitter, dtype=dtype)
    atype = atype.copy()
    atype[5] = -1  # one virtual centre atom: zero rows
    s = extended_system(olib, coord, atype, box, 6.8, dtype=dtype)
    sec = [0, 46, 138]
    avg, std = avg_std(2, 138, dtype)
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    w = olib.prod_env_mat_a(s["coord"], s["atype"], off, neigh, avg, std, s["nloc"], 6.0, 0.5, sec)
    got = ops.prod_env_mat_a(T(s["coord"]), T(s["atype"]), T(s["numneigh"]), T(s["rows"]), T(avg), T(std), s["nloc"],
                             len(s["atype"]), 6.0, 0.5, sec)
    assert np.array_equal(N(got[3]), w[3])
    for a, b in zip(got[:3], w[:3]):
        close(N(a), b, dtype)


def test_prod_env_mat_a_ilist_permutation(ops, olib):
    """ilist in arbitrary order writes row ilist[r] (prod_env_mat.cc:43-46)."""
    coord, atype, box = water_like_box(ncopy=1, seed=6, jitter=0.05)
    s = extended_system(olib, coord, atype, box, 6.0)
    sec = [0, 20, 50]
    avg, std = avg_std(2, 50, np.float64)
    rng = np.random.default_rng(1)
    perm = rng.permutation(s["nloc"]).astype(np.int32)
    off, neigh = ocpu.dense_to_csr(s["rows"][perm], s["numneigh"][perm])
    w = olib.prod_env_mat_a(s["coord"], s["atype"], off, neigh, avg, std, s["nloc"], 6.0, 1.0, sec, ilist=perm)
    got = ops.prod_env_mat_a(T(s["coord"]), T(s["atype"]), T(s["numneigh"][perm]), T(s["rows"][perm]), T(avg), T(std),
                             s["nloc"], len(s["atype"]), 6.0, 1.0, sec, ilist=T(perm))
    assert np.array_equal(N(got[3]), w[3])
    close(N(got[0]), w[0], np.float64)
    close(N(got[1]), w[1], np.float64)


# ------------------------------------------------------------------ a8-a10: tabulate ---------
def _tab_golden(dtype):
    g = golden("tabulate_se_a.json")["TestTabulateSeA"]
    nloc, nnei, M = g["nloc"], g["nnei"], g["last_layer_size"]
    return (g, np.array(g["table"], dtype).reshape(-1, M * 6), np.array(g["info"], dtype),
            np.array(g["em_x"], dtype).reshape(nloc * nnei, 1), np.array(g["em"], dtype).reshape(nloc, nnei, 4), M)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tabulate_golden(ops, olib, dtype):
    """source/lib/tests/test_tabulate_se_a.cc + source/tests/pt/test_tabulate_fusion_se_a.py literals."""
    g, table, info, em_x, em, M = _tab_golden(dtype)
    nloc, nnei = em.shape[:2]
    out = ops.tabulate_fusion_se_a(T(table), torch.as_tensor(info), T(em_x), T(em), M)
    close(N(out), olib.tabulate_fusion_se_a(table, info, em_x, em, M), dtype, fac=4)
    gold(N(out), g["expected_xyz_scatter"])
    dy = np.ones((nloc, 4, M), dtype)
    gx, gem, _ = ops.tabulate_fusion_se_a_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dy), M)
    gold(N(gx), g["expected_dy_dem_x"])
    gold(N(gem), g["expected_dy_dem"])
    two = np.array(g["two_embed"], dtype).reshape(nloc * nnei, M)
    out2 = ops.tabulate_fusion_se_a(T(table), torch.as_tensor(info), T(em_x), T(em), M, two_embed=T(two))
    gold(N(out2), g["expected_xyz_scatter_with_two_embed"])
    gx2, gem2, gtwo = ops.tabulate_fusion_se_a_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dy), M,
                                                    two_embed=T(two))
    gold(N(gx2), g["expected_dy_dem_x_with_two_embed"])
    gold(N(gem2), g["expected_dy_dem_with_two_embed"])
    assert gtwo.shape == two.shape


def _random_tab_case(rng, dtype, nloc, nnei, M, nreal_max=None, unsorted=False):
    """Env-mat-like inputs: descending em_x, trailing padding (em_x = pad value, angular part 0),
    values reaching below `lower` and above `max` (extrapolation branches)."""
    info = np.array([-0.4, 2.0, 6.0, 0.05, 0.5, -1.0], dtype)
    nspline = int((info[1] - info[0]) / info[3]) + int((info[2] - info[1]) / info[4]) + 1
    table = random_table(nspline, M, rng, dtype)
    em = rng.normal(size=(nloc, nnei, 4)).astype(dtype)
    em_x = np.sort(rng.uniform(-0.8, 7.5, size=(nloc, nnei)), axis=1)[:, ::-1].astype(dtype)
    pad = dtype(-0.37)
    for i in range(nloc):
        nreal = rng.integers(0, (nreal_max or nnei) + 1)
        em_x[i, nreal:] = pad
        em[i, nreal:, 0] = pad
        em[i, nreal:, 1:] = 0
    if unsorted:
        for i in range(nloc):
            p = rng.permutation(nnei)
            em_x[i] = em_x[i, p]
            em[i] = em[i, p]
    em[:, :, 0] = em_x
    return table, info, np.ascontiguousarray(em_x.reshape(-1, 1)), np.ascontiguousarray(em)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nnei,M", [(46, 100), (92, 100), (7, 8), (33, 32), (64, 40), (20, 160), (138, 128), (50, 80), (41, 64), (30, 97), (29, 26),
                                    (9, 104)])
@pytest.mark.parametrize("is_sorted", [True, False])
def test_tabulate_vs_oracle(ops, olib, dtype, nnei, M, is_sorted):
    rng = np.random.default_rng(nnei * 1000 + M)
    nloc = 37
    table, info, em_x, em = _random_tab_case(rng, dtype, nloc, nnei, M, unsorted=not is_sorted)
    want = olib.tabulate_fusion_se_a(table, info, em_x, em, M, is_sorted=is_sorted)
    got = ops.tabulate_fusion_se_a(T(table), torch.as_tensor(info), T(em_x), T(em), M, is_sorted=is_sorted)
    close(N(got), want, dtype, fac=4)
    dy = rng.normal(size=(nloc, 4, M)).astype(dtype)
    wx, wem, _ = olib.tabulate_fusion_se_a_grad(table, info, em_x, em, dy, M, is_sorted=is_sorted)
    gx, gem, _ = ops.tabulate_fusion_se_a_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dy), M,
                                               is_sorted=is_sorted)
    close(N(gx), wx, dtype, fac=4)
    close(N(gem), wem, dtype, fac=4)
    dzx = rng.normal(size=em_x.shape).astype(dtype)
    dzem = rng.normal(size=em.shape).astype(dtype)
    wgg = olib.tabulate_fusion_se_a_grad_grad(table, info, em_x, em, dzx, dzem, M, is_sorted=is_sorted)
    ggg = ops.tabulate_fusion_se_a_grad_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dzx), T(dzem), M,
                                             is_sorted=is_sorted)
    close(N(ggg), wgg, dtype, fac=4)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("is_sorted", [True, False])
def test_tabulate_atten_vs_oracle(ops, olib, dtype, is_sorted):
    rng = np.random.default_rng(11)
    nloc, nnei, M = 19, 120, 100
    table, info, em_x, em = _random_tab_case(rng, dtype, nloc, nnei, M, unsorted=not is_sorted)
    two = rng.normal(size=(nloc * nnei, M)).astype(dtype)
    want = olib.tabulate_fusion_se_a(table, info, em_x, em, M, two_embed=two, is_sorted=is_sorted)
    got = ops.tabulate_fusion_se_a(T(table), torch.as_tensor(info), T(em_x), T(em), M, two_embed=T(two),
                                   is_sorted=is_sorted)
    close(N(got), want, dtype, fac=4)
    dy = rng.normal(size=(nloc, 4, M)).astype(dtype)
    wx, wem, wtwo = olib.tabulate_fusion_se_a_grad(table, info, em_x, em, dy, M, two_embed=two, is_sorted=is_sorted)
    gx, gem, gtwo = ops.tabulate_fusion_se_a_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dy), M,
                                                  two_embed=T(two), is_sorted=is_sorted)
    close(N(gx), wx, dtype, fac=4)
    close(N(gem), wem, dtype, fac=4)
    close(N(gtwo), wtwo.reshape(nloc * nnei, M), dtype, fac=4)
    dzx = rng.normal(size=em_x.shape).astype(dtype)
    dzem = rng.normal(size=em.shape).astype(dtype)
    dztwo = rng.normal(size=two.shape).astype(dtype)
    wgg = olib.tabulate_fusion_se_a_grad_grad(table, info, em_x, em, dzx, dzem, M, two_embed=two, dz_dtwo=dztwo,
                                              is_sorted=is_sorted)
    ggg = ops.tabulate_fusion_se_a_grad_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dzx), T(dzem), M,
                                             two_embed=T(two), dz_dy_dtwo=T(dztwo), is_sorted=is_sorted)
    close(N(ggg), wgg, dtype, fac=4)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nd", [9, 16, 25])
@pytest.mark.parametrize("two", [False, True])
@pytest.mark.parametrize("is_sorted", [True, False])
def test_tabulate_higher_basis_vs_oracle(ops, olib, dtype, nd, two, is_sorted):
    """ndescrpt = 9 / 16 / 25 (tabulate.cc:456-560; csrc/tabulate_nd.cu): forward, backward and second order against
    the CPU checkers, with two_embed, padding folds (the fold test reads components 1..3 whatever the basis is), both
    extrapolation branches, a channel count that is not a multiple of 32."""
    rng = np.random.default_rng(100 * nd + 7)
    nloc, nnei, M = 23, 37, 45
    table, info, em_x, em4 = _random_tab_case(rng, dtype, nloc, nnei, M, unsorted=not is_sorted)
    em = rng.normal(size=(nloc, nnei, nd)).astype(dtype)
    em[:, :, :4] = em4
    pad = em4[:, :, 1:].reshape(nloc, nnei, 3).any(axis=2) == 0  # padded slots keep a zero angular part
    em[pad, 4:] = 0
    te = rng.normal(size=(nloc * nnei, M)).astype(dtype) if two else None
    tt = None if te is None else T(te)
    want = olib.tabulate_fusion_se_a(table, info, em_x, em, M, two_embed=te, is_sorted=is_sorted)
    got = ops.tabulate_fusion_se_a(T(table), torch.as_tensor(info), T(em_x), T(em), M, two_embed=tt, is_sorted=is_sorted)
    assert got.shape == (nloc, nd, M)
    close(N(got), want, dtype, fac=4)
    dy = rng.normal(size=(nloc, nd, M)).astype(dtype)
    wx, wem, wtwo = olib.tabulate_fusion_se_a_grad(table, info, em_x, em, dy, M, two_embed=te, is_sorted=is_sorted)
    gx, gem, gtwo = ops.tabulate_fusion_se_a_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dy), M,
                                                  two_embed=tt, is_sorted=is_sorted)
    close(N(gx), wx, dtype, fac=4)
    close(N(gem), wem, dtype, fac=4)
    if two:
        close(N(gtwo), wtwo.reshape(nloc * nnei, M), dtype, fac=4)
    dzx = rng.normal(size=em_x.shape).astype(dtype)
    dzem = rng.normal(size=em.shape).astype(dtype)
    dzt = rng.normal(size=te.shape).astype(dtype) if two else None
    wgg = olib.tabulate_fusion_se_a_grad_grad(table, info, em_x, em, dzx, dzem, M, two_embed=te, dz_dtwo=dzt,
                                              is_sorted=is_sorted)
    ggg = ops.tabulate_fusion_se_a_grad_grad(T(table), torch.as_tensor(info), T(em_x), T(em), T(dzx), T(dzem), M,
                                             two_embed=tt, dz_dy_dtwo=None if dzt is None else T(dzt),
                                             is_sorted=is_sorted)
    close(N(ggg), wgg, dtype, fac=4)


def test_tabulate_rejects_other_basis_dimensions(ops):
    table = torch.zeros(4, 48, dtype=torch.float64, device=DEV)
    info = torch.tensor([0, 0.2, 0.4, 0.01, 0.1, -1], dtype=torch.float64)
    with pytest.raises(ValueError):  # tabulate.h is_supported_se_a_basis_dimension: 4, 9, 16, 25
        ops.tabulate_fusion_se_a(table, info, torch.zeros(6, 1, dtype=torch.float64, device=DEV),
                                 torch.zeros(2, 3, 5, dtype=torch.float64, device=DEV), 8)


def test_tabulate_empty_neighbors(ops):
    """test_tabulate_se_a.cc:761-786: nnei == 0 gives a zero descriptor and empty gradients."""
    table = torch.zeros(4, 48, dtype=torch.float64, device=DEV)
    info = torch.tensor([0, 0.2, 0.4, 0.01, 0.1, -1], dtype=torch.float64)
    em_x = torch.zeros(0, 1, dtype=torch.float64, device=DEV)
    em = torch.zeros(3, 0, 4, dtype=torch.float64, device=DEV)
    out = ops.tabulate_fusion_se_a(table, info, em_x, em, 8)
    assert out.shape == (3, 4, 8) and float(out.abs().sum()) == 0.0
    gx, gem, _ = ops.tabulate_fusion_se_a_grad(table, info, em_x, em, torch.ones(3, 4, 8, dtype=torch.float64,
                                                                               device=DEV), 8)
    assert gx.numel() == 0 and gem.numel() == 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tabulate_sections(ops, olib, dtype):
    """Strided per-type sections over the full env-mat == per-section calls on copies, summed."""
    rng = np.random.default_rng(21)
    nloc, M = 23, 100
    sec = [0, 46, 138]
    ems, tabs, infos = [], [], []
    want = 0
    dy = rng.normal(size=(nloc, 4, M)).astype(dtype)
    wg = []
    for t in range(2):
        table, info, em_x, em = _random_tab_case(rng, dtype, nloc, sec[t + 1] - sec[t], M)
        ems.append(em)
        tabs.append(table)
        infos.append(info)
        want = want + olib.tabulate_fusion_se_a(table, info, em_x, em, M)
        gx, gem, _ = olib.tabulate_fusion_se_a_grad(table, info, em_x, em, dy, M)
        gem = gem.copy()
        gem[:, :, 0] += gx
        wg.append(gem)
    em_full = np.concatenate(ems, axis=1).reshape(nloc, -1)
    got = ops.tabulate_sections_fwd([T(t) for t in tabs], [torch.as_tensor(i) for i in infos], T(em_full), sec, M)
    close(N(got), want, dtype, fac=4)
    gg = ops.tabulate_sections_grad([T(t) for t in tabs], [torch.as_tensor(i) for i in infos], T(em_full), T(dy), sec,
                                    M)
    close(N(gg), np.concatenate(wg, axis=1), dtype, fac=4)


def test_torch_ops_autograd(ops, olib):
    """torch.ops.deepmd.tabulate_fusion_se_a/_se_atten: forward, backward, double backward
    (source/tests/pt/test_tabulate_fusion_se_a.py:1424-1511)."""
    dtype = np.float64
    g, table, info, em_x, em, M = _tab_golden(dtype)
    nloc, nnei = em.shape[:2]
    tt, ti = T(table), torch.as_tensor(info)
    ex = T(em_x).requires_grad_(True)
    ee = T(em).requires_grad_(True)
    out = torch.ops.deepmd.tabulate_fusion_se_a(tt, ti, ex, ee, M)[0]
    gold(N(out), g["expected_xyz_scatter"])
    gx, gem = torch.autograd.grad(out, [ex, ee], torch.ones_like(out), create_graph=True)
    gold(N(gx), g["expected_dy_dem_x"])
    gold(N(gem), g["expected_dy_dem"])
    # double backward == oracle grad_grad contracted with the same cotangents
    rng = np.random.default_rng(0)
    cx = rng.normal(size=em_x.shape)
    cem = rng.normal(size=em.shape)
    dy0 = torch.ones_like(out).requires_grad_(True)
    out = torch.ops.deepmd.tabulate_fusion_se_a(tt, ti, ex, ee, M)[0]
    gx, gem = torch.autograd.grad(out, [ex, ee], dy0, create_graph=True)
    (gdy,) = torch.autograd.grad([gx, gem], [dy0], [T(cx), T(cem)])
    want = olib.tabulate_fusion_se_a_grad_grad(table, info, em_x, em, cx, cem, M)
    close(N(gdy), want, dtype, fac=10)
    # se_atten schema
    two = np.array(g["two_embed"], dtype).reshape(nloc * nnei, M)
    tw = T(two).requires_grad_(True)
    out2 = torch.ops.deepmd.tabulate_fusion_se_atten(tt, ti, ex, ee, tw, M, True)[0]
    gold(N(out2), g["expected_xyz_scatter_with_two_embed"])
    g2 = torch.autograd.grad(out2, [ex, ee, tw], torch.ones_like(out2))
    gold(N(g2[0]), g["expected_dy_dem_x_with_two_embed"])
    gold(N(g2[1]), g["expected_dy_dem_with_two_embed"])
    # device validation (source/tests/pt/test_tabulate_device_validation.py)
    with pytest.raises(RuntimeError):
        torch.ops.deepmd.tabulate_fusion_se_a(tt, ti.to(DEV), ex, ee, M)
    with pytest.raises(RuntimeError):
        torch.ops.deepmd.tabulate_fusion_se_a(tt, ti, ex.cpu(), ee, M)


# ------------------------------------------------------------------ a11/a12: force, virial ----
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prod_force_virial_golden(ops, olib, dtype):
    gf = golden("prod_force_a.json")["TestProdForceA"]
    gv = golden("prod_virial_a.json")["TestProdVirialA"]
    s = six_atom_system(olib, rc=gf["rc"], dtype=dtype)
    sec = gf["sec_a"]
    nnei = sec[-1]
    nloc, nall = s["nloc"], len(s["atype"])
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    nl, _ = olib.format_nlist(s["coord"], s["atype"], off, neigh, gf["rc"], sec)
    em, dv, rij = olib.env_mat_a(s["coord"], s["atype"], nl, gf["rc_smth"], gf["rc"], sec)
    # same net_deriv as the reference fixtures (test_prod_force_a.cc / test_prod_virial_a.cc SetUp)
    nd = (10 - 0.01 * np.arange(nloc * nnei * 4, dtype=np.float64)).astype(dtype).reshape(nloc, nnei * 4)
    f1 = ops.prod_force_a(T(nd), T(dv), T(nl), nloc, nall, nnei)
    np.testing.assert_allclose(N(f1).reshape(-1), np.array(gf["expected_force"])[: nall * 3], atol=2e-4)
    wf = olib.prod_force_a(nd, dv, nl, nall)
    close(N(f1), wf, dtype, fac=4)
    # two identical frames in one call (nframes axis of prod_force_a_gpu)
    nf = 2
    f = ops.prod_force_a(T(np.tile(nd, (nf, 1))), T(np.tile(dv, (nf, 1))), T(np.tile(nl, (nf, 1))), nloc, nall, nnei,
                         nframes=nf)
    close(N(f)[0], wf, dtype, fac=4)
    close(N(f)[1], wf, dtype, fac=4)
    v, av = ops.prod_virial_a(T(nd), T(dv), T(rij), T(nl), nloc, nall, nnei)
    np.testing.assert_allclose(N(v), np.array(gv["expected_virial"]), rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(N(av), np.array(gv["expected_atom_virial"]), rtol=1e-5, atol=2e-4)
    wv, wav = olib.prod_virial_a(nd, dv, rij, nl, nall)
    close(N(v), wv, dtype, fac=4)
    close(N(av), wav, dtype, fac=4)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prod_force_virial_water(ops, olib, dtype):
    coord, atype, box = water_like_box(ncopy=2, seed=8, jitter=0.05, dtype=dtype)
    s = extended_system(olib, coord, atype, box, 6.0, dtype=dtype)
    sec = [0, 46, 138]
    nnei = 138
    nloc, nall = s["nloc"], len(s["atype"])
    avg, std = avg_std(2, nnei, dtype)
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    em, dv, rij, nl = olib.prod_env_mat_a(s["coord"], s["atype"], off, neigh, avg, std, nloc, 6.0, 0.5, sec)
    rng = np.random.default_rng(4)
    nd = rng.normal(size=(nloc, nnei * 4)).astype(dtype)
    wf = olib.prod_force_a(nd, dv, nl, nall)
    wv, wav = olib.prod_virial_a(nd, dv, rij, nl, nall)
    f = ops.prod_force_a(T(nd), T(dv), T(nl), nloc, nall, nnei)
    close(N(f), wf, dtype, fac=4)
    v, av = ops.prod_virial_a(T(nd), T(dv), T(rij), T(nl), nloc, nall, nnei)
    close(N(v), wv, dtype, fac=20)
    close(N(av), wav, dtype, fac=4)
    f2, v2, av2 = ops.prod_force_virial_a(T(nd), T(dv), T(rij), T(nl), nloc, nall, nnei, atom_virial=True)
    close(N(f2), wf, dtype, fac=4)
    close(N(v2), wv, dtype, fac=20)
    close(N(av2), wav, dtype, fac=4)
    f3, v3, av3 = ops.prod_force_virial_a(T(nd), T(dv), T(rij), T(nl), nloc, nall, nnei)
    assert av3 is None
    close(N(v3), wv, dtype, fac=20)
    # size-independent property: translation invariance => net force vanishes after folding ghosts
    fold = np.zeros((nloc, 3))
    np.add.at(fold, s["mapping"], N(f).reshape(-1, 3).astype(np.float64))
    assert np.abs(fold.sum(0)).max() <= 1e3 * TOL[dtype] * np.abs(wf).max()
    # torch op schemas (TF-style)
    nat = torch.tensor([nloc, nall, 0, 0], dtype=torch.int32)
    fo = torch.ops.deepmd.prod_force_se_a(T(nd).reshape(1, -1), T(dv).reshape(1, -1), T(nl).reshape(1, -1), nat, nnei, 0)
    close(N(fo), wf, dtype, fac=4)
    vo, avo = torch.ops.deepmd.prod_virial_se_a(T(nd).reshape(1, -1), T(dv).reshape(1, -1), T(rij).reshape(1, -1),
                                                T(nl).reshape(1, -1), nat, nnei, 0)
    close(N(vo), wv, dtype, fac=20)
    close(N(avo), wav, dtype, fac=4)


# ------------------------------------------------------------------ a2-a4: list front end ----
def _ghost_key(c, t, m, nloc):
    c = np.asarray(c, np.float64)
    rows = [(int(m[i]), int(t[i])) + tuple(np.round(c[i], 6)) for i in range(nloc, len(t))]
    return sorted(rows)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_normalize_copy_build(ops, olib, dtype):
    coord, atype, box = water_like_box(ncopy=2, seed=9, jitter=0.3, dtype=dtype)
    coord = coord + dtype(7.3)  # push atoms outside the cell
    w = olib.normalize_coord(coord, box)
    gc = ops.normalize_coord(T(coord).clone(), box)
    assert np.array_equal(N(gc), w)  # same operation order, no contraction: bitwise
    for rc in (6.0, 4.0, 10.5):
        wc, wt, wm = olib.copy_coord(w, atype, box, rc)
        c, t, m = ops.copy_coord(T(w), T(atype), box, rc)
        nloc = len(atype)
        assert len(wt) == t.numel()
        assert np.array_equal(N(c)[:nloc], w) and np.array_equal(N(m)[:nloc], np.arange(nloc))
        assert _ghost_key(N(c), N(t), N(m), nloc) == _ghost_key(wc, wt, wm, nloc)
    # raw list on the oracle's extended system: rows identical element by element
    ext_c, ext_t, ext_m = olib.copy_coord(w, atype, box, 6.0)
    nloc = len(atype)
    wn, wr = olib.build_nlist(ext_c, nloc, 6.0)
    nn, rows = ops.build_nlist(T(ext_c), nloc, 6.0)
    assert np.array_equal(N(nn), wn)
    rows = N(rows)
    for i in range(nloc):
        assert np.array_equal(rows[i, : wn[i]], wr[i, : wn[i]])
    # virtual atoms + capacity error path
    ty = ext_t.copy()
    ty[::5] = -1
    wn2, wr2 = olib.build_nlist(ext_c, nloc, 6.0, atype=ty)
    nn2, rows2 = ops.build_nlist(T(ext_c), nloc, 6.0, atype=T(ty))
    assert np.array_equal(N(nn2), wn2)
    with pytest.raises(MemoryError):
        ops.build_nlist(T(ext_c), nloc, 6.0, mem_size=8)


def test_copy_coord_golden(ops):
    for cls in ("TestCopyCoord", "TestCopyCoordMoreCell"):
        g = golden("coord.json")[cls]
        posi = np.array(g["posi"]).reshape(-1, 3)
        atype = np.array(g["atype"], np.int32)
        box = np.array(g["boxt"]).reshape(3, 3)
        c, t, m = ops.copy_coord(T(posi), T(atype), box, g["rc"])
        want_c = np.array(g["_expected_posi_cpy"]).reshape(-1, 3)
        want_t = np.array(g["_expected_atype_cpy"])
        want_m = np.array(g["_expected_mapping"])
        nloc = len(atype)
        assert t.numel() == len(want_t)
        assert _ghost_key(N(c), N(t), N(m), nloc) == _ghost_key(want_c, want_t, want_m, nloc)
        with pytest.raises(MemoryError):
            ops.copy_coord(T(posi), T(atype), box, g["rc"], mem_nall=40)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_op_prod_env_mat_a_modes(ops, olib, dtype):
    """torch.ops.deepmd.prod_env_mat_a with ALL mesh encodings: PBC self-built list (len 6), pointer-free
    inline list (len > 16), raw host pointers (len 16, the LAMMPS hand-off) and no PBC (len 0), against the oracle."""
    coord, atype, box = water_like_box(ncopy=2, seed=12, jitter=0.1, dtype=dtype)
    nloc = len(atype)
    sec = [0, 46, 138]
    avg, std = avg_std(2, 138, dtype)
    cn = olib.normalize_coord(coord, box)
    ext_c, ext_t, mapping = olib.copy_coord(cn, atype, box, 6.0)
    nn, rows = olib.build_nlist(ext_c, nloc, 6.0)
    off, neigh = ocpu.dense_to_csr(rows, nn)
    w = olib.prod_env_mat_a(ext_c, ext_t, off, neigh, avg, std, nloc, 6.0, 0.5, sec)
    nat = torch.tensor([nloc, nloc, (atype == 0).sum(), (atype == 1).sum()], dtype=torch.int32)
    mesh6 = torch.zeros(6, dtype=torch.int32)
    em, dv, rij, nl = torch.ops.deepmd.prod_env_mat_a(T(coord).reshape(1, -1), T(atype).reshape(1, -1), nat,
                                                      T(box).reshape(1, 9), mesh6, T(avg), T(std), 0.0, 6.0, 0.5,
                                                      [46, 92], [0, 0])
    # ghost numbering differs from the oracle's copy_coord; compare per slot after mapping to owners
    want_nl = np.where(w[3] >= 0, mapping[np.maximum(w[3], 0)], -1)
    got_nl = N(nl).reshape(nloc, -1)
    # equal distances between different images of one owner can swap slots: compare em instead of
    # requiring identical owner ids slot by slot, then check the owner multiset per type section
    close(N(em), w[0], dtype)
    for i in range(nloc):
        for a, b in ((0, 46), (46, 138)):
            assert sorted(got_nl[i, a:b].tolist()) == sorted(want_nl[i, a:b].tolist())
    # inline list: extended system and raw rows handed over explicitly
    nall = len(ext_t)
    nat2 = torch.tensor([nloc, nall, 0, 0], dtype=torch.int32)
    head = np.zeros(16, np.int32)
    head[1] = nloc
    mesh = np.concatenate([head, np.arange(nloc, dtype=np.int32), nn.astype(np.int32), neigh[: off[-1]]])
    em2, dv2, rij2, nl2 = torch.ops.deepmd.prod_env_mat_a(T(ext_c).reshape(1, -1), T(ext_t).reshape(1, -1), nat2,
                                                          T(box).reshape(1, 9), T(mesh), T(avg), T(std), 0.0, 6.0,
                                                          0.5, [46, 92], [0, 0])
    assert np.array_equal(N(nl2).reshape(nloc, -1), w[3])
    close(N(em2), w[0], dtype)
    close(N(dv2), w[1], dtype)
    close(N(rij2), w[2], dtype)
    # length 16: the LAMMPS hand-off, raw HOST pointers to ilist / numneigh / firstneigh packed as int32 pairs
    # (source/op/tf/prod_env_mat_multi_device.cc:529-552: mesh[1] = inum, mesh[4..5], [8..9], [12..13] = pointers)
    import ctypes as C

    ilist_h = np.arange(nloc, dtype=np.int32)
    nn_h = np.ascontiguousarray(nn, dtype=np.int32)
    rows_h = [np.ascontiguousarray(neigh[off[i]:off[i + 1]], dtype=np.int32) for i in range(nloc)]
    first_h = (C.POINTER(C.c_int) * nloc)(*[r.ctypes.data_as(C.POINTER(C.c_int)) for r in rows_h])
    mesh16 = np.zeros(16, np.int32)
    mesh16[1] = nloc
    for k, ptr in ((4, ilist_h.ctypes.data), (8, nn_h.ctypes.data), (12, C.addressof(first_h))):
        mesh16[k:k + 2] = np.frombuffer(np.uint64(ptr).tobytes(), dtype=np.int32)
    em3, dv3, rij3, nl3 = torch.ops.deepmd.prod_env_mat_a(T(ext_c).reshape(1, -1), T(ext_t).reshape(1, -1), nat2,
                                                          T(box).reshape(1, 9), torch.from_numpy(mesh16), T(avg), T(std),
                                                          0.0, 6.0, 0.5, [46, 92], [0, 0])
    assert np.array_equal(N(nl3).reshape(nloc, -1), w[3])
    assert torch.equal(em3, em2) and torch.equal(dv3, dv2) and torch.equal(rij3, rij2)
    # length 0: no PBC, the op builds the list over the atoms it is given (an isolated cluster)
    wn, wrows = olib.build_nlist(coord, nloc, 6.0)
    woff, wneigh = ocpu.dense_to_csr(wrows, wn)
    w0 = olib.prod_env_mat_a(coord, atype, woff, wneigh, avg, std, nloc, 6.0, 0.5, sec)
    em0, dv0, rij0, nl0 = torch.ops.deepmd.prod_env_mat_a(T(coord).reshape(1, -1), T(atype).reshape(1, -1), nat,
                                                          T(box).reshape(1, 9), torch.zeros(0, dtype=torch.int32), T(avg),
                                                          T(std), 0.0, 6.0, 0.5, [46, 92], [0, 0])
    assert np.array_equal(N(nl0).reshape(nloc, -1), w0[3])
    close(N(em0), w0[0], dtype)
    close(N(dv0), w0[1], dtype)
    close(N(rij0), w0[2], dtype)


# ------------------------------------------------------------------ descriptor contraction ----
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("M,axis", [(100, 16), (8, 4), (33, 33), (16, 16), (97, 16), (128, 16), (64, 8), (50, 12)])
def test_descriptor_contraction(ops, dtype, M, axis):
    """dpb200_se_a_descriptor / _grad against the plain torch statement of
    deepmd/pt/model/descriptor/se_a.py:843-850 (and autograd for the backward)."""
    rng = np.random.default_rng(M)
    n, nnei = 57, 138
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    x = torch.as_tensor(rng.normal(size=(n, 4, M)).astype(dtype)).to(DEV).requires_grad_(True)
    xs = x / nnei
    want = torch.matmul(xs.permute(0, 2, 1), xs[:, :, :axis]).reshape(n, -1)
    got = ops.se_a_descriptor(x.detach(), axis, 1.0 / nnei)
    close(N(got), N(want), dtype, fac=4)
    cot = torch.as_tensor(rng.normal(size=(n, M * axis)).astype(dtype)).to(DEV)
    (wg,) = torch.autograd.grad(want, x, cot, retain_graph=True)
    gg = ops.se_a_descriptor_grad(cot, x.detach(), axis, 1.0 / nnei)
    close(N(gg), N(wg), dtype, fac=8)
    assert got.dtype == tdt
    # gather / scatter through a row list (the per-type selection of the fitting net)
    rows = torch.as_tensor(rng.permutation(n)[: n // 2].astype(np.int32)).to(DEV)
    got_r = ops.se_a_descriptor(x.detach(), axis, 1.0 / nnei, rows=rows)
    close(N(got_r), N(want)[N(rows)], dtype, fac=4)
    out = torch.zeros_like(x.detach())
    ops.se_a_descriptor_grad(cot[: n // 2].contiguous(), x.detach(), axis, 1.0 / nnei, rows=rows, out=out)
    (wg_r,) = torch.autograd.grad(want[rows.long()], x, cot[: n // 2])
    close(N(out), N(wg_r), dtype, fac=8)


def test_fitting_forward_backward_matches_autograd():
    import __graft_entry__ as g

    g.load_package()
    from deepmd_kit_b200.model import FittingNet

    for dims, dt, tol in (((1600, (240, 240, 240)), torch.float64, 1e-12), ((64, (32, 64, 64)), torch.float32, 2e-5)):
        f = FittingNet(dims[0], dims[1], True, 3, dt, DEV)
        x = torch.randn(301, dims[0], dtype=dt, device=DEV, requires_grad=True)
        e = f(x)
        (ga,) = torch.autograd.grad(e.sum(), x)
        e2, g2 = f.forward_backward(x.detach())
        assert float((e.detach() - e2).abs().max()) <= tol * float(e.detach().abs().max())
        assert float((ga - g2).abs().max()) <= tol * float(ga.abs().max())


# ------------------------------------------------------------------------------------------------
# fused descriptor epilogue + split-operand fitting GEMMs (csrc/fitting.cu, tabulate.cu desc_epilogue)
# ------------------------------------------------------------------------------------------------
def _two_section_case(dtype, nloc=97, sel=(46, 92), M=100, seed=11):
    rng = np.random.default_rng(seed)
    nnei = sum(sel)
    tables, infos = [], []
    for t in range(len(sel)):
        info = np.array([-0.4, 2.0, 6.0, 0.05, 0.5, -1.0], dtype)
        nspline = int((info[1] - info[0]) / info[3]) + int((info[2] - info[1]) / info[4]) + 1
        tables.append(T(random_table(nspline, M, rng, dtype)))
        infos.append(torch.as_tensor(info))
    em = rng.normal(scale=0.3, size=(nloc, nnei, 4)).astype(dtype)
    em[:, :, 0] = np.sort(rng.uniform(-0.2, 2.5, size=(nloc, nnei)), axis=1)[:, ::-1]
    # trailing padding per section, like a real env-mat
    for i in range(nloc):
        for t, (a, b) in enumerate(zip(np.cumsum((0,) + sel[:-1]), np.cumsum(sel))):
            k = int(rng.integers(0, (b - a) // 2))
            if k:
                em[i, b - k:b, 0] = -0.37
                em[i, b - k:b, 1:] = 0
    sec = [0] + list(np.cumsum(sel))
    return tables, infos, T(em.reshape(nloc, -1)), sec, M


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tabulate_desc_epilogue(ops, dtype):
    tables, infos, em, sec, M = _two_section_case(dtype)
    axis, nnei = 16, sec[-1]
    nloc = em.shape[0]
    want_out = ops.tabulate_sections_fwd(tables, infos, em, sec, M)
    want_d = ops.se_a_descriptor(want_out, axis, 1.0 / nnei)
    perm = torch.randperm(nloc, device=DEV).to(torch.int32)
    # mode 1: plain D, permuted rows
    out, d1, _ = ops.tabulate_sections_desc(tables, infos, em, sec, M, axis, 1.0 / nnei, desc_row=perm, mode=1)
    assert torch.equal(out, want_out)
    got = torch.empty_like(want_d)
    got[:] = d1[perm.long()]
    close(N(got), N(want_d), dtype)
    # mode 2: split operand
    out, d2, ex = ops.tabulate_sections_desc(tables, infos, em, sec, M, axis, 1.0 / nnei, desc_row=perm, mode=2,
                                             nslice=7, pad_rows=32)
    assert torch.equal(out, want_out)
    assert d2.shape[0] == nloc + 32 and int(d2[nloc:].abs().sum()) == 0
    K = M * axis
    d2 = d2[perm.long()]
    if dtype == np.float64:
        sl = d2.reshape(nloc, 7, K).to(torch.float64)
        assert int(sl.abs().max()) <= 128
        w = torch.tensor([2.0 ** (-7 - 8 * s) for s in range(7)], dtype=torch.float64, device=DEV)
        rec = (sl * w[None, :, None]).sum(1) * torch.ldexp(torch.ones((), dtype=torch.float64, device=DEV),
                                                           ex[perm.long()].to(torch.int32))[:, None]
        rowmax = want_d.abs().amax(1, keepdim=True)
        err = ((rec - want_d).abs() / rowmax).max().item()
        assert err < 2.0 ** -50, err
        # the production slice count takes a specialised (compile-time) path: same contract at 47 fraction bits
        out, d6, ex6 = ops.tabulate_sections_desc(tables, infos, em, sec, M, axis, 1.0 / nnei, desc_row=perm, mode=2,
                                                  nslice=6, pad_rows=32)
        assert torch.equal(out, want_out) and torch.equal(ex6, ex)
        sl = d6[perm.long()].reshape(nloc, 6, K).to(torch.float64)
        assert int(sl.abs().max()) <= 128
        rec = (sl * w[None, :6, None]).sum(1) * torch.ldexp(torch.ones((), dtype=torch.float64, device=DEV),
                                                            ex6[perm.long()].to(torch.int32))[:, None]
        err = ((rec - want_d).abs() / rowmax).max().item()
        assert err < 2.0 ** -43, err
    else:
        hi, lo = d2[:, :K], d2[:, K:]
        assert torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros_like(hi, dtype=torch.int32))
        assert torch.equal(lo.view(torch.int32) & 0x1FFF, torch.zeros_like(lo, dtype=torch.int32))
        rowmax = want_d.abs().amax(1, keepdim=True)
        assert (((hi + lo) - want_d).abs() / rowmax).max().item() < 2e-6
        # mode 3: four int8 digit slices of the 32-bit fixed-point image + row exponent (fp32 only)
        out, d3, ex = ops.tabulate_sections_desc(tables, infos, em, sec, M, axis, 1.0 / nnei, desc_row=perm, mode=3,
                                                 nslice=4, pad_rows=32)
        assert torch.equal(out, want_out)
        assert d3.dtype == torch.int8 and d3.shape == (nloc + 32, 4 * K) and int(d3[nloc:].abs().sum()) == 0
        sl = d3[perm.long()].reshape(nloc, 4, K).to(torch.float64)
        assert int(sl.abs().max()) <= 128
        w = torch.tensor([2.0 ** (-7 - 8 * s) for s in range(4)], dtype=torch.float64, device=DEV)
        rec = (sl * w[None, :, None]).sum(1) * torch.ldexp(torch.ones((), dtype=torch.float64, device=DEV),
                                                           ex[perm.long()].to(torch.int32))[:, None]
        wd = want_d.to(torch.float64)
        # the image resolves 2^-31 of 2^row_exp; the rest is fp32 rounding of D itself
        assert ((rec - wd).abs() / rowmax.to(torch.float64)).max().item() < 2e-6
        assert bool((wd.abs().amax(1) < torch.ldexp(torch.ones(nloc, dtype=torch.float64, device=DEV),
                                                     ex[perm.long()].to(torch.int32) - 1)).all())
    with pytest.raises(ValueError):
        ops.tabulate_sections_desc(tables, infos, em, sec, M, axis, 1.0 / nnei, mode=3, nslice=6)


def test_split_i8_gemm_matches_fp64(ops):
    """sum over orders of K-concatenated int8 GEMMs == the fp64 product to ~2^-45 of the row/column scales."""
    from deepmd_kit_b200.model import split_i8_cols

    torch.manual_seed(5)
    n, K, Nn, ns = 300, 1600, 240, 7
    x = (torch.randn(n, K, dtype=torch.float64, device=DEV) * torch.logspace(-6, 0, n, dtype=torch.float64, device=DEV)[:, None])
    w = torch.randn(K, Nn, dtype=torch.float64) * 0.03
    xs, ex = ops.split_i8_rows(x, ns)
    sl, ce = split_i8_cols(w, ns)
    wrev = torch.cat([sl[k] for k in range(ns - 1, -1, -1)], 0).contiguous().to(DEV)
    acc = torch.empty((ns, n, Nn), dtype=torch.int32, device=DEV)
    for d in range(ns):
        torch._int_mm(xs[:, : (d + 1) * K], wrev[(ns - 1 - d) * K:], out=acc[d])
    z = ops.split_i8_combine(acc, ex, ce.to(DEV), activation=False)
    want = x @ w.to(DEV)
    scale = x.abs().amax(1, keepdim=True) * w.abs().amax(0, keepdim=True).to(DEV) * K ** 0.5
    assert ((z - want).abs() / scale).max().item() < 1e-13
    # with the fused layer epilogue
    b = torch.randn(Nn, dtype=torch.float64, device=DEV)
    idt = torch.rand(Nn, dtype=torch.float64, device=DEV)
    a, y = ops.split_i8_combine(acc, ex, ce.to(DEV), bias=b, idt=idt)
    assert torch.allclose(a, torch.tanh(want + b), rtol=0, atol=1e-11)
    assert torch.allclose(y, torch.tanh(want + b) * idt, rtol=0, atol=1e-11)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fitting_split_matches_plain(ops, dtype):
    """Tensor-core fitting net (int8 split / 3xTF32) against the plain GEMM forward/backward."""
    from deepmd_kit_b200.model import FittingNet

    torch.manual_seed(2)
    n, K = 700, 1600
    net = FittingNet(K, (240, 240, 240), True, 7, dtype, DEV).prepare_split(7)
    d = torch.randn(n, K, dtype=dtype, device=DEV) * 0.02
    e0, g0 = net.forward_backward(d)
    if dtype == torch.float64:
        xs, ex = ops.split_i8_rows(d, 7)
        e1, g1 = net.forward_backward_split(xs, ex, n)
        tol = 1e-11
        ref_e, ref_g = e0, g0
    else:
        xs = ops.split_tf32(d, 2)
        e1, g1 = net.forward_backward_split(xs, None, n)
        net64 = FittingNet(K, (240, 240, 240), True, 7, torch.float64, DEV)
        ref_e, ref_g = net64.forward_backward(d.double())
        tol = 1e-5
        # the plain fp32 path is the yardstick: the split path must not be worse than twice its error
        tol_e = max(tol, 2 * ((e0.double() - ref_e).abs().max() / ref_e.abs().max()).item())
        assert ((e1.double() - ref_e).abs().max() / ref_e.abs().max()).item() <= tol_e
    assert ((e1.double() - ref_e.double()).abs().max() / ref_e.abs().max()).item() <= tol
    assert ((g1.double() - ref_g.double()).abs().max() / ref_g.abs().max()).item() <= tol


def test_compressed_coefficients_match_full_table(ops):
    """DPB200_TAB_COMPRESSED_COEF (fp32 a3/a4, fp16 a5 on the stride-0 rows of a dp-compress table) against the
    full fp64 table on the same inputs, forward + fused descriptor and backward; inputs reach the stride-1
    (never compressed) rows and both extrapolation branches as well."""
    import __graft_entry__ as g

    g.load_package()
    from deepmd_kit_b200.model import SeAConfig, SeAModel

    cfg = SeAConfig()
    model = SeAModel(cfg, torch.float64, DEV)
    assert model.coef_flags is not None and all(f & 1 for f in model.coef_flags)
    rng = np.random.default_rng(3)
    nloc, nnei = 301, cfg.nnei
    em = rng.normal(scale=0.4, size=(nloc, nnei, 4))
    x = np.sort(rng.uniform(-0.9, 4.0, size=(nloc, nnei)), axis=1)[:, ::-1].copy()
    x[:7, :5] = rng.uniform(9.5, 40.0, size=(7, 5))   # stride-1 rows
    x[7:9, :3] = 50.0                                  # beyond max
    x[9:11, -3:] = -1.5                                # below lower
    em[:, :, 0] = x
    for i in range(nloc):  # trailing padding in both sections
        for a, b in ((0, 46), (46, 138)):
            k = int(rng.integers(0, (b - a) // 2))
            if k:
                em[i, b - k:b, 0] = -0.36
                em[i, b - k:b, 1:] = 0
    em_t = T(em.reshape(nloc, -1))
    inv = 1.0 / nnei
    out0, d0, e0 = ops.tabulate_sections_desc(model.tables, model.infos, em_t, cfg.sec, model.M, 16, inv, mode=2, nslice=6)
    out1, d1, e1 = ops.tabulate_sections_desc(model.tables, model.infos, em_t, cfg.sec, model.M, 16, inv, mode=2, nslice=6,
                                              flags=model.coef_flags)
    assert ((out1 - out0).abs().max() / out0.abs().max()).item() < 1e-11
    # the slices of the two descriptors may differ in the last digits only: compare the reconstructed rows
    K = model.M * 16
    w6 = torch.tensor([2.0 ** (-7 - 8 * s) for s in range(6)], dtype=torch.float64, device=DEV)
    r0_ = (d0[:nloc].reshape(nloc, 6, K).double() * w6[None, :, None]).sum(1) * torch.ldexp(
        torch.ones(nloc, dtype=torch.float64, device=DEV), e0[:nloc])[:, None]
    r1_ = (d1[:nloc].reshape(nloc, 6, K).double() * w6[None, :, None]).sum(1) * torch.ldexp(
        torch.ones(nloc, dtype=torch.float64, device=DEV), e1[:nloc])[:, None]
    assert ((r1_ - r0_).abs().max() / r0_.abs().max()).item() < 1e-11
    dy = torch.randn_like(out0)
    g0 = ops.tabulate_sections_grad(model.tables, model.infos, em_t, dy, cfg.sec, model.M)
    g1 = ops.tabulate_sections_grad(model.tables, model.infos, em_t, dy, cfg.sec, model.M, flags=model.coef_flags)
    assert ((g1 - g0).abs().max() / g0.abs().max()).item() < 1e-11
    # a table that does not qualify is refused by the gate
    tab = random_table(57, 100, rng, np.float64)
    assert ops.compressed_coef_flags(torch.as_tensor(tab), np.array([-0.4, 2.0, 6.0, 0.05, 0.5, -1.0])) == 0


def test_compressed_coefficients_fp32_cubic(ops):
    """fp32 flavour: {a0..a3} only (float4 per row and channel) against the full fp32 table, forward and backward."""
    import __graft_entry__ as g

    g.load_package()
    from deepmd_kit_b200.model import SeAConfig, SeAModel

    cfg = SeAConfig()
    model = SeAModel(cfg, torch.float32, DEV)
    assert model.coef_flags is not None and all(model.coef_flags)
    rng = np.random.default_rng(4)
    nloc, nnei = 257, cfg.nnei
    em = rng.normal(scale=0.4, size=(nloc, nnei, 4)).astype(np.float32)
    x = np.sort(rng.uniform(-0.9, 4.0, size=(nloc, nnei)), axis=1)[:, ::-1].copy()
    x[:5, :4] = rng.uniform(9.5, 40.0, size=(5, 4))
    x[5:7, :2] = 50.0
    x[7:9, -3:] = -1.5
    em[:, :, 0] = x
    em_t = T(em.reshape(nloc, -1))
    inv = 1.0 / nnei
    out0, d0, _ = ops.tabulate_sections_desc(model.tables, model.infos, em_t, cfg.sec, model.M, 16, inv, mode=2)
    out1, d1, _ = ops.tabulate_sections_desc(model.tables, model.infos, em_t, cfg.sec, model.M, 16, inv, mode=2,
                                             flags=model.coef_flags)
    assert ((out1 - out0).abs().max() / out0.abs().max()).item() < 2e-6
    assert ((d1 - d0).abs().max() / d0.abs().max()).item() < 5e-6
    dy = torch.randn_like(out0)
    g0 = ops.tabulate_sections_grad(model.tables, model.infos, em_t, dy, cfg.sec, model.M)
    g1 = ops.tabulate_sections_grad(model.tables, model.infos, em_t, dy, cfg.sec, model.M, flags=model.coef_flags)
    assert ((g1 - g0).abs().max() / g0.abs().max()).item() < 5e-6


# ------------------------------------------------------------------------------------------------
# gradients of prod_force_a / prod_virial_a with respect to net_deriv
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prod_force_virial_grad_vs_oracle(ops, olib, dtype):
    """dpb200_prod_force_grad_a / _virial_grad_a against the restated reference (prod_force_grad.cc:22-77,
    prod_virial_grad.cc:21-63) on a periodic box, two frames, incl. ghost indices folded with j % nloc."""
    coord, atype, box = water_like_box(ncopy=2, seed=3, jitter=0.05)
    s = extended_system(olib, coord, atype, box, 6.0, dtype=np.float64)
    sec = [0, 20, 60]
    nnei = sec[-1]
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    nlist, _ = olib.format_nlist(s["coord"], s["atype"], off, neigh, 6.0, sec)
    em, dv, rij = olib.env_mat_a(s["coord"], s["atype"], nlist, 0.5, 6.0, sec)
    nloc = s["nloc"]
    rng = np.random.default_rng(9)
    dv, rij = dv.astype(dtype), rij.astype(dtype)
    gf = rng.normal(size=(2 * nloc, 3)).astype(dtype)
    gv = rng.normal(size=9).astype(dtype)
    dv2, nl2 = np.concatenate([dv, dv]), np.concatenate([nlist, nlist])
    want_f = olib.prod_force_grad_a(gf, dv2, nl2, nframes=2)
    want_v = olib.prod_virial_grad_a(gv, dv, rij, nlist)
    got_f = ops.prod_force_grad_a(T(gf.reshape(2, -1)), T(dv2), T(nl2), nloc, nnei, nframes=2)
    got_v = ops.prod_virial_grad_a(T(gv), T(dv), T(rij), T(nlist), nloc, nnei)
    close(N(got_f), want_f, dtype, fac=4)
    close(N(got_v), want_v, dtype, fac=4)
    assert (nlist >= nloc).any()  # ghosts present: the j % nloc fold was exercised


def test_force_virial_ops_are_differentiable(ops, olib):
    """torch.ops.deepmd.prod_force_se_a / prod_virial_se_a backward == the adjoint of the forward (exact for ghosts:
    ngrad = nall), and the explicit *_grad ops follow the TF schemas."""
    torch.ops.deepmd  # registered at import
    coord, atype, box = water_like_box(ncopy=1, seed=4, jitter=0.05)
    s = extended_system(olib, coord, atype, box, 6.0, dtype=np.float64)
    sec = [0, 20, 60]
    nnei = sec[-1]
    off, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    nlist, _ = olib.format_nlist(s["coord"], s["atype"], off, neigh, 6.0, sec)
    em, dv, rij = olib.env_mat_a(s["coord"], s["atype"], nlist, 0.5, 6.0, sec)
    nloc, nall = s["nloc"], len(s["atype"])
    natoms = torch.tensor([nloc, nall, 0, 0], dtype=torch.int32)
    rng = np.random.default_rng(1)
    nd = T(rng.normal(size=(1, nloc * nnei * 4))).requires_grad_(True)
    dv_t, rij_t, nl_t = T(dv.reshape(1, -1)), T(rij.reshape(1, -1)), T(nlist.reshape(1, -1))
    force = torch.ops.deepmd.prod_force_se_a(nd, dv_t, nl_t, natoms, nnei, 0)
    virial, _ = torch.ops.deepmd.prod_virial_se_a(nd, dv_t, rij_t, nl_t, natoms, nnei, 0)
    wf, wv = torch.randn_like(force), torch.randn_like(virial)
    (g,) = torch.autograd.grad((force * wf).sum() + (virial * wv).sum(), nd)
    # adjoint through linearity: d/d(nd) <w, F(nd)> = F^T w, checked with a random direction
    u = torch.randn_like(nd)
    fu = torch.ops.deepmd.prod_force_se_a(u, dv_t, nl_t, natoms, nnei, 0)
    vu, _ = torch.ops.deepmd.prod_virial_se_a(u, dv_t, rij_t, nl_t, natoms, nnei, 0)
    lhs = (g * u).sum().item()
    rhs = ((fu * wf).sum() + (vu * wv).sum()).item()
    assert abs(lhs - rhs) <= 1e-10 * abs(rhs)
    # explicit grad ops (TF schema: grad over the nloc local atoms)
    gf = torch.randn(1, nloc * 3, dtype=torch.float64, device=DEV)
    got = torch.ops.deepmd.prod_force_se_a_grad(gf, nd.detach(), dv_t, nl_t, natoms, nnei, 0)
    close(N(got), olib.prod_force_grad_a(N(gf).reshape(nloc, 3), dv, nlist), np.float64, fac=4)
    gvv = torch.randn(1, 9, dtype=torch.float64, device=DEV)
    got = torch.ops.deepmd.prod_virial_se_a_grad(gvv, nd.detach(), dv_t, rij_t, nl_t, natoms, nnei, 0)
    close(N(got), olib.prod_virial_grad_a(N(gvv).reshape(9), dv, rij, nlist), np.float64, fac=4)
