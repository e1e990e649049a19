"""Pins the CPU restatement (oracle/dp_oracle.c) — CPU only, no GPU, no product code.

1. against the literal golden vectors of the reference's own unit tests (tests/golden/*.json);
2. against the unmodified reference library compiled from /root/reference (oracle/_ref), on the
   golden system and on seeded random systems, fp64 and fp32.  Integer outputs must be identical;
   floating-point outputs are required to be *bitwise* identical as well (same operation order,
   same build flags), which is stronger than the 1e-10 the reference asks of itself
   (source/lib/tests/test_env_mat_a.cc:540-590).
"""
import numpy as np
import pytest

from oracle import cpu as ocpu
from _systems import extended_system, golden, random_table, six_atom_system, water_like_box


def csr(sysd):
    return ocpu.dense_to_csr(sysd["rows"], sysd["numneigh"])


# ----------------------------------------------------------------------------- goldens
def test_golden_copy_coord_and_raw_list(port):
    g = golden("neighbor_list.json")["TestNeighborList"]
    s = six_atom_system(port)
    assert s["numneigh"].max() == 5
    for i, want in enumerate(g["expect_nlist_cpy"]):
        got = sorted(s["rows"][i, : s["numneigh"][i]].tolist())
        assert got == sorted(want)


@pytest.mark.parametrize("cls", ["TestCopyCoord", "TestCopyCoordMoreCell"])
def test_golden_copy_coord(port, cls):
    g = golden("coord.json")[cls]
    posi = np.array(g["posi"]).reshape(-1, 3)
    atype = np.array(g["atype"], np.int32)
    box = np.array(g["boxt"]).reshape(3, 3)
    c, t, m = port.copy_coord(posi, atype, box, g["rc"])
    want_c = np.array(g["_expected_posi_cpy"]).reshape(-1, 3)
    want_t = np.array(g["_expected_atype_cpy"])
    want_m = np.array(g["_expected_mapping"])
    assert len(t) == len(want_t)
    nloc = len(atype)
    # the reference test sorts ghosts before comparing (test_coord.cc:137-165, 236-252)
    def key(cc, tt, mm):
        rows = [tuple(np.round(cc[i], 9)) + (int(tt[i]), int(mm[i])) for i in range(nloc, len(tt))]
        return sorted(rows)
    np.testing.assert_allclose(c[:nloc], want_c[:nloc], atol=1e-12)
    got, want = key(c, t, m), key(want_c, want_t, want_m)
    assert [r[3:] for r in got] == [r[3:] for r in want]
    np.testing.assert_allclose(np.array([r[:3] for r in got]), np.array([r[:3] for r in want]), atol=1e-9)
    # too little memory => status 1 and the needed size (test_coord.cc:255-272)
    _, _, need = port.copy_coord(posi, atype, box, g["rc"], mem_nall=40)
    assert need == len(want_t)


def test_golden_normalize_coord(port):
    g = golden("coord.json")["TestNormCoord"]
    box = np.array(g["boxt"]).reshape(3, 3)
    for k in ("r0", "r1", "r2"):
        out = port.normalize_coord(np.array(g[k]).reshape(-1, 3), box)
        np.testing.assert_allclose(out.reshape(-1), g["posi"], atol=1e-12)


@pytest.mark.parametrize("cls", ["TestFormatNlist", "TestFormatNlistShortSel"])
def test_golden_format_nlist(port, cls):
    g = golden("fmt_nlist.json")[cls]
    s = six_atom_system(port, rc=g["rc"])
    off, neigh = csr(s)
    nlist, over = port.format_nlist(s["coord"], s["atype"], off, neigh, g["rc"], g["sec_a"])
    assert nlist.reshape(-1).tolist() == g["expect_nlist_cpy"]
    if cls == "TestFormatNlistShortSel":
        # test_fmt_nlist.cc:257-317: type-1 section overflows for every atom
        assert (over >= 0).any()
    else:
        assert (over == -1).all()


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-5), (np.float32, 1e-5)])
def test_golden_env_mat(port, dtype, tol):
    g = golden("env_mat_a.json")["TestEnvMatA"]
    s = six_atom_system(port, rc=g["rc"], dtype=dtype)
    off, neigh = csr(s)
    nnei = g["sec_a"][-1]
    avg = np.zeros((2, nnei * 4), dtype)
    std = np.ones((2, nnei * 4), dtype)
    em, dv, rij, nlist = port.prod_env_mat_a(s["coord"], s["atype"], off, neigh, avg, std, s["nloc"], g["rc"],
                                            g["rc_smth"], g["sec_a"])
    np.testing.assert_allclose(em.reshape(-1), g["expected_env"], atol=tol)
    # prod == format + per-atom env (test_env_mat_a.cc:540-590)
    nl2, _ = port.format_nlist(s["coord"], s["atype"], off, neigh, g["rc"], g["sec_a"])
    assert (nl2 == nlist).all()
    em2, dv2, rij2 = port.env_mat_a(s["coord"], s["atype"], nl2, g["rc_smth"], g["rc"], g["sec_a"])
    assert (em2 == em).all() and (dv2 == dv).all() and (rij2 == rij).all()


def test_golden_tabulate(port):
    g = golden("tabulate_se_a.json")["TestTabulateSeA"]
    nloc, nnei, M = g["nloc"], g["nnei"], g["last_layer_size"]
    table = np.array(g["table"])
    em_x = np.array(g["em_x"]).reshape(nloc, nnei)
    em = np.array(g["em"]).reshape(nloc, nnei, 4)
    two = np.array(g["two_embed"]).reshape(nloc, nnei, M)
    dy = np.ones((nloc, 4, M))
    out = port.tabulate_fusion_se_a(table, g["info"], em_x, em, M)
    np.testing.assert_allclose(out.reshape(-1), g["expected_xyz_scatter"], rtol=0, atol=1e-5)
    out2 = port.tabulate_fusion_se_a(table, g["info"], em_x, em, M, two_embed=two)
    np.testing.assert_allclose(out2.reshape(-1), g["expected_xyz_scatter_with_two_embed"], atol=1e-5)
    gx, gem, _ = port.tabulate_fusion_se_a_grad(table, g["info"], em_x, em, dy, M)
    np.testing.assert_allclose(gx.reshape(-1), g["expected_dy_dem_x"], atol=1e-5)
    np.testing.assert_allclose(gem.reshape(-1), g["expected_dy_dem"], atol=1e-5)
    gx, gem, gtwo = port.tabulate_fusion_se_a_grad(table, g["info"], em_x, em, dy, M, two_embed=two)
    np.testing.assert_allclose(gx.reshape(-1), g["expected_dy_dem_x_with_two_embed"], atol=1e-5)
    np.testing.assert_allclose(gem.reshape(-1), g["expected_dy_dem_with_two_embed"], atol=1e-5)
    # empty neighbour axis is a valid empty reduction (test_tabulate_se_a.cc:761-786)
    z = port.tabulate_fusion_se_a(table, g["info"], np.zeros((nloc, 0)), np.zeros((nloc, 0, 4)), M)
    assert (z == 0).all()


def _force_virial_inputs(port, g):
    s = six_atom_system(port, rc=g["rc"])
    off, neigh = csr(s)
    sec = g["sec_a"]
    nnei = sec[-1]
    nlist, _ = port.format_nlist(s["coord"], s["atype"], off, neigh, g["rc"], sec)
    em, dv, rij = port.env_mat_a(s["coord"], s["atype"], nlist, g["rc_smth"], g["rc"], sec)
    nd = 10 - 0.01 * np.arange(s["nloc"] * nnei * 4, dtype=np.float64)
    return s, nlist, dv, rij, nd.reshape(s["nloc"], -1)


def test_golden_prod_force(port):
    g = golden("prod_force_a.json")["TestProdForceA"]
    s, nlist, dv, rij, nd = _force_virial_inputs(port, g)
    nall = len(s["atype"])
    f = port.prod_force_a(nd, dv, nlist, nall)
    np.testing.assert_allclose(f.reshape(-1), g["expected_force"][: nall * 3], atol=1e-5)


def test_golden_prod_virial(port):
    g = golden("prod_virial_a.json")["TestProdVirialA"]
    s, nlist, dv, rij, nd = _force_virial_inputs(port, g)
    nall = len(s["atype"])
    v, av = port.prod_virial_a(nd, dv, rij, nlist, nall)
    np.testing.assert_allclose(v, g["expected_virial"], atol=1e-5)
    np.testing.assert_allclose(av.reshape(-1), g["expected_atom_virial"], atol=1e-5)


def test_golden_prod_force_grad(port):
    """source/lib/tests/test_prod_force_grad_a.cc:16-131: grad = 10 - 0.1*k, two identical frames."""
    g = golden("prod_force_grad_a.json")["TestProdForceGradA"]
    s, nlist, dv, rij, _ = _force_virial_inputs(port, g)
    nloc = s["nloc"]
    grad = (10 - 0.1 * np.arange(nloc * 3, dtype=np.float64)).reshape(nloc, 3)
    out = port.prod_force_grad_a(np.concatenate([grad, grad]), np.concatenate([dv, dv]), np.concatenate([nlist, nlist]),
                                 nframes=2)
    np.testing.assert_allclose(out.reshape(-1), np.tile(g["expected_grad_net"], 2), atol=1e-5)


def test_golden_prod_virial_grad(port):
    """source/lib/tests/test_prod_virial_grad_a.cc:11-135: grad = 10 - k."""
    g = golden("prod_virial_grad_a.json")["TestProdVirialGradA"]
    s, nlist, dv, rij, _ = _force_virial_inputs(port, g)
    out = port.prod_virial_grad_a(10 - np.arange(9, dtype=np.float64), dv, rij, nlist)
    np.testing.assert_allclose(out.reshape(-1), g["expected_grad_net"], atol=1e-5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_grad_ops_port_matches_reference_and_adjoint(port, ref, dtype):
    """The restated gradient ops equal the reference library bitwise, and they ARE the adjoints of
    prod_force_a / prod_virial_a:  <grad, force(nd)> = <force_grad(grad), nd>."""
    g = golden("prod_force_a.json")["TestProdForceA"]
    s, nlist, dv, rij, nd = _force_virial_inputs(port, g)
    rng = np.random.default_rng(5)
    nloc, nall = s["nloc"], len(s["atype"])
    dv, rij, nd = dv.astype(dtype), rij.astype(dtype), rng.normal(size=nd.shape).astype(dtype)
    gf = rng.normal(size=(nloc, 3)).astype(dtype)
    gv = rng.normal(size=9).astype(dtype)
    a = port.prod_force_grad_a(gf, dv, nlist)
    b = port.prod_virial_grad_a(gv, dv, rij, nlist)
    assert (a == ref.prod_force_grad_a(gf, dv, nlist)).all()
    np.testing.assert_allclose(b, ref.prod_virial_grad_a(gv, dv, rij, nlist), rtol=0, atol=(1e-13 if dtype == np.float64 else 1e-5))
    if dtype == np.float64:
        # adjoint identity on the local block (ghosts folded with j % nloc, as prod_force_grad does)
        nl_loc = np.where(nlist >= 0, nlist % nloc, -1).astype(np.int32)
        f = port.prod_force_a(nd, dv, nl_loc, nloc)
        assert abs((gf * f).sum() - (a * nd).sum()) < 1e-10 * abs((a * nd).sum())
        v, _ = port.prod_virial_a(nd, dv, rij, nlist, nall)
        assert abs((gv * v).sum() - (b * nd).sum()) < 1e-10 * abs((b * nd).sum())


# ----------------------------------------------------------- restatement == reference
def test_ref_legacy_fixture_matches_port(port, ref):
    g = golden("env_mat_a.json")["TestEnvMatA"]
    leg = ref.legacy_copy_and_build(np.array(g["posi"]).reshape(-1, 3), g["atype"], np.diag([13.0] * 3), 6.0, 6.0)
    s = six_atom_system(port)
    assert (leg["coord"] == s["coord"]).all() and (leg["atype"] == s["atype"]).all()
    assert (leg["mapping"] == s["mapping"]).all()
    for i in range(s["nloc"]):
        a = sorted(leg["neigh"][leg["offsets"][i]: leg["offsets"][i + 1]].tolist())
        assert a == s["rows"][i, : s["numneigh"][i]].tolist()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("jitter", [0.0, 0.05])
def test_port_equals_reference_geometry(port, ref, dtype, jitter):
    coord, atype, box = water_like_box(ncopy=2, seed=3, jitter=jitter, dtype=dtype)
    a = port.normalize_coord(coord, box)
    b = ref.normalize_coord(coord, box)
    assert (a == b).all()
    ca, ta, ma = port.copy_coord(a, atype, box, 7.0)
    cb, tb, mb = ref.copy_coord(a, atype, box, 7.0)
    assert ca.shape == cb.shape and (ca == cb).all() and (ta == tb).all() and (ma == mb).all()
    assert (port.compute_cell_info(box, 7.0) == ref.compute_cell_info(box, 7.0)).all()
    nloc = len(atype)
    na, ra = port.build_nlist(ca, nloc, 7.0)
    nb, rb = ref.build_nlist(ca, nloc, 7.0)
    assert (na == nb).all()
    w = min(ra.shape[1], rb.shape[1])
    mask = np.arange(w)[None, :] < na[:, None]
    assert (ra[:, :w][mask] == rb[:, :w][mask]).all()
    # row capacity too small => status 1 + needed size (neighbor_list.cc:914-917)
    # (the reference stops at the first overflowing row and reports that row's size; we report the max)
    _, need = port.build_nlist(ca, nloc, 7.0, mem_size=8)
    assert need == na.max() and 8 < ref.build_nlist(ca, nloc, 7.0, mem_size=8)[1] <= need


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("jitter", [0.0, 0.03])
@pytest.mark.parametrize("sel", [(46, 92), (6, 10)])
def test_port_equals_reference_env_mat(port, ref, dtype, jitter, sel):
    coord, atype, box = water_like_box(ncopy=2, seed=5, jitter=jitter, dtype=dtype)
    s = extended_system(port, coord, atype, box, 8.0, dtype=dtype)
    rng = np.random.default_rng(1)
    s["rows"] = np.ascontiguousarray(rng.permuted(s["rows"], axis=1)) if False else s["rows"]
    off, neigh = csr(s)
    # shuffle each raw row: formatting must not depend on the input order
    for i in range(s["nloc"]):
        seg = neigh[off[i]: off[i + 1]]
        rng.shuffle(seg)
    sec = np.array([0, sel[0], sel[0] + sel[1]], np.int32)
    nnei = int(sec[-1])
    avg = rng.normal(size=(2, nnei * 4)).astype(dtype) * 0.1
    std = (0.5 + rng.random(size=(2, nnei * 4))).astype(dtype)
    A = port.prod_env_mat_a(s["coord"], s["atype"], off, neigh, avg, std, s["nloc"], 6.0, 0.5, sec)
    B = ref.prod_env_mat_a(s["coord"], s["atype"], off, neigh, avg, std, s["nloc"], 6.0, 0.5, sec)
    assert (A[3] == B[3]).all(), "formatted neighbour list must be bit-identical"
    for x, y in zip(A[:3], B[:3]):
        assert (x == y).all()
    fa, oa = port.format_nlist(s["coord"], s["atype"], off, neigh, 6.0, sec)
    fb, ob = ref.format_nlist(s["coord"], s["atype"], off, neigh, 6.0, sec)
    assert (fa == fb).all() and (oa == ob).all()
    if sel == (6, 10):
        assert (oa >= 0).any()
    # virtual atoms (type < 0) are never neighbours and produce all-zero rows
    ty2 = s["atype"].copy()
    ty2[::7] = -1
    A = port.prod_env_mat_a(s["coord"], ty2, off, neigh, avg, std, s["nloc"], 6.0, 0.5, sec)
    B = ref.prod_env_mat_a(s["coord"], ty2, off, neigh, avg, std, s["nloc"], 6.0, 0.5, sec)
    for x, y in zip(A, B):
        assert (x == y).all()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("two", [False, True])
@pytest.mark.parametrize("is_sorted", [True, False])
@pytest.mark.parametrize("nd", [4, 9, 16, 25])
def test_port_equals_reference_tabulate(port, ref, dtype, two, is_sorted, nd):
    """nd = NDESCRPT (tabulate.cc:456-560 dispatch: 4 for se_a / se_atten, 9 / 16 / 25 for the higher angular bases)."""
    rng = np.random.default_rng(7)
    nloc, nnei, M = 9, 23, 20
    lower, upper, vmax, s0, s1 = -0.5, 2.0, 10.0, 0.01, 0.1
    nspline = int((upper - lower) / s0 + (vmax - upper) / s1)
    info = np.array([lower, upper, vmax, s0, s1, -1], dtype)
    table = random_table(nspline, M, rng, dtype)
    em_x = rng.uniform(-1.0, 11.0, size=(nloc, nnei)).astype(dtype)  # both extrapolation sides
    em_x[0, :3] = [lower, upper, vmax]  # exact boundaries
    em = rng.normal(size=(nloc, nnei, nd)).astype(dtype)
    # trailing padding: identical em_x, zero angular part
    for i in range(nloc):
        npad = i % 5
        if npad:
            em_x[i, nnei - npad:] = em_x[i, nnei - 1]
            em[i, nnei - npad:, 1:] = 0
            em[i, nnei - npad:, 0] = em[i, nnei - 1, 0]
    te = rng.normal(size=(nloc, nnei, M)).astype(dtype) if two else None
    dy = rng.normal(size=(nloc, nd, M)).astype(dtype)
    a = port.tabulate_fusion_se_a(table, info, em_x, em, M, te, is_sorted)
    assert a.shape == (nloc, nd, M)
    b = ref.tabulate_fusion_se_a(table, info, em_x, em, M, te, is_sorted)
    assert (a == b).all()
    ga = port.tabulate_fusion_se_a_grad(table, info, em_x, em, dy, M, te, is_sorted)
    gb = ref.tabulate_fusion_se_a_grad(table, info, em_x, em, dy, M, te, is_sorted)
    for x, y in zip(ga, gb):
        assert (x is None and y is None) or (x == y).all()
    dzx = rng.normal(size=(nloc, nnei)).astype(dtype)
    dze = rng.normal(size=(nloc, nnei, nd)).astype(dtype)
    dzt = rng.normal(size=(nloc, nnei, M)).astype(dtype) if two else None
    ha = port.tabulate_fusion_se_a_grad_grad(table, info, em_x, em, dzx, dze, M, te, dzt, is_sorted)
    hb = ref.tabulate_fusion_se_a_grad_grad(table, info, em_x, em, dzx, dze, M, te, dzt, is_sorted)
    assert (ha == hb).all()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_port_equals_reference_force_virial(port, ref, dtype):
    rng = np.random.default_rng(11)
    nloc, nall, nnei = 40, 70, 18
    nlist = rng.integers(-1, nall, size=(nloc, nnei)).astype(np.int32)
    nd = rng.normal(size=(nloc, nnei * 4)).astype(dtype)
    dv = rng.normal(size=(nloc, nnei * 12)).astype(dtype)
    rij = rng.normal(size=(nloc, nnei * 3)).astype(dtype)
    assert (port.prod_force_a(nd, dv, nlist, nall) == ref.prod_force_a(nd, dv, nlist, nall)).all()
    va, aa = port.prod_virial_a(nd, dv, rij, nlist, nall)
    vb, ab = ref.prod_virial_a(nd, dv, rij, nlist, nall)
    # the reference accumulates with `omp atomic` in arbitrary order -> tolerance, not bits
    tol = 1e-12 if dtype == np.float64 else 2e-5
    np.testing.assert_allclose(va, vb, rtol=tol, atol=tol * 10)
    np.testing.assert_allclose(aa, ab, rtol=tol, atol=tol * 10)
