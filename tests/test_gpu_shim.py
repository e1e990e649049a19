"""The reference-side C++ binding on a real GPU: tests/shim/_build/shim_driver (compiled here against
the reference's own headers, see tests/shim/build.sh) calls deepmd::prod_env_mat_a_gpu,
tabulate_fusion_se_a_gpu, tabulate_fusion_se_a_grad_gpu, prod_force_a_gpu and prod_virial_a_gpu from
lib/libdeepmd_op_cuda.so exactly as the reference's op layers do; outputs are checked against the
CPU oracle.  Plus a CPU-side check of the exported (mangled) symbol set."""
import os
import subprocess

import numpy as np
import pytest

from oracle import cpu as ocpu
from _systems import extended_system, random_table, water_like_box

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "shim", "_build", "shim_driver")
SHIM = os.path.join(ROOT, "deepmd-kit_b200", "lib", "libdeepmd_op_cuda.so")

WANT = ["prod_env_mat_a_gpu", "format_nbor_list_gpu", "tabulate_fusion_se_a_gpu", "tabulate_fusion_se_a_grad_gpu",
        "tabulate_fusion_se_a_grad_grad_gpu", "prod_force_a_gpu", "prod_virial_a_gpu", "prod_force_grad_a_gpu",
        "prod_virial_grad_a_gpu", "normalize_coord_gpu",
        "copy_coord_gpu", "build_nlist_gpu"]


def test_shim_exports_reference_symbols():
    if not os.path.exists(SHIM):
        pytest.skip("libdeepmd_op_cuda.so not built (needs the reference headers)")
    out = subprocess.run(["nm", "-DC", "--defined-only", SHIM], capture_output=True, text=True, check=True).stdout
    for name in WANT:
        for fp in ("double", "float"):
            assert f"deepmd::{name}<{fp}>(" in out, f"missing deepmd::{name}<{fp}>"
    assert "deepmd::use_nlist_map(int*, int const*, int, int)" in out
    # Itanium-mangled names as a caller compiled against the reference headers would reference them
    raw = subprocess.run(["nm", "-D", "--defined-only", SHIM], capture_output=True, text=True, check=True).stdout
    assert "_ZN6deepmd16prod_force_a_gpuIdEEvPT_PKS1_S4_PKiiiii" in raw


@pytest.mark.gpu
def test_shim_driver_matches_oracle(olib, tmp_path):
    if not os.path.exists(DRIVER):
        pytest.skip("tests/shim/_build/shim_driver not built (needs the reference headers)")
    rng = np.random.default_rng(5)
    coord, atype, box = water_like_box(ncopy=2, seed=4, jitter=0.05)
    s = extended_system(olib, coord, atype, box, 6.5)
    sec = np.array([0, 46, 138], np.int32)
    nnei, nloc, nall = 138, s["nloc"], len(s["atype"])
    avg = rng.normal(scale=0.05, size=(2, nnei * 4))
    avg[:, 1::4] = avg[:, 2::4] = avg[:, 3::4] = 0
    std = 0.08 + 0.1 * rng.random(size=(2, nnei * 4))
    M = 100
    info = np.array([-1.0, 3.0, 15.0, 0.05, 0.5, -1.0])
    nspline = int((info[1] - info[0]) / info[3] + (info[2] - info[1]) / info[4])
    table = random_table(nspline, M, rng)
    nd = rng.normal(size=(nloc, nnei * 4))
    dy = rng.normal(size=(nloc, 4, M))
    max_nbor = s["rows"].shape[1]
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        np.array([nloc, nall, nnei, 2, max_nbor, nspline, M], np.int32).tofile(f)
        sec.tofile(f)
        np.array([6.0, 0.5], np.float32).tofile(f)
        for a, dt in ((s["coord"], np.float64), (s["atype"], np.int32), (s["numneigh"], np.int32), (s["rows"], np.int32),
                      (avg, np.float64), (std, np.float64), (table, np.float64), (info, np.float64), (nd, np.float64),
                      (dy, np.float64)):
            np.ascontiguousarray(a, dtype=dt).tofile(f)
    r = subprocess.run([DRIVER, str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHIM_DRIVER_OK" in r.stdout, r.stderr[-2000:]
    raw = open(fout, "rb").read()
    off = 0

    def take(n, dt):
        nonlocal off
        a = np.frombuffer(raw, dtype=dt, count=n, offset=off)
        off += n * np.dtype(dt).itemsize
        return a

    em, dv, rij = take(nloc * nnei * 4, np.float64), take(nloc * nnei * 12, np.float64), take(nloc * nnei * 3, np.float64)
    nlist = take(nloc * nnei, np.int32)
    desc, gx, gem = take(nloc * 4 * M, np.float64), take(nloc * nnei, np.float64), take(nloc * nnei * 4, np.float64)
    force, virial, av = take(nall * 3, np.float64), take(9, np.float64), take(nall * 9, np.float64)
    gn_f, gn_v = take(nloc * nnei * 4, np.float64), take(nloc * nnei * 4, np.float64)
    o, neigh = ocpu.dense_to_csr(s["rows"], s["numneigh"])
    w_em, w_dv, w_rij, w_nl = olib.prod_env_mat_a(s["coord"], s["atype"], o, neigh, avg, std, nloc, 6.0, 0.5, sec)
    assert np.array_equal(nlist.reshape(nloc, nnei), w_nl)

    def close(a, b, tol=1e-10):
        b = np.asarray(b).reshape(-1)
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)

    close(em, w_em)
    close(dv, w_dv)
    close(rij, w_rij)
    em3 = w_em.reshape(nloc, nnei, 4)
    em_x = np.ascontiguousarray(em3[:, :, 0]).reshape(-1, 1)
    close(desc, olib.tabulate_fusion_se_a(table, info, em_x, em3, M), 4e-10)
    wx, wem, _ = olib.tabulate_fusion_se_a_grad(table, info, em_x, em3, dy, M)
    close(gx, wx, 4e-10)
    close(gem, wem, 4e-10)
    close(force, olib.prod_force_a(nd, w_dv, w_nl, nall), 4e-10)
    wv, wav = olib.prod_virial_a(nd, w_dv, w_rij, w_nl, nall)
    close(virial, wv, 2e-9)
    close(av, wav, 4e-10)
    wf = olib.prod_force_a(nd, w_dv, w_nl, nall)
    close(gn_f, olib.prod_force_grad_a(wf[:nloc], w_dv, w_nl), 4e-10)
    close(gn_v, olib.prod_virial_grad_a(wv, w_dv, w_rij, w_nl), 2e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("nd", [9, 16, 25])
def test_shim_tabulate_higher_basis(olib, tmp_path, nd):
    """deepmd::tabulate_fusion_se_a{,_grad,_grad_grad}_gpu with the trailing ndescrpt argument (tabulate.h:175-218),
    called from C++ exactly as source/op/pt/tabulate_multi_device.cc:101-117 does with ndescrpt = em.size(2); an
    unsupported basis dimension must throw as the reference's check_se_a_basis_dimension does."""
    if not os.path.exists(DRIVER):
        pytest.skip("tests/shim/_build/shim_driver not built (needs the reference headers)")
    rng = np.random.default_rng(40 + nd)
    nloc, nnei, M = 17, 29, 40
    info = np.array([-0.4, 2.0, 6.0, 0.05, 0.5, -1.0])
    nspline = int((info[1] - info[0]) / info[3]) + int((info[2] - info[1]) / info[4]) + 1
    table = random_table(nspline, M, rng)
    em_x = np.sort(rng.uniform(-0.8, 7.5, size=(nloc, nnei)), axis=1)[:, ::-1].copy()
    em = rng.normal(size=(nloc, nnei, nd))
    for i in range(nloc):  # trailing padding: constant em_x, zero angular part
        k = int(rng.integers(0, 6))
        if k:
            em_x[i, nnei - k:] = -0.37
            em[i, nnei - k:, 1:] = 0
    em[:, :, 0] = em_x
    dy = rng.normal(size=(nloc, nd, M))
    dzx = rng.normal(size=(nloc, nnei))
    dzem = rng.normal(size=(nloc, nnei, nd))
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        np.array([nloc, nnei, M, nd, nspline, 1], np.int32).tofile(f)
        for a in (table, info, em_x, em, dy, dzx, dzem):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
    r = subprocess.run([DRIVER, "tabnd", str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHIM_DRIVER_OK" in r.stdout, r.stderr[-2000:]
    raw = np.fromfile(fout, dtype=np.float64)
    n1, n2, n3 = nloc * nd * M, nloc * nnei, nloc * nnei * nd
    desc, gx, gem, gg = raw[:n1], raw[n1:n1 + n2], raw[n1 + n2:n1 + n2 + n3], raw[n1 + n2 + n3:]
    ex = em_x.reshape(-1, 1)

    def close(a, b, tol=1e-10):
        b = np.asarray(b).reshape(-1)
        assert a.shape == b.shape and np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)

    close(desc, olib.tabulate_fusion_se_a(table, info, ex, em, M))
    wx, wem, _ = olib.tabulate_fusion_se_a_grad(table, info, ex, em, dy, M)
    close(gx, wx)
    close(gem, wem)
    close(gg, olib.tabulate_fusion_se_a_grad_grad(table, info, ex, em, dzx.reshape(-1, 1), dzem, M))


@pytest.mark.gpu
def test_shim_neighbour_front_end(olib, tmp_path):
    """normalize_coord_gpu / copy_coord_gpu / build_nlist_gpu driven as _norm_copy_coord_gpu and _build_nlist_gpu of
    source/op/tf/prod_env_mat_multi_device.cc:2399-2600 drive them: Region and cell_info in DEVICE memory, rows
    written into the caller-owned jlist through firstneigh (which must stay untouched), two frames per call,
    `1` returned for a copy buffer / row capacity that is too small."""
    if not os.path.exists(DRIVER):
        pytest.skip("tests/shim/_build/shim_driver not built (needs the reference headers)")
    coord, atype, box = water_like_box(ncopy=1, seed=21, jitter=0.2)
    coord = coord + 3.7  # some atoms outside the cell: normalize_coord_gpu has work to do
    nloc, rc, nframes = len(atype), 6.0, 2
    fin, fout = tmp_path / "nl_in.bin", tmp_path / "nl_out.bin"
    mem_cpy = 40 * nloc
    with open(fin, "wb") as f:
        np.array([nloc, mem_cpy, nframes], np.int32).tofile(f)
        np.array([rc], np.float32).tofile(f)
        np.ascontiguousarray(box, np.float64).tofile(f)
        np.ascontiguousarray(coord, np.float64).tofile(f)
        np.ascontiguousarray(atype, np.int32).tofile(f)
    r = subprocess.run([DRIVER, "nlist", str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHIM_DRIVER_OK" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
    raw = open(fout, "rb").read()
    off = 0

    def take(n, dt):
        nonlocal off
        a = np.frombuffer(raw, dtype=dt, count=n, offset=off)
        off += n * np.dtype(dt).itemsize
        return a

    nall, ret_small, ret_cap, max_nnei, mem_nnei, nf = take(6, np.int32)
    assert ret_small == 1 and ret_cap == 1 and nf == nframes and mem_nnei == nall
    cn = take(nloc * 3, np.float64).reshape(nloc, 3)
    ext_c = take(nall * 3, np.float64).reshape(nall, 3)
    ext_t, ext_m = take(nall, np.int32), take(nall, np.int32)
    ilist, numneigh = take(nframes * nloc, np.int32), take(nframes * nloc, np.int32)
    rows = [take(int(k), np.int32) for k in numneigh]
    w = olib.normalize_coord(coord, box)
    assert np.array_equal(cn, w)
    wc, wt, wm = olib.copy_coord(w, atype, box, rc)
    assert nall == len(wt)
    assert np.array_equal(ext_c[:nloc], w) and np.array_equal(ext_m[:nloc], np.arange(nloc))

    def key(c, t, m):
        return sorted((int(m[i]), int(t[i])) + tuple(np.round(c[i], 6)) for i in range(nloc, len(t)))

    assert key(ext_c, ext_t, ext_m) == key(wc, wt, wm)
    wn, wr = olib.build_nlist(ext_c, nloc, rc, atype=ext_t)
    assert max_nnei == wn.max()
    for fr in range(nframes):
        assert np.array_equal(ilist[fr * nloc:(fr + 1) * nloc], np.arange(nloc))
        assert np.array_equal(numneigh[fr * nloc:(fr + 1) * nloc], wn)
        for i in range(nloc):
            assert np.array_equal(rows[fr * nloc + i], wr[i, : wn[i]])
