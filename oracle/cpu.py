"""TEST INFRASTRUCTURE ONLY — ctypes front-end for the two CPU checkers.

* ``CpuLib("port")``      -> oracle/libdp_oracle.so   (our C restatement, dp_oracle.c)
* ``CpuLib("reference")`` -> oracle/_ref/libdeepmd_ref.so (the unmodified reference CPU
  library compiled from /root/reference by oracle/Makefile + the glue in ref_shim.cc)

Both expose the same numpy-in / numpy-out methods so a test can run one against the other
and either against the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": os.path.join(_HERE, "libdp_oracle.so"),
    "reference": os.path.join(_HERE, "_ref", "libdeepmd_ref.so"),
}


def build(kind: str = "all", reference_root: str = "/root/reference") -> None:
    """Compile the checker libraries (the reference one only where its sources exist)."""
    if kind in ("all", "port"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if kind in ("all", "reference") and os.path.isdir(os.path.join(reference_root, "source", "lib", "src")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={reference_root}"])


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fp(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(dtype)


def to_csr(rows):
    """list of per-atom index lists -> (offsets int64[n+1], neigh int32[nnz])."""
    off = np.zeros(len(rows) + 1, dtype=np.int64)
    for i, r in enumerate(rows):
        off[i + 1] = off[i] + len(r)
    neigh = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
    for i, r in enumerate(rows):
        neigh[off[i]:off[i + 1]] = r
    return off, neigh


def dense_to_csr(rows2d, numneigh):
    rows2d = np.asarray(rows2d)
    numneigh = np.asarray(numneigh)
    off = np.zeros(len(numneigh) + 1, dtype=np.int64)
    np.cumsum(numneigh, out=off[1:])
    mask = np.arange(rows2d.shape[1])[None, :] < numneigh[:, None]
    neigh = np.ascontiguousarray(rows2d[mask], dtype=np.int32)
    if neigh.size == 0:
        neigh = np.zeros(1, dtype=np.int32)
    return off, neigh


class CpuLib:
    def __init__(self, kind: str = "port"):
        if kind not in _PATHS:
            raise ValueError(kind)
        if not available(kind):
            raise FileNotFoundError(
                f"{_PATHS[kind]} missing - run `make -C oracle` (the reference flavour needs /root/reference)")
        self.kind = kind
        self.lib = C.CDLL(_PATHS[kind])
        self.pre = "dpo_" if kind == "port" else "ref_"
        if kind == "reference":
            self.lib.ref_last_error.restype = C.c_char_p
            self.lib.ref_legacy_copy_and_build.restype = C.c_void_p
            self.lib.ref_legacy_nnz.restype = C.c_int64

    # -- helpers -------------------------------------------------------------------
    def _call(self, name, *args):
        fn = getattr(self.lib, self.pre + name)
        fn.restype = C.c_int
        rc = fn(*args)
        if rc < 0:
            msg = self.lib.ref_last_error().decode() if self.kind == "reference" else ""
            raise RuntimeError(f"{self.pre}{name} failed: {msg}")
        return rc

    # -- a5 ------------------------------------------------------------------------
    def format_nlist(self, coord, atype, offsets, neigh, rcut, sec, ilist=None):
        coord = np.ascontiguousarray(coord)
        s = _fp(coord.dtype)
        atype = np.ascontiguousarray(atype, dtype=np.int32)
        sec = np.ascontiguousarray(sec, dtype=np.int32)
        inum = len(offsets) - 1
        nall = atype.shape[0]
        nnei = int(sec[-1])
        nlist = np.full((inum, nnei), -1, dtype=np.int32)
        over = np.full(inum, -1, dtype=np.int32)
        il = None if ilist is None else np.ascontiguousarray(ilist, dtype=np.int32)
        args = [_p(nlist), _p(over), _p(coord), _p(atype), C.c_int(inum), _p(il), _p(offsets), _p(neigh)]
        if self.kind == "reference":
            args.append(C.c_int(nall))
        args += [C.c_float(rcut), _p(sec), C.c_int(len(sec))]
        self._call("format_nlist_" + s, *args)
        return nlist, over

    # -- a6 ------------------------------------------------------------------------
    def env_mat_a(self, coord, atype, fmt_nlist, rcut_smth, rcut, sec):
        coord = np.ascontiguousarray(coord)
        s = _fp(coord.dtype)
        atype = np.ascontiguousarray(atype, dtype=np.int32)
        sec = np.ascontiguousarray(sec, dtype=np.int32)
        fmt_nlist = np.ascontiguousarray(fmt_nlist, dtype=np.int32)
        nloc, nnei = fmt_nlist.shape
        em = np.zeros((nloc, nnei * 4), coord.dtype)
        dv = np.zeros((nloc, nnei * 12), coord.dtype)
        rij = np.zeros((nloc, nnei * 3), coord.dtype)
        if self.kind == "reference":
            self._call("env_mat_a_" + s, _p(em), _p(dv), _p(rij), _p(coord), _p(atype), _p(fmt_nlist),
                       C.c_int(nloc), C.c_int(atype.shape[0]), C.c_float(rcut_smth), C.c_float(rcut),
                       _p(sec), C.c_int(len(sec)))
        else:
            self._call("env_mat_a_" + s, _p(em), _p(dv), _p(rij), _p(coord), _p(fmt_nlist), C.c_int(nloc),
                       C.c_float(rcut_smth), C.c_float(rcut), _p(sec), C.c_int(len(sec)))
        return em, dv, rij

    # -- a7 ------------------------------------------------------------------------
    def prod_env_mat_a(self, coord, atype, offsets, neigh, avg, std, nloc, rcut, rcut_smth, sec,
                       ilist=None, f_type=None):
        coord = np.ascontiguousarray(coord)
        s = _fp(coord.dtype)
        atype = np.ascontiguousarray(atype, dtype=np.int32)
        sec = np.ascontiguousarray(sec, dtype=np.int32)
        avg = np.ascontiguousarray(avg, dtype=coord.dtype)
        std = np.ascontiguousarray(std, dtype=coord.dtype)
        nall = atype.shape[0]
        nnei = int(sec[-1])
        inum = len(offsets) - 1
        em = np.zeros((nloc, nnei * 4), coord.dtype)
        dv = np.zeros((nloc, nnei * 12), coord.dtype)
        rij = np.zeros((nloc, nnei * 3), coord.dtype)
        nlist = np.full((nloc, nnei), -1, dtype=np.int32)
        il = None if ilist is None else np.ascontiguousarray(ilist, dtype=np.int32)
        ft = None if f_type is None else np.ascontiguousarray(f_type, dtype=np.int32)
        args = [_p(em), _p(dv), _p(rij), _p(nlist), _p(coord), _p(atype), _p(ft), C.c_int(inum), _p(il),
                _p(offsets), _p(neigh), _p(avg), _p(std), C.c_int(nloc)]
        if self.kind == "reference":
            args += [C.c_int(nall), C.c_int(1)]
        args += [C.c_float(rcut), C.c_float(rcut_smth), _p(sec), C.c_int(len(sec))]
        self._call("prod_env_mat_a_" + s, *args)
        return em, dv, rij, nlist

    # -- a8 / a9 -------------------------------------------------------------------
    def tabulate_fusion_se_a(self, table, info, em_x, em, M, two_embed=None, is_sorted=True):
        table = np.ascontiguousarray(table)
        dt = table.dtype
        s = _fp(dt)
        info = np.ascontiguousarray(info, dtype=dt)
        em_x = np.ascontiguousarray(em_x, dtype=dt)
        em = np.ascontiguousarray(em, dtype=dt)
        nloc, nnei = em.shape[0], em.shape[1]
        nd = em.shape[2] if em.ndim == 3 else 4  # NDESCRPT (tabulate.cc:456-497: 4, 9, 16 or 25)
        te = None if two_embed is None else np.ascontiguousarray(two_embed, dtype=dt)
        out = np.zeros((nloc, nd, M), dt)
        self._call("tabulate_fusion_se_a_nd_" + s, _p(out), _p(table), _p(info), _p(em_x), _p(em), _p(te),
                   C.c_int(nloc), C.c_int(nnei), C.c_int(M), C.c_int(int(is_sorted)), C.c_int(nd))
        return out

    def tabulate_fusion_se_a_grad(self, table, info, em_x, em, dy, M, two_embed=None, is_sorted=True):
        table = np.ascontiguousarray(table)
        dt = table.dtype
        s = _fp(dt)
        info = np.ascontiguousarray(info, dtype=dt)
        em_x = np.ascontiguousarray(em_x, dtype=dt)
        em = np.ascontiguousarray(em, dtype=dt)
        dy = np.ascontiguousarray(dy, dtype=dt)
        nloc, nnei = em.shape[0], em.shape[1]
        nd = em.shape[2] if em.ndim == 3 else 4
        te = None if two_embed is None else np.ascontiguousarray(two_embed, dtype=dt)
        g_x = np.zeros((nloc, nnei), dt)
        g_em = np.zeros((nloc, nnei, nd), dt)
        g_two = np.zeros((nloc, nnei, M), dt) if te is not None else None
        self._call("tabulate_fusion_se_a_grad_nd_" + s, _p(g_x), _p(g_em), _p(g_two), _p(table), _p(info),
                   _p(em_x), _p(em), _p(te), _p(dy), C.c_int(nloc), C.c_int(nnei), C.c_int(M),
                   C.c_int(int(is_sorted)), C.c_int(nd))
        return g_x, g_em, g_two

    def tabulate_fusion_se_a_grad_grad(self, table, info, em_x, em, dz_dem_x, dz_dem, M, two_embed=None,
                                       dz_dtwo=None, is_sorted=True):
        table = np.ascontiguousarray(table)
        dt = table.dtype
        s = _fp(dt)
        info = np.ascontiguousarray(info, dtype=dt)
        em_x = np.ascontiguousarray(em_x, dtype=dt)
        em = np.ascontiguousarray(em, dtype=dt)
        dz_dem_x = np.ascontiguousarray(dz_dem_x, dtype=dt)
        dz_dem = np.ascontiguousarray(dz_dem, dtype=dt)
        nloc, nnei = em.shape[0], em.shape[1]
        nd = em.shape[2] if em.ndim == 3 else 4
        te = None if two_embed is None else np.ascontiguousarray(two_embed, dtype=dt)
        dzt = None if dz_dtwo is None else np.ascontiguousarray(dz_dtwo, dtype=dt)
        out = np.zeros((nloc, nd, M), dt)
        self._call("tabulate_fusion_se_a_grad_grad_nd_" + s, _p(out), _p(table), _p(info), _p(em_x), _p(em),
                   _p(te), _p(dz_dem_x), _p(dz_dem), _p(dzt), C.c_int(nloc), C.c_int(nnei), C.c_int(M),
                   C.c_int(int(is_sorted)), C.c_int(nd))
        return out

    # -- a11 / a12 -----------------------------------------------------------------
    def prod_force_a(self, net_deriv, env_deriv, nlist, nall):
        net_deriv = np.ascontiguousarray(net_deriv)
        dt = net_deriv.dtype
        s = _fp(dt)
        env_deriv = np.ascontiguousarray(env_deriv, dtype=dt)
        nlist = np.ascontiguousarray(nlist, dtype=np.int32)
        nloc, nnei = nlist.shape
        force = np.zeros((nall, 3), dt)
        args = [_p(force), _p(net_deriv), _p(env_deriv), _p(nlist), C.c_int(nloc), C.c_int(nall), C.c_int(nnei)]
        if self.kind == "reference":
            args.append(C.c_int(1))
        self._call("prod_force_a_" + s, *args)
        return force

    def prod_virial_a(self, net_deriv, env_deriv, rij, nlist, nall):
        net_deriv = np.ascontiguousarray(net_deriv)
        dt = net_deriv.dtype
        s = _fp(dt)
        env_deriv = np.ascontiguousarray(env_deriv, dtype=dt)
        rij = np.ascontiguousarray(rij, dtype=dt)
        nlist = np.ascontiguousarray(nlist, dtype=np.int32)
        nloc, nnei = nlist.shape
        virial = np.zeros(9, dt)
        atom_virial = np.zeros((nall, 9), dt)
        self._call("prod_virial_a_" + s, _p(virial), _p(atom_virial), _p(net_deriv), _p(env_deriv), _p(rij),
                   _p(nlist), C.c_int(nloc), C.c_int(nall), C.c_int(nnei))
        return virial, atom_virial

    # -- f2: gradients of a11 / a12 w.r.t. net_deriv -----------------------------------
    def prod_force_grad_a(self, grad, env_deriv, nlist, nframes=1):
        """grad [nframes*nloc, 3] -> grad_net [nframes*nloc, nnei*4]; nlist [nframes*nloc, nnei]."""
        grad = np.ascontiguousarray(grad)
        dt = grad.dtype
        s = _fp(dt)
        env_deriv = np.ascontiguousarray(env_deriv, dtype=dt)
        nlist = np.ascontiguousarray(nlist, dtype=np.int32)
        nrow, nnei = nlist.shape
        nloc = nrow // nframes
        out = np.zeros((nrow, nnei * 4), dt)
        self._call("prod_force_grad_a_" + s, _p(out), _p(grad), _p(env_deriv), _p(nlist), C.c_int(nloc), C.c_int(nnei),
                   C.c_int(nframes))
        return out

    def prod_virial_grad_a(self, grad, env_deriv, rij, nlist):
        """grad [9] -> grad_net [nloc, nnei*4]."""
        grad = np.ascontiguousarray(grad)
        dt = grad.dtype
        s = _fp(dt)
        env_deriv = np.ascontiguousarray(env_deriv, dtype=dt)
        rij = np.ascontiguousarray(rij, dtype=dt)
        nlist = np.ascontiguousarray(nlist, dtype=np.int32)
        nloc, nnei = nlist.shape
        out = np.zeros((nloc, nnei * 4), dt)
        self._call("prod_virial_grad_a_" + s, _p(out), _p(grad), _p(env_deriv), _p(rij), _p(nlist), C.c_int(nloc),
                   C.c_int(nnei))
        return out

    # -- a2 / a3 / a4 --------------------------------------------------------------
    def normalize_coord(self, coord, box):
        coord = np.array(coord, copy=True, order="C")
        s = _fp(coord.dtype)
        box = np.ascontiguousarray(box, dtype=coord.dtype).reshape(9)
        self._call("normalize_coord_" + s, _p(coord), C.c_int(coord.size // 3), _p(box))
        return coord

    def copy_coord(self, coord, atype, box, rcut, mem_nall=None):
        coord = np.ascontiguousarray(coord)
        s = _fp(coord.dtype)
        atype = np.ascontiguousarray(atype, dtype=np.int32)
        box = np.ascontiguousarray(box, dtype=coord.dtype).reshape(9)
        nloc = atype.shape[0]
        mem = mem_nall if mem_nall is not None else max(64, nloc * 4)
        while True:
            out_c = np.zeros((mem, 3), coord.dtype)
            out_t = np.zeros(mem, np.int32)
            mapping = np.zeros(mem, np.int32)
            nall = C.c_int(0)
            rc = self._call("copy_coord_" + s, _p(out_c), _p(out_t), _p(mapping), C.byref(nall), _p(coord),
                            _p(atype), C.c_int(nloc), C.c_int(mem), C.c_float(rcut), _p(box))
            if rc == 0:
                n = nall.value
                return out_c[:n].copy(), out_t[:n].copy(), mapping[:n].copy()
            if mem_nall is not None:
                return None, None, nall.value
            mem = max(nall.value, mem * 2)

    def build_nlist(self, coord, nloc, rcut, atype=None, mem_size=None):
        """Raw list with build_nlist_cpu semantics: returns (numneigh[nloc], rows[nloc, mem])."""
        coord = np.ascontiguousarray(coord)
        s = _fp(coord.dtype)
        nall = coord.size // 3
        ty = None if atype is None else np.ascontiguousarray(atype, dtype=np.int32)
        mem = mem_size if mem_size is not None else 64
        while True:
            numneigh = np.zeros(nloc, np.int32)
            rows = np.zeros((nloc, mem), np.int32)
            mx = C.c_int(0)
            name = "build_nlist_cpu_" if self.kind == "reference" else "build_nlist_"
            rc = self._call(name + s, _p(numneigh), _p(rows), C.byref(mx), _p(coord), C.c_int(nloc),
                            C.c_int(nall), C.c_int(mem), C.c_float(rcut), _p(ty))
            if rc == 0:
                return numneigh, rows
            if mem_size is not None:
                return None, mx.value
            mem = max(mx.value, mem * 2)

    def compute_cell_info(self, box, rcut):
        box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        ci = np.zeros(23, np.int32)
        self._call("compute_cell_info", _p(ci), C.c_float(rcut), _p(box))
        return ci

    # -- reference-only fixture generator (legacy copy_coord + cell-list build_nlist) ---
    def legacy_copy_and_build(self, posi, atype, box, rc_copy, rc_list):
        if self.kind != "reference":
            raise RuntimeError("legacy fixture generator exists only in the reference flavour")
        posi = np.ascontiguousarray(posi, dtype=np.float64)
        atype = np.ascontiguousarray(atype, dtype=np.int32)
        box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        nloc = atype.shape[0]
        h = self.lib.ref_legacy_copy_and_build(_p(posi), _p(atype), C.c_int(nloc), _p(box),
                                               C.c_double(rc_copy), C.c_double(rc_list))
        if not h:
            raise RuntimeError(self.lib.ref_last_error().decode())
        h = C.c_void_p(h)
        try:
            nall = self.lib.ref_legacy_nall(h)
            nnz = self.lib.ref_legacy_nnz(h)
            posi_cpy = np.zeros((nall, 3), np.float64)
            atype_cpy = np.zeros(nall, np.int32)
            mapping = np.zeros(nall, np.int32)
            ncell = np.zeros(3, np.int32)
            ngcell = np.zeros(3, np.int32)
            offsets = np.zeros(nloc + 1, np.int64)
            neigh = np.zeros(max(nnz, 1), np.int32)
            self.lib.ref_legacy_fetch(h, _p(posi_cpy), _p(atype_cpy), _p(mapping), _p(ncell), _p(ngcell),
                                      _p(offsets), _p(neigh))
        finally:
            self.lib.ref_legacy_free(h)
        return dict(coord=posi_cpy, atype=atype_cpy, mapping=mapping, ncell=ncell, ngcell=ngcell,
                    offsets=offsets, neigh=neigh)
