"""Host side of the hot path: tensor-level wrappers over the C ABI and the ``torch.ops.deepmd.*``
operator surface.

Two layers, both thin:

* ``raw`` functions (``prod_env_mat_a``, ``format_nlist``, ``tabulate_fusion_se_a`` ...): take CUDA
  tensors, check shapes/devices, allocate outputs and the workspace and call ``dpb200_*`` on the
  current torch stream.  Argument meaning follows the reference library functions they replace
  (``deepmd::prod_env_mat_a_gpu`` etc., source/lib/include/*.h).
* ``torch.ops.deepmd.*``: the reference's PyTorch operator schemas
  (source/op/pt/tabulate_multi_device.cc:1630-1696) for ``tabulate_fusion_se_a`` /
  ``tabulate_fusion_se_atten`` -- differentiable twice, like the reference -- plus
  ``prod_env_mat_a`` / ``prod_force_se_a`` / ``prod_virial_se_a`` whose argument lists mirror the
  TensorFlow op schemas (source/op/tf/prod_env_mat_multi_device.cc:13-30,
  prod_force_multi_device.cc:6-14, prod_virial_multi_device.cc:5-15).

No CPU implementation exists here: CPU tensors are rejected.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import MAX_NBOR_SIZE, lib

__all__ = [
    "prod_env_mat_a", "format_nlist", "tabulate_fusion_se_a", "tabulate_fusion_se_a_grad",
    "tabulate_fusion_se_a_grad_grad", "prod_force_a", "prod_virial_a", "prod_force_virial_a",
    "normalize_coord", "copy_coord", "build_nlist", "use_nlist_map", "csr_row_pointers",
]


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _suffix(t: torch.Tensor) -> str:
    if t.dtype == torch.float64:
        return "f64"
    if t.dtype == torch.float32:
        return "f32"
    raise TypeError(f"dpb200: unsupported floating type {t.dtype} (float32 / float64 only)")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(*named):
    dev = None
    for name, t in named:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"dpb200: `{name}` is on {t.device}; the hot path has no CPU implementation "
                               "(CUDA tensors required)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"dpb200: `{name}` is on {t.device} but other arguments are on {dev}")
    return dev


def _c(t: torch.Tensor, dtype=None) -> torch.Tensor:
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


def _sec_arr(sec: Sequence[int]):
    sec = [int(s) for s in sec]
    return (C.c_int * len(sec))(*sec), len(sec), sec[-1]


def _workspace(nbytes: int, dev) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def csr_row_pointers(neigh: torch.Tensor, numneigh: torch.Tensor) -> torch.Tensor:
    """Device array of device row pointers (the `firstneigh` of a converted InputNlist,
    source/lib/include/neighbor_list.h:20-57) for rows stored back to back in `neigh`."""
    off = torch.cumsum(numneigh.to(torch.int64), 0) - numneigh.to(torch.int64)
    return off * 4 + neigh.data_ptr()


# ------------------------------------------------------------------------------------------------
# a5-a7: neighbour formatting + environment matrix
# ------------------------------------------------------------------------------------------------
def _rows_args(numneigh, rows, firstneigh, max_nbor_size):
    if firstneigh is not None:
        if max_nbor_size is None:
            max_nbor_size = int(numneigh.max().item()) if numneigh.numel() else 0
        return None, 0, firstneigh, int(max_nbor_size)
    if rows.dim() != 2:
        raise ValueError("dpb200: `rows` must be a dense [nrows, capacity] int32 block")
    cap = rows.shape[1]
    if max_nbor_size is None:
        max_nbor_size = cap
    return rows, cap, None, int(min(max_nbor_size, cap))


def prod_env_mat_a(coord, atype, numneigh, rows, avg, std, nloc, nall, rcut, rcut_smth, sec, *, ilist=None,
                   f_type=None, firstneigh=None, nframes=1, max_nbor_size=None, row_range=None):
    """deepmd::prod_env_mat_a_gpu (source/lib/include/prod_env_mat.h:91-110): returns
    (em[nf*nloc, nnei*4], em_deriv[nf*nloc, nnei*12], rij[nf*nloc, nnei*3], nlist[nf*nloc, nnei]).
    row_range=(a, b) (single frame, dense rows): only centre atoms a..b-1, outputs have b-a rows."""
    if row_range is not None:
        a, b = int(row_range[0]), int(row_range[1])
        if nframes != 1 or rows is None or ilist is not None:
            raise ValueError("dpb200: row_range needs one frame, dense rows and no ilist")
        numneigh = numneigh[a:b]
        rows = rows[a:b]
        ilist = torch.arange(a, b, dtype=torch.int32, device=coord.device)
        nloc_eff, shift = b - a, a
    else:
        nloc_eff, shift = nloc, 0
    dev = _need_cuda(("coord", coord), ("type", atype), ("numneigh", numneigh), ("rows", rows), ("avg", avg),
                     ("std", std), ("ilist", ilist), ("f_type", f_type), ("firstneigh", firstneigh))
    s = _suffix(coord)
    coord = _c(coord)
    atype = _c(atype, torch.int32)
    numneigh = _c(numneigh, torch.int32)
    avg = _c(avg, coord.dtype)
    std = _c(std, coord.dtype)
    if rows is not None:
        rows = _c(rows, torch.int32)
    if ilist is not None:
        ilist = _c(ilist, torch.int32)
    if f_type is not None:
        f_type = _c(f_type, torch.int32)
    sec_c, nsec, nnei = _sec_arr(sec)
    # avg / std carry one row per CENTRE type; with a coarser f_type (se_atten: one distance-ordered section)
    # that is more rows than sections
    ntypes = avg.numel() // (nnei * 4) if nnei else nsec - 1
    if avg.numel() != ntypes * nnei * 4 or std.numel() != avg.numel() or ntypes < 1 or (f_type is None and ntypes != nsec - 1):
        raise ValueError("dpb200: avg/std must have ntypes*nnei*4 elements")
    if coord.numel() != nframes * nall * 3 or atype.numel() != nframes * nall:
        raise ValueError("dpb200: coord/type do not match nframes*nall")
    rows_t, stride, fn, mx = _rows_args(numneigh, rows, firstneigh, max_nbor_size)
    if mx > MAX_NBOR_SIZE:
        raise ValueError(f"dpb200: neighbour rows wider than {MAX_NBOR_SIZE} are not supported")
    n = nframes * nloc_eff
    em = torch.empty((n, nnei * 4), dtype=coord.dtype, device=dev)
    dv = torch.empty((n, nnei * 12), dtype=coord.dtype, device=dev)
    rij = torch.empty((n, nnei * 3), dtype=coord.dtype, device=dev)
    nlist = torch.empty((n, nnei), dtype=torch.int32, device=dev)
    L = lib()
    wsb = L.cdll.dpb200_prod_env_mat_a_workspace_bytes(ntypes, nnei, nall, nframes, coord.element_size())
    ws = _workspace(wsb, dev)
    esz = coord.element_size()
    # the kernel writes row ilist[r] of each output: with a row range the output pointers are moved
    # back by `shift` rows so that centre atom a lands in row 0 of the compact chunk buffers
    po = lambda t, width, b_: C.c_void_p(t.data_ptr() - shift * width * b_)
    L.call("prod_env_mat_a_ex_" + s, po(em, nnei * 4, esz), po(dv, nnei * 12, esz), po(rij, nnei * 3, esz),
           po(nlist, nnei, 4), _p(coord), _p(atype), _p(f_type), _p(ilist), _p(numneigh), _p(fn), _p(rows_t), stride, mx,
           _p(avg), _p(std), ntypes, nloc_eff, nall, nframes, float(rcut), float(rcut_smth), sec_c, nsec, _p(ws), ws.numel(),
           _stream(dev))
    return em, dv, rij, nlist


def format_nlist(coord, atype, numneigh, rows, nloc, nall, rcut, sec, *, ilist=None, firstneigh=None, nframes=1,
                 max_nbor_size=None):
    """deepmd::format_nbor_list_gpu (source/lib/include/fmt_nlist.h:22-34) with the CPU semantics of
    format_nlist_i_cpu (source/lib/src/fmt_nlist.cc:98-143): nlist[nf*nloc, nnei] int32."""
    dev = _need_cuda(("coord", coord), ("type", atype), ("numneigh", numneigh), ("rows", rows), ("ilist", ilist),
                     ("firstneigh", firstneigh))
    s = _suffix(coord)
    coord = _c(coord)
    atype = _c(atype, torch.int32)
    numneigh = _c(numneigh, torch.int32)
    if rows is not None:
        rows = _c(rows, torch.int32)
    if ilist is not None:
        ilist = _c(ilist, torch.int32)
    sec_c, nsec, nnei = _sec_arr(sec)
    rows_t, stride, fn, mx = _rows_args(numneigh, rows, firstneigh, max_nbor_size)
    if mx > MAX_NBOR_SIZE:
        raise ValueError(f"dpb200: neighbour rows wider than {MAX_NBOR_SIZE} are not supported")
    nlist = torch.empty((nframes * nloc, nnei), dtype=torch.int32, device=dev)
    L = lib()
    wsb = L.cdll.dpb200_prod_env_mat_a_workspace_bytes(nsec - 1, nnei, nall, nframes, coord.element_size())
    ws = _workspace(wsb, dev)
    L.call("format_nlist_" + s, _p(nlist), _p(coord), _p(atype), _p(ilist), _p(numneigh), _p(fn), _p(rows_t), stride,
           mx, nloc, nall, nframes, float(rcut), sec_c, nsec, _p(ws), ws.numel(), _stream(dev))
    return nlist


# ------------------------------------------------------------------------------------------------
# a8-a10: compressed embedding table
# ------------------------------------------------------------------------------------------------
def _info_host(table_info: torch.Tensor, dtype) -> torch.Tensor:
    if table_info.device.type != "cpu":
        raise RuntimeError("table_info must be on the CPU")
    ti = table_info.detach().reshape(-1).to(dtype).contiguous()
    if ti.numel() < 5:
        raise ValueError("table_info needs at least 5 entries (lower, upper, max, stride0, stride1)")
    return ti


def _check_tab(table, em_x, em, two_embed):
    if table.dim() != 2:
        raise ValueError("Dim of table should be 2")
    if em_x.dim() != 2:
        raise ValueError("Dim of input should be 2")
    if em.dim() != 3:
        raise ValueError("Dim of input should be 3")
    if two_embed is not None and two_embed.dim() != 2:
        raise ValueError("Dim of input should be 2")
    if em.shape[2] not in (4, 9, 16, 25):  # tabulate.h is_supported_se_a_basis_dimension
        raise ValueError(f"The environment basis dimension must be 4, 9, 16 or 25, got {em.shape[2]}")
    for name, t in (("em_x", em_x), ("em", em), ("two_embed", two_embed)):
        if t is not None and t.device != table.device:
            raise RuntimeError(f"{name} must be on the same device as table; table is on {table.device} but "
                               f"{name} is on {t.device}")


def tabulate_fusion_se_a(table, table_info, em_x, em, last_layer_size, two_embed=None, is_sorted=True):
    """deepmd::tabulate_fusion_se_a_gpu (source/lib/include/tabulate.h:175-186): descriptor [nloc,4,M]."""
    _check_tab(table, em_x, em, two_embed)
    dev = _need_cuda(("table", table), ("em_x", em_x), ("em", em), ("two_embed", two_embed))
    s = _suffix(table)
    ti = _info_host(table_info, table.dtype)
    table, em_x, em = _c(table), _c(em_x, table.dtype), _c(em, table.dtype)
    if two_embed is not None:
        two_embed = _c(two_embed, table.dtype)
    nloc, nnei, nd = em.shape[0], em.shape[1], em.shape[2]
    M = int(last_layer_size)
    out = torch.empty((nloc, nd, M), dtype=table.dtype, device=dev)
    if nd != 4:  # higher angular bases: csrc/tabulate_nd.cu
        lib().call("tabulate_fusion_se_a_nd_" + s, _p(out), _p(table), C.c_void_p(ti.data_ptr()), _p(em_x), _p(em),
                   _p(two_embed), nloc, nnei, M, int(bool(is_sorted)), nd, _stream(dev))
        return out
    lib().call("tabulate_fusion_se_a_" + s, _p(out), _p(table), C.c_void_p(ti.data_ptr()), _p(em_x), _p(em),
               _p(two_embed), nloc, nnei, M, int(bool(is_sorted)), _stream(dev))
    return out


def tabulate_fusion_se_a_grad(table, table_info, em_x, em, dy, last_layer_size, two_embed=None, is_sorted=True):
    """deepmd::tabulate_fusion_se_a_grad_gpu (tabulate.h:188-202): (dy_dem_x, dy_dem, dy_dtwo|None)."""
    _check_tab(table, em_x, em, two_embed)
    dev = _need_cuda(("table", table), ("em_x", em_x), ("em", em), ("two_embed", two_embed), ("dy", dy))
    s = _suffix(table)
    ti = _info_host(table_info, table.dtype)
    table, em_x, em, dy = _c(table), _c(em_x, table.dtype), _c(em, table.dtype), _c(dy, table.dtype)
    if two_embed is not None:
        two_embed = _c(two_embed, table.dtype)
    nloc, nnei = em.shape[0], em.shape[1]
    M = int(last_layer_size)
    g_x = torch.zeros_like(em_x) if nnei == 0 else torch.empty_like(em_x)
    g_em = torch.zeros_like(em) if nnei == 0 else torch.empty_like(em)
    g_two = None if two_embed is None else torch.empty_like(two_embed)
    if em.shape[2] != 4:
        lib().call("tabulate_fusion_se_a_grad_nd_" + s, _p(g_x), _p(g_em), _p(g_two), _p(table),
                   C.c_void_p(ti.data_ptr()), _p(em_x), _p(em), _p(two_embed), _p(dy), nloc, nnei, M,
                   int(bool(is_sorted)), int(em.shape[2]), _stream(dev))
        return g_x, g_em, g_two
    lib().call("tabulate_fusion_se_a_grad_" + s, _p(g_x), _p(g_em), _p(g_two), _p(table), C.c_void_p(ti.data_ptr()),
               _p(em_x), _p(em), _p(two_embed), _p(dy), nloc, nnei, M, int(bool(is_sorted)), _stream(dev))
    return g_x, g_em, g_two


def tabulate_fusion_se_atten_gate(table, table_info, em_x, em, tt_full, pair, sw, last_layer_size, is_sorted=True,
                                  flags=0):
    """tabulate_fusion_se_atten with two_embed = tt_full[pair] * sw formed inside the kernel (never materialised):
    pair int32 [nloc, nnei] rows of tt_full [(ntypes+1)^2, M], sw [nloc, nnei].  Returns [nloc, 4, M]."""
    dev = _need_cuda(("table", table), ("em_x", em_x), ("em", em), ("tt_full", tt_full), ("pair", pair), ("sw", sw))
    s = _suffix(table)
    ti = _info_host(table_info, table.dtype)
    table, em_x, em = _c(table), _c(em_x, table.dtype), _c(em, table.dtype)
    tt_full, pair, sw = _c(tt_full, table.dtype), _c(pair, torch.int32), _c(sw, table.dtype)
    nloc, nnei = em.shape[0], em.shape[1]
    M = int(last_layer_size)
    if tt_full.shape[-1] != M or pair.numel() != nloc * nnei or sw.numel() != nloc * nnei:
        raise ValueError("dpb200: gate tensors do not match [nloc, nnei] / M")
    out = torch.empty((nloc, 4, M), dtype=table.dtype, device=dev)
    lib().call("tabulate_fusion_se_atten_gate_" + s, _p(out), _p(table), C.c_void_p(ti.data_ptr()), _p(em_x), _p(em),
               _p(tt_full), _p(pair), _p(sw), nloc, nnei, M, int(bool(is_sorted)), int(flags), _stream(dev))
    return out


def tabulate_fusion_se_atten_gate_desc(table, table_info, em_x, em, tt_full, pair, sw, last_layer_size, axis, scale,
                                       dim_in, nslice, min_row_exp, pad_rows=0, desc_row=None, is_sorted=True, flags=0):
    """tabulate_fusion_se_atten_gate with the descriptor epilogue (fp64): returns (out [nloc, 4, M], desc int8
    [nloc+pad, nslice*dim_in] whose first M*axis columns of every slice hold D = (scale out)^T (scale out)[:, :axis]
    (the remaining dim_in - M*axis columns are zero, for the caller's own operand columns: fit_slice_cols), row_exp
    int32 [nloc+pad] >= min_row_exp)."""
    dev = _need_cuda(("table", table), ("em_x", em_x), ("em", em), ("tt_full", tt_full), ("pair", pair), ("sw", sw))
    if table.dtype != torch.float64:
        raise ValueError("dpb200: the gated descriptor epilogue is the fp64 form")
    s = _suffix(table)
    ti = _info_host(table_info, table.dtype)
    table, em_x, em = _c(table), _c(em_x, table.dtype), _c(em, table.dtype)
    tt_full, pair, sw = _c(tt_full, table.dtype), _c(pair, torch.int32), _c(sw, table.dtype)
    nloc, nnei = em.shape[0], em.shape[1]
    M = int(last_layer_size)
    out = torch.empty((nloc, 4, M), dtype=table.dtype, device=dev)
    rows = nloc + int(pad_rows)
    desc = torch.zeros((rows, int(nslice) * int(dim_in)), dtype=torch.int8, device=dev)
    row_exp = torch.zeros((rows,), dtype=torch.int32, device=dev)
    if desc_row is not None:
        desc_row = _c(desc_row, torch.int32)
    lib().call("tabulate_fusion_se_atten_gate_desc_" + s, _p(out), _p(table), C.c_void_p(ti.data_ptr()), _p(em_x), _p(em),
               _p(tt_full), _p(pair), _p(sw), nloc, nnei, M, int(bool(is_sorted)), int(axis), float(scale), _p(desc_row),
               2, _p(desc), int(desc.shape[1]), int(dim_in), int(nslice), _p(row_exp), int(min_row_exp), int(flags),
               _stream(dev))
    return out, desc, row_exp


def fit_slice_cols(desc, slice_stride, col0, nslice, row_exp, src, idx=None):
    """Columns [col0, col0 + src.shape[1]) of every digit slice of desc[r] := digits of src[idx[r]] at row_exp[r]
    (dpb200_fit_slice_cols_f64)."""
    dev = _need_cuda(("desc", desc), ("row_exp", row_exp), ("src", src), ("idx", idx))
    src = _c(src, torch.float64)
    n = int(row_exp.numel()) if idx is None else int(idx.numel())
    if idx is not None:
        idx = _c(idx, torch.int32)
    lib().call("fit_slice_cols_f64", _p(desc), int(desc.stride(0)), int(slice_stride), int(col0), int(src.shape[1]),
               int(nslice), _p(row_exp), _p(src), int(src.stride(0)), _p(idx), n, _stream(dev))


def tabulate_fusion_se_atten_gate_grad(table, table_info, em_x, em, tt_full, pair, sw, dy, last_layer_size,
                                       is_sorted=True, fuse_x=False, flags=0):
    """Backward of tabulate_fusion_se_atten_gate: (dy_dem_x [nloc*nnei, 1], dy_dem [nloc, nnei, 4], dy_dsw [nloc, nnei]).
    fuse_x: em_x is component 0 of em, its gradient is added into dy_dem[..., 0] and dy_dem_x is returned as None."""
    dev = _need_cuda(("table", table), ("em_x", em_x), ("em", em), ("tt_full", tt_full), ("pair", pair), ("sw", sw),
                     ("dy", dy))
    s = _suffix(table)
    ti = _info_host(table_info, table.dtype)
    table, em_x, em, dy = _c(table), _c(em_x, table.dtype), _c(em, table.dtype), _c(dy, table.dtype)
    tt_full, pair, sw = _c(tt_full, table.dtype), _c(pair, torch.int32), _c(sw, table.dtype)
    nloc, nnei = em.shape[0], em.shape[1]
    M = int(last_layer_size)
    g_x = None if fuse_x else torch.empty_like(em_x)
    g_em = torch.empty_like(em)
    g_sw = torch.empty((nloc, nnei), dtype=table.dtype, device=dev)
    lib().call("tabulate_fusion_se_atten_gate_grad_" + s, _p(g_x), _p(g_em), _p(g_sw), _p(table),
               C.c_void_p(ti.data_ptr()), _p(em_x), _p(em), _p(tt_full), _p(pair), _p(sw), _p(dy), nloc, nnei, M,
               int(bool(is_sorted)), int(flags), _stream(dev))
    return g_x, g_em, g_sw


def tabulate_fusion_se_a_grad_grad(table, table_info, em_x, em, dz_dy_dem_x, dz_dy_dem, last_layer_size,
                                   two_embed=None, dz_dy_dtwo=None, is_sorted=True):
    """deepmd::tabulate_fusion_se_a_grad_grad_gpu (tabulate.h:204-218): dz_dy [nloc,4,M]."""
    _check_tab(table, em_x, em, two_embed)
    dev = _need_cuda(("table", table), ("em_x", em_x), ("em", em), ("two_embed", two_embed),
                     ("dz_dy_dem_x", dz_dy_dem_x), ("dz_dy_dem", dz_dy_dem), ("dz_dy_dtwo", dz_dy_dtwo))
    s = _suffix(table)
    ti = _info_host(table_info, table.dtype)
    table, em_x, em = _c(table), _c(em_x, table.dtype), _c(em, table.dtype)
    dz_x, dz_em = _c(dz_dy_dem_x, table.dtype), _c(dz_dy_dem, table.dtype)
    if two_embed is not None:
        two_embed = _c(two_embed, table.dtype)
        dz_dy_dtwo = torch.zeros_like(two_embed) if dz_dy_dtwo is None else _c(dz_dy_dtwo, table.dtype)
    nloc, nnei = em.shape[0], em.shape[1]
    M = int(last_layer_size)
    out = torch.empty((nloc, em.shape[2], M), dtype=table.dtype, device=dev)
    if em.shape[2] != 4:
        lib().call("tabulate_fusion_se_a_grad_grad_nd_" + s, _p(out), _p(table), C.c_void_p(ti.data_ptr()), _p(em_x),
                   _p(em), _p(two_embed), _p(dz_x), _p(dz_em), _p(dz_dy_dtwo if two_embed is not None else None),
                   nloc, nnei, M, int(bool(is_sorted)), int(em.shape[2]), _stream(dev))
        return out
    lib().call("tabulate_fusion_se_a_grad_grad_" + s, _p(out), _p(table), C.c_void_p(ti.data_ptr()), _p(em_x), _p(em),
               _p(two_embed), _p(dz_x), _p(dz_em), _p(dz_dy_dtwo if two_embed is not None else None), nloc, nnei, M,
               int(bool(is_sorted)), _stream(dev))
    return out


def tabulate_sections_fwd(tables, infos, em, sec, last_layer_size, is_sorted=True):
    """All type sections of one env-mat in place (no slicing copies): sum_t tabulate(em[:, sec_t]) ->
    [nloc,4,M].  `em` is the full [nloc, nnei*4] matrix of prod_env_mat_a; em_x is its component 0.
    Mirrors the loop of deepmd/pt/model/descriptor/se_a.py:810-841 (type_one_side)."""
    dev = _need_cuda(("em", em))
    s = _suffix(em)
    em = _c(em)
    nloc = em.shape[0]
    nnei = int(sec[-1])
    M = int(last_layer_size)
    out = torch.empty((nloc, 4, M), dtype=em.dtype, device=dev)
    esz = em.element_size()
    first = True
    for t, (table, info) in enumerate(zip(tables, infos)):
        n_t = int(sec[t + 1] - sec[t])
        if n_t == 0:
            continue
        ti = _info_host(info, em.dtype)
        base = em.data_ptr() + int(sec[t]) * 4 * esz
        lib().call("tabulate_fusion_se_a_ex_" + s, _p(out), _p(_c(table)), C.c_void_p(ti.data_ptr()),
                   C.c_void_p(base), nnei * 4, 4, C.c_void_p(base), nnei * 4, None, nloc, n_t, M,
                   int(bool(is_sorted)), 0 if first else 1, _stream(dev))
        first = False
    if first:
        out.zero_()
    return out


def compressed_coef_flags_f32(table: torch.Tensor, info, tol_v: float = 1e-7, tol_d: float = 2e-6) -> int:
    """fp32 flavour of the gate: the kernels keep {a0..a3} only (one float4 per row and channel).  Allowed when
    dropping a4 x^4 + a5 x^5 changes the quintic by less than tol_v * max|a0| (about one fp32 ulp) and its derivative
    by less than tol_d * max|a1| on the stride-0 rows; ten times those bounds (still within the 1e-5 fp32 parity
    tolerance) on the stride-1 rows, which only inputs outside the tabulated range reach."""
    t = table.detach().to("cpu", torch.float64)
    lower, upper, vmax, s0, s1 = [float(x) for x in info[:5]]
    nrow = t.shape[0]
    a = t.reshape(nrow, -1, 6).abs()
    first = min(nrow, int((upper - lower) / s0))
    m0, m1 = float(a[..., 0].max()), float(a[..., 1].max())
    if not (m0 > 0 and m1 > 0):
        return 0
    for blk, s, f in ((a[:first], s0, 1.0), (a[first:], s1, 10.0)):
        if blk.numel() == 0:
            continue
        ev = blk[..., 4] * s ** 4 + blk[..., 5] * s ** 5
        ed = 4 * blk[..., 4] * s ** 3 + 5 * blk[..., 5] * s ** 4
        if float(ev.max()) > f * tol_v * m0 or float(ed.max()) > f * tol_d * m1:
            return 0
    return 1


def compressed_coef_flags(table: torch.Tensor, info, tol: float = 3e-12) -> int:
    """Host-side gate for DPB200_TAB_COMPRESSED_COEF (include/dpb200.h): returns the flags word for THIS table
    ([nrow, M*6] fp64) if, on every stride-0 row (the stride-1 extrapolation rows are never compressed), storing
    a3, a4 as fp32, a5 as fp16 (scaled by a power of two) and a2 with 36 mantissa bits changes the quintic by less
    than tol * max|a0| and its derivative by less than tol * max|a1|; else 0.
    Rounding model: a2 2^-37; a3, a4 2^-25 storage + 2^-24 fp32 Horner step; a5 2^-11."""
    t = table.detach().to("cpu", torch.float64)
    lower, upper, vmax, s0, s1 = [float(x) for x in info[:5]]
    nrow = t.shape[0]
    a = t.reshape(nrow, -1, 6).abs()
    first = min(nrow, int((upper - lower) / s0))
    m0, m1 = float(a[..., 0].max()), float(a[..., 1].max())
    if not (m0 > 0 and m1 > 0 and first > 0):
        return 0
    e2, e34, e5 = 2.0 ** -37, 1.5 * 2.0 ** -24, 2.0 ** -11 + 2.0 ** -23
    blk, s = a[:first], s0
    ev = e2 * blk[..., 2] * s ** 2 + e34 * (blk[..., 3] * s ** 3 + blk[..., 4] * s ** 4) + e5 * blk[..., 5] * s ** 5
    ed = 2 * e2 * blk[..., 2] * s + e34 * (3 * blk[..., 3] * s ** 2 + 4 * blk[..., 4] * s ** 3) + e5 * 5 * blk[..., 5] * s ** 4
    if float(ev.max()) > tol * m0 or float(ed.max()) > tol * m1:
        return 0
    # the forward (value only) reads the compressed form on the stride-1 rows as well
    if first < nrow:
        c = a[first:]
        evc = e2 * c[..., 2] * s1 ** 2 + e34 * (c[..., 3] * s1 ** 3 + c[..., 4] * s1 ** 4) + e5 * c[..., 5] * s1 ** 5
        if float(evc.max()) > tol * m0:
            return 0
    m5 = float(blk[..., 5].max())
    k = 0
    if m5 > 0:
        import math

        k = max(-120, min(120, 13 - int(math.floor(math.log2(m5)))))  # max|a5| * 2^k in [2^13, 2^14) on stride-0 rows
    if float(a[..., 5].max()) * 2.0 ** k >= 6.0e4:  # fp16 range on every row
        return 0
    return 1 | ((k & 0xff) << 8)


def tabulate_sections_grad(tables, infos, em, dy, sec, last_layer_size, is_sorted=True, flags=None):
    """Backward of tabulate_sections_fwd: d/d(em) as one [nloc, nnei*4] matrix (the em_x gradient is
    added into component 0 inside the kernel, which is what autograd does with the
    `ss = rr[..., :1]` slice of deepmd/pt/model/descriptor/se_a.py:818-820)."""
    dev = _need_cuda(("em", em), ("dy", dy))
    s = _suffix(em)
    em, dy = _c(em), _c(dy, em.dtype)
    nloc = em.shape[0]
    nnei = int(sec[-1])
    M = int(last_layer_size)
    g_em = torch.empty_like(em)
    esz = em.element_size()
    for t, (table, info) in enumerate(zip(tables, infos)):
        n_t = int(sec[t + 1] - sec[t])
        if n_t == 0:
            continue
        ti = _info_host(info, em.dtype)
        off = int(sec[t])
        base = em.data_ptr() + off * 4 * esz
        lib().call("tabulate_fusion_se_a_grad_fx_" + s, None,
                   C.c_void_p(g_em.data_ptr() + off * 4 * esz), _p(_c(table)), C.c_void_p(ti.data_ptr()),
                   C.c_void_p(base), nnei * 4, 4, C.c_void_p(base), nnei * 4, _p(dy), nloc, n_t, M,
                   int(bool(is_sorted)), 0 if flags is None else int(flags[t]), _stream(dev))
    return g_em


# ------------------------------------------------------------------------------------------------
# a11-a12: force / virial
# ------------------------------------------------------------------------------------------------
def prod_force_a(net_deriv, in_deriv, nlist, nloc, nall, nnei, nframes=1):
    """deepmd::prod_force_a_gpu (source/lib/include/prod_force.h:71-79): force [nframes, nall*3]."""
    dev = _need_cuda(("net_deriv", net_deriv), ("in_deriv", in_deriv), ("nlist", nlist))
    s = _suffix(net_deriv)
    net_deriv, in_deriv, nlist = _c(net_deriv), _c(in_deriv, net_deriv.dtype), _c(nlist, torch.int32)
    if net_deriv.numel() != nframes * nloc * nnei * 4 or in_deriv.numel() != nframes * nloc * nnei * 12:
        raise ValueError("dpb200: net_deriv / in_deriv do not match nframes*nloc*nnei")
    force = torch.empty((nframes, nall * 3), dtype=net_deriv.dtype, device=dev)
    lib().call("prod_force_a_" + s, _p(force), _p(net_deriv), _p(in_deriv), _p(nlist), nloc, nall, nnei, nframes,
               _stream(dev))
    return force


def prod_virial_a(net_deriv, in_deriv, rij, nlist, nloc, nall, nnei):
    """deepmd::prod_virial_a_gpu (source/lib/include/prod_virial.h:30-39): (virial[9], atom_virial[nall*9])."""
    dev = _need_cuda(("net_deriv", net_deriv), ("in_deriv", in_deriv), ("rij", rij), ("nlist", nlist))
    s = _suffix(net_deriv)
    dt = net_deriv.dtype
    net_deriv, in_deriv, rij, nlist = _c(net_deriv), _c(in_deriv, dt), _c(rij, dt), _c(nlist, torch.int32)
    virial = torch.empty(9, dtype=dt, device=dev)
    atom_virial = torch.empty(nall * 9, dtype=dt, device=dev)
    lib().call("prod_virial_a_" + s, _p(virial), _p(atom_virial), _p(net_deriv), _p(in_deriv), _p(rij), _p(nlist),
               nloc, nall, nnei, _stream(dev))
    return virial, atom_virial


def prod_force_virial_a(net_deriv, in_deriv, rij, nlist, nloc, nall, nnei, atom_virial=False):
    """Fused single pass over net_deriv / in_deriv: (force[nall*3], virial[9], atom_virial|None)."""
    dev = _need_cuda(("net_deriv", net_deriv), ("in_deriv", in_deriv), ("rij", rij), ("nlist", nlist))
    s = _suffix(net_deriv)
    dt = net_deriv.dtype
    net_deriv, in_deriv, rij, nlist = _c(net_deriv), _c(in_deriv, dt), _c(rij, dt), _c(nlist, torch.int32)
    force = torch.empty(nall * 3, dtype=dt, device=dev)
    virial = torch.empty(9, dtype=dt, device=dev)
    av = torch.empty(nall * 9, dtype=dt, device=dev) if atom_virial else None
    lib().call("prod_force_virial_a_" + s, _p(force), _p(virial), _p(av), _p(net_deriv), _p(in_deriv), _p(rij),
               _p(nlist), nloc, nall, nnei, _stream(dev))
    return force, virial, av


def prod_force_virial_a_pair(net_deriv, in_deriv, rij, nlist, pair_q, pair_w, nloc, nall, nnei, atom_virial=False):
    """prod_force_virial_a plus the central pair force -(pair_q * pair_w) * rij of every slot (se_atten switch path;
    dpb200_prod_force_virial_a_pair): (force[nall*3], virial[9], atom_virial|None)."""
    dev = _need_cuda(("net_deriv", net_deriv), ("in_deriv", in_deriv), ("rij", rij), ("nlist", nlist),
                     ("pair_q", pair_q), ("pair_w", pair_w))
    s = _suffix(net_deriv)
    dt = net_deriv.dtype
    net_deriv, in_deriv, rij, nlist = _c(net_deriv), _c(in_deriv, dt), _c(rij, dt), _c(nlist, torch.int32)
    pair_q, pair_w = _c(pair_q, dt), _c(pair_w, dt)
    force = torch.empty(nall * 3, dtype=dt, device=dev)
    virial = torch.empty(9, dtype=dt, device=dev)
    av = torch.empty(nall * 9, dtype=dt, device=dev) if atom_virial else None
    lib().call("prod_force_virial_a_pair_" + s, _p(force), _p(virial), _p(av), _p(net_deriv), _p(in_deriv), _p(rij),
               _p(nlist), _p(pair_q), _p(pair_w), nloc, nall, nnei, _stream(dev))
    return force, virial, av


def se_atten_gate_scalars(nlist, ext_type, rij, nloc, nnei, ntypes, rcut_smth, rcut):
    """Per (centre, slot) of the formatted list (extended indices): (pair int32 = row of tt_full, sw, sw'(r)/r), all
    [nloc, nnei] (dpb200_se_atten_gate_scalars; se_atten.py:916-926, 979-983)."""
    dev = _need_cuda(("nlist", nlist), ("ext_type", ext_type), ("rij", rij))
    s = _suffix(rij)
    nlist, ext_type, rij = _c(nlist, torch.int32), _c(ext_type, torch.int32), _c(rij)
    pair = torch.empty((nloc, nnei), dtype=torch.int32, device=dev)
    sw = torch.empty((nloc, nnei), dtype=rij.dtype, device=dev)
    dswr = torch.empty((nloc, nnei), dtype=rij.dtype, device=dev)
    lib().call("se_atten_gate_scalars_" + s, _p(pair), _p(sw), _p(dswr), _p(nlist), _p(ext_type), _p(rij), int(nloc),
               int(nnei), int(ntypes), float(rcut_smth), float(rcut), _stream(dev))
    return pair, sw, dswr


# ---- DPA-1 attention layers (csrc/attn_layers.cu): the stages between the dense products of a layer ------------
def se_atten_embed(table, table_info, em, tt_full, pair, sw, last_layer_size):
    """em [natoms, nnei, 4] -> (x0, g_s, g_s') [natoms, nnei, M]: the tabulated geometric embedding of every neighbour
    and x0 = g_s (1 + tt_full[pair] sw) (se_atten.py:979-1003)."""
    dev = _need_cuda(("table", table), ("em", em), ("tt_full", tt_full), ("pair", pair), ("sw", sw))
    s = _suffix(em)
    table, em, tt_full, pair, sw = _c(table, em.dtype), _c(em), _c(tt_full, em.dtype), _c(pair, torch.int32), _c(sw, em.dtype)
    info = _info_host(table_info, em.dtype)
    n, nnei = em.shape[0], em.shape[1]
    M = int(last_layer_size)
    x0 = torch.empty((n, nnei, M), dtype=em.dtype, device=dev)
    gs, dgs = torch.empty_like(x0), torch.empty_like(x0)
    lib().call("se_atten_embed_" + s, _p(x0), _p(gs), _p(dgs), _p(table), info.data_ptr(), _p(em), 4, _p(tt_full),
               _p(pair), _p(sw), n * nnei, M, _stream(dev))
    return x0, gs, dgs


def se_atten_embed_grad(d_em, d_sw, dx0, gs, dgs, tt_full, pair, sw):
    """Accumulates dE/ds into d_em[:, :, 0] and dE/d(sw) into d_sw (both in place)."""
    dev = _need_cuda(("d_em", d_em), ("dx0", dx0))
    s = _suffix(dx0)
    assert d_em.is_contiguous() and d_sw.is_contiguous()
    n, nnei, M = dx0.shape
    lib().call("se_atten_embed_grad_" + s, _p(d_em), 4, _p(d_sw), _p(_c(dx0)), _p(gs), _p(dgs), _p(_c(tt_full, dx0.dtype)),
               _p(_c(pair, torch.int32)), _p(_c(sw, dx0.dtype)), n * nnei, M, _stream(dev))


def se_atten_rhat(em):
    """em [natoms, nnei, 4] -> (rhat [natoms, nnei, 3] = normalize(em[..., 1:4]), rinv [natoms, nnei])."""
    dev = _need_cuda(("em", em))
    s = _suffix(em)
    em = _c(em)
    n, nnei = em.shape[0], em.shape[1]
    rhat = torch.empty((n, nnei, 3), dtype=em.dtype, device=dev)
    rinv = torch.empty((n, nnei), dtype=em.dtype, device=dev)
    lib().call("se_atten_rhat_" + s, _p(rhat), _p(rinv), _p(em), n * nnei, _stream(dev))
    return rhat, rinv


def se_atten_rhat_grad(d_em, d_rhat, rhat, rinv):
    """d_em[..., 1:4] += d normalize / d em applied to d_rhat (in place)."""
    dev = _need_cuda(("d_em", d_em), ("d_rhat", d_rhat))
    assert d_em.is_contiguous()
    lib().call("se_atten_rhat_grad_" + _suffix(d_em), _p(d_em), _p(_c(d_rhat)), _p(rhat), _p(rinv), rinv.numel(),
               _stream(dev))


def attn_qkv_normalize(qkv, hidden, q_scale, normalize=True):
    """qkv [rows, 3 * hidden] in place -> inv_norm [rows, 3] (se_atten.py:1368-1373)."""
    dev = _need_cuda(("qkv", qkv))
    assert qkv.is_contiguous()
    rows = qkv.numel() // (3 * hidden)
    inv = torch.empty((rows, 3), dtype=qkv.dtype, device=dev)
    lib().call("attn_qkv_normalize_" + _suffix(qkv), _p(qkv), _p(inv), rows, int(hidden), float(q_scale),
               int(bool(normalize)), _stream(dev))
    return inv


def attn_qkv_normalize_grad(d_qkv, qkv_hat, inv, hidden, q_scale, normalize=True):
    dev = _need_cuda(("d_qkv", d_qkv), ("qkv_hat", qkv_hat))
    assert d_qkv.is_contiguous() and qkv_hat.is_contiguous()
    rows = d_qkv.numel() // (3 * hidden)
    lib().call("attn_qkv_normalize_grad_" + _suffix(d_qkv), _p(d_qkv), _p(qkv_hat), _p(inv), rows, int(hidden),
               float(q_scale), int(bool(normalize)), _stream(dev))


def attn_weights(S, sw, rhat, shift=20.0, dotr=True, nnei_full=None):
    """S [natoms, nnei, nnei] -> (P, A): gated softmax and the attention weights (se_atten.py:1386-1416).
    nnei_full > nnei: the slab holds only the first nnei slots, the omitted ones are empty (trailing padding)."""
    dev = _need_cuda(("S", S), ("sw", sw), ("rhat", rhat))
    S = _c(S)
    natoms, n = S.shape[0], S.shape[1]
    P, A = torch.empty_like(S), torch.empty_like(S)
    lib().call("attn_weights_" + _suffix(S), _p(P), _p(A), _p(S), _p(_c(sw, S.dtype)), _p(_c(rhat, S.dtype)), natoms, n,
               int(n if nnei_full is None else nnei_full), float(shift), int(bool(dotr)), _stream(dev))
    return P, A


def attn_weights_grad(dA, P, S, sw, rhat, d_sw, d_rhat, shift=20.0, dotr=True):
    """dA -> dS, written over dA; accumulates into d_sw [natoms, nnei] and d_rhat [natoms, nnei, 3]."""
    dev = _need_cuda(("dA", dA), ("P", P), ("S", S))
    assert dA.is_contiguous() and d_sw.is_contiguous() and d_rhat.is_contiguous()
    natoms, n = S.shape[0], S.shape[1]
    lib().call("attn_weights_grad_" + _suffix(S), _p(dA), _p(d_sw), _p(d_rhat), _p(dA), _p(P), _p(S), _p(_c(sw, S.dtype)),
               _p(_c(rhat, S.dtype)), natoms, n, float(shift), int(bool(dotr)), _stream(dev))
    return dA


def attn_residual_layernorm(x, y, gamma, beta, eps):
    """LayerNorm(x + y) * gamma + beta over the last axis; y is overwritten with the normalised rows.
    Returns (out, zhat (= y's storage), rstd)."""
    dev = _need_cuda(("x", x), ("y", y))
    assert x.is_contiguous() and y.is_contiguous()
    C_ = x.shape[-1]
    rows = x.numel() // C_
    out = torch.empty_like(x)
    rstd = torch.empty(rows, dtype=x.dtype, device=dev)
    lib().call("attn_residual_layernorm_" + _suffix(x), _p(out), _p(y), _p(rstd), _p(x), _p(_c(gamma, x.dtype)),
               _p(_c(beta, x.dtype)), rows, C_, float(eps), _stream(dev))
    return out, y, rstd


def attn_residual_layernorm_grad(dout, zhat, rstd, gamma):
    dev = _need_cuda(("dout", dout), ("zhat", zhat))
    dout = _c(dout)
    C_ = dout.shape[-1]
    dz = torch.empty_like(dout)
    lib().call("attn_residual_layernorm_grad_" + _suffix(dout), _p(dz), _p(dout), _p(zhat), _p(rstd),
               _p(_c(gamma, dout.dtype)), rstd.numel(), C_, _stream(dev))
    return dz


def prod_force_grad_a(grad, in_deriv, nlist, nloc, nnei, nframes=1, ngrad=None):
    """deepmd::prod_force_grad_a_gpu (source/lib/include/prod_force_grad.h:26-33): gradient of prod_force_a with
    respect to net_deriv, grad [nframes, ngrad*3] -> grad_net [nframes*nloc, nnei*4].  ngrad (default nloc, the
    reference op): atoms per frame in `grad`; neighbour indices beyond it are folded with j % ngrad."""
    dev = _need_cuda(("grad", grad), ("in_deriv", in_deriv), ("nlist", nlist))
    s = _suffix(grad)
    dt = grad.dtype
    grad, in_deriv, nlist = _c(grad), _c(in_deriv, dt), _c(nlist, torch.int32)
    ngrad = nloc if ngrad is None else int(ngrad)
    out = torch.empty((nframes * nloc, nnei * 4), dtype=dt, device=dev)
    lib().call("prod_force_grad_a_ex_" + s, _p(out), _p(grad), _p(in_deriv), _p(nlist), int(nloc), ngrad, int(nnei),
               int(nframes), _stream(dev))
    return out


def prod_virial_grad_a(grad, in_deriv, rij, nlist, nloc, nnei):
    """deepmd::prod_virial_grad_a_gpu (prod_virial_grad.h:26-33): grad [9] -> grad_net [nloc, nnei*4]."""
    dev = _need_cuda(("grad", grad), ("in_deriv", in_deriv), ("rij", rij), ("nlist", nlist))
    s = _suffix(grad)
    dt = grad.dtype
    grad, in_deriv, rij, nlist = _c(grad), _c(in_deriv, dt), _c(rij, dt), _c(nlist, torch.int32)
    out = torch.empty((nloc, nnei * 4), dtype=dt, device=dev)
    lib().call("prod_virial_grad_a_" + s, _p(out), _p(grad), _p(in_deriv), _p(rij), _p(nlist), int(nloc), int(nnei),
               _stream(dev))
    return out


def prod_force_virial_a_ex(force, virial, atom_virial, net_deriv, in_deriv, rij, nlist, nrows, center_offset, nall,
                           nnei, accumulate):
    """Atom-chunked fused scatter into caller-provided force[nall*3] / virial[9] / atom_virial|None."""
    dev = _need_cuda(("net_deriv", net_deriv), ("in_deriv", in_deriv), ("rij", rij), ("nlist", nlist), ("force", force))
    s = _suffix(net_deriv)
    dt = net_deriv.dtype
    net_deriv, in_deriv, rij, nlist = _c(net_deriv), _c(in_deriv, dt), _c(rij, dt), _c(nlist, torch.int32)
    lib().call("prod_force_virial_a_ex_" + s, _p(force), _p(virial), _p(atom_virial), _p(net_deriv), _p(in_deriv),
               _p(rij), _p(nlist), int(nrows), int(center_offset), int(nall), int(nnei), int(bool(accumulate)),
               _stream(dev))
    return force, virial, atom_virial


# ------------------------------------------------------------------------------------------------
# a2-a4: neighbour-list front end
# ------------------------------------------------------------------------------------------------
def _box_host(box, dtype):
    b = torch.as_tensor(box).detach().to("cpu").reshape(9).to(dtype).contiguous()
    return b


def normalize_coord(coord, box):
    """deepmd::normalize_coord_gpu (source/lib/include/coord.h:47-55), in place; returns coord."""
    dev = _need_cuda(("coord", coord))
    s = _suffix(coord)
    if not coord.is_contiguous():
        raise ValueError("dpb200: normalize_coord works in place and needs a contiguous tensor")
    b = _box_host(box, coord.dtype)
    lib().call("normalize_coord_" + s, _p(coord), coord.numel() // 3, C.c_void_p(b.data_ptr()), _stream(dev))
    return coord


def _buf(cache, name, shape, dtype, dev):
    """A tensor of `shape` carved from a cached flat buffer (grown when too small): neighbour-list
    rebuilds then cause no allocator traffic (a 1.5 GB cudaMalloc in the middle of an MD run costs more
    than the rebuild itself)."""
    n = 1
    for d in shape:
        n *= int(d)
    if cache is None:
        return torch.empty(shape, dtype=dtype, device=dev)
    t = cache.get(name)
    if t is None or t.numel() < n or t.dtype != dtype or t.device != dev:
        cache[name] = None
        t = torch.empty(max(n, 1), dtype=dtype, device=dev)
        cache[name] = t
    return t[:n].view(shape)


def copy_coord(coord, atype, box, rcut, mem_nall=None, cache=None):
    """deepmd::copy_coord_gpu (coord.h:57-85) with the caller-side retry folded in:
    returns (ext_coord[nall,3], ext_type[nall], mapping[nall])."""
    dev = _need_cuda(("coord", coord), ("type", atype))
    s = _suffix(coord)
    coord = _c(coord).reshape(-1, 3)
    atype = _c(atype, torch.int32).reshape(-1)
    nloc = atype.numel()
    b = _box_host(box, coord.dtype)
    L = lib()
    ws = _buf(cache, "cc_ws", (max(int(L.cdll.dpb200_copy_coord_workspace_bytes(nloc)), 256),), torch.uint8, dev)
    mem = int(mem_nall) if mem_nall is not None else max(64, 2 * nloc)
    if cache is not None and mem_nall is None and cache.get("cc_nall"):
        mem = max(64, int(cache["cc_nall"] * 1.05) + 64)
    for _ in range(8):
        out_c = _buf(cache, "cc_c", (mem, 3), coord.dtype, dev)
        out_t = _buf(cache, "cc_t", (mem,), torch.int32, dev)
        mapping = _buf(cache, "cc_m", (mem,), torch.int32, dev)
        nall = C.c_int(0)
        rc = L.call("copy_coord_" + s, _p(out_c), _p(out_t), _p(mapping), C.byref(nall), _p(coord), _p(atype), nloc,
                    mem, float(rcut), C.c_void_p(b.data_ptr()), _p(ws), ws.numel(), _stream(dev))
        if rc == 0:
            n = nall.value
            if cache is not None:
                cache["cc_nall"] = n
            return out_c[:n], out_t[:n], mapping[:n]
        if mem_nall is not None:
            raise MemoryError(f"copy_coord: nall={nall.value} exceeds mem_nall={mem_nall}")
        mem = nall.value
    raise RuntimeError("copy_coord: retry limit reached")


def build_nlist(coord, nloc, rcut, atype=None, mem_size=None, cache=None):
    """deepmd::build_nlist_gpu (source/lib/include/neighbor_list.h:256-266), cell list, with the
    caller-side doubling retry folded in: returns (numneigh[nloc], rows[nloc, mem_size])."""
    dev = _need_cuda(("coord", coord), ("type", atype))
    s = _suffix(coord)
    coord = _c(coord).reshape(-1, 3)
    nall = coord.shape[0]
    if atype is not None:
        atype = _c(atype, torch.int32)
    L = lib()
    ws = _buf(cache, "bn_ws", (max(int(L.cdll.dpb200_build_nlist_workspace_bytes(nall)), 256),), torch.uint8, dev)
    mem = int(mem_size) if mem_size is not None else 256
    if cache is not None and mem_size is None and cache.get("bn_mem"):
        mem = int(cache["bn_mem"])
    for _ in range(8):
        numneigh = _buf(cache, "bn_n", (nloc,), torch.int32, dev)
        rows = _buf(cache, "bn_rows", (nloc, mem), torch.int32, dev)
        mx = C.c_int(0)
        rc = L.call("build_nlist_" + s, _p(numneigh), _p(rows), C.byref(mx), _p(coord), nloc, nall, mem, float(rcut),
                    _p(atype), _p(ws), ws.numel(), _stream(dev))
        if rc == 0:
            if cache is not None:
                cache["bn_mem"] = mem
            return numneigh, rows
        if mem_size is not None:
            raise MemoryError(f"build_nlist: a row needs {mx.value} entries, mem_size={mem_size}")
        mem = (mx.value + 31) // 32 * 32
    raise RuntimeError("build_nlist: retry limit reached")


def use_nlist_map(nlist, mapping):
    """deepmd::use_nlist_map (neighbor_list.h:219-222), in place."""
    dev = _need_cuda(("nlist", nlist), ("mapping", mapping))
    if nlist.dtype != torch.int32 or not nlist.is_contiguous():
        raise ValueError("dpb200: nlist must be a contiguous int32 tensor")
    mapping = _c(mapping, torch.int32)
    nnei = nlist.shape[-1]
    lib().cdll.dpb200_use_nlist_map(_p(nlist), _p(mapping), nlist.numel() // max(nnei, 1), nnei, _stream(dev))
    return nlist


def se_a_descriptor(gr, axis, scale, rows=None):
    """D[i] = (gr[r]*scale)^T (gr[r]*scale)[:, :axis], r = rows[i] (or i): [*,4,M] -> [n, M*axis]."""
    dev = _need_cuda(("gr", gr), ("rows", rows))
    s = _suffix(gr)
    gr = _c(gr)
    M = gr.shape[2]
    if rows is not None:
        rows = _c(rows, torch.int32)
    n = gr.shape[0] if rows is None else rows.numel()
    out = torch.empty((n, M * int(axis)), dtype=gr.dtype, device=dev)
    lib().call("se_a_descriptor_" + s, _p(out), _p(gr), _p(rows), n, M, int(axis), float(scale), _stream(dev))
    return out


def se_a_descriptor_grad(dD, gr, axis, scale, rows=None, out=None):
    """dE/d(gr[r]) for r = rows[i] (or i) from dE/dD[i]; written into `out` ([*,4,M], default new)."""
    dev = _need_cuda(("dD", dD), ("gr", gr), ("rows", rows))
    s = _suffix(gr)
    gr, dD = _c(gr), _c(dD, gr.dtype)
    M = gr.shape[2]
    if rows is not None:
        rows = _c(rows, torch.int32)
    n = dD.shape[0]
    if out is None:
        out = torch.empty_like(gr)
    lib().call("se_a_descriptor_grad_" + s, _p(out), _p(dD), _p(gr), _p(rows), n, M, int(axis), float(scale),
               _stream(dev))
    return out


def tabulate_sections_desc(tables, infos, em, sec, last_layer_size, axis, scale, desc_row=None, mode=1, nslice=7,
                           pad_rows=0, is_sorted=True, flags=None):
    """tabulate_sections_fwd with the se_e2_a descriptor contraction fused into the last section's
    epilogue (dpb200_tabulate_fusion_se_a_desc).  Returns (out [nloc,4,M], desc, row_exp|None):
      mode 1            desc = D [nloc+pad, M*axis] in em.dtype;
      mode 2, float64   desc = int8 [nloc+pad, nslice*M*axis] balanced base-256 digit slices, row_exp int32 [nloc+pad];
      mode 2, float32   desc = float32 [nloc+pad, 2*M*axis] = TF32 head | tail;
      mode 3, float32   desc = int8 [nloc+pad, 4*M*axis] digit slices of the 32-bit fixed-point image, row_exp int32
                        (operand of the int8 tensor-core fitting net, nslice = 4).
    Row desc_row[i] (int32; None: i) belongs to atom i; `pad_rows` extra zero rows are appended.
    flags: per-table DPB200_TAB_COMPRESSED_COEF words from `compressed_coef_flags` (None: full fp64 coefficients)."""
    dev = _need_cuda(("em", em), ("desc_row", desc_row))
    s = _suffix(em)
    em = _c(em)
    nloc = em.shape[0]
    nnei = int(sec[-1])
    M = int(last_layer_size)
    axis = int(axis)
    K = M * axis
    out = torch.empty((nloc, 4, M), dtype=em.dtype, device=dev)
    rows = nloc + int(pad_rows)
    row_exp = None
    if mode == 1:
        desc = torch.empty((rows, K), dtype=em.dtype, device=dev)
    elif em.dtype == torch.float64 or mode == 3:
        if mode == 3 and (em.dtype != torch.float32 or int(nslice) != 4):
            raise ValueError("dpb200: descriptor mode 3 is the float32 form with nslice = 4")
        desc = torch.empty((rows, int(nslice) * K), dtype=torch.int8, device=dev)
        row_exp = torch.empty((rows,), dtype=torch.int32, device=dev)
        if pad_rows:
            row_exp[nloc:].zero_()
    else:
        desc = torch.empty((rows, 2 * K), dtype=torch.float32, device=dev)
    if pad_rows:
        desc[nloc:].zero_()
    if desc_row is not None:
        desc_row = _c(desc_row, torch.int32)
    live = [t for t in range(len(tables)) if int(sec[t + 1] - sec[t]) > 0]
    if not live:
        raise ValueError("dpb200: tabulate_sections_desc needs at least one non-empty type section")
    esz = em.element_size()
    for n_done, t in enumerate(live):
        n_t = int(sec[t + 1] - sec[t])
        ti = _info_host(infos[t], em.dtype)
        base = C.c_void_p(em.data_ptr() + int(sec[t]) * 4 * esz)
        acc = 0 if n_done == 0 else 1
        fl = 0 if flags is None else int(flags[t])
        if t != live[-1]:
            lib().call("tabulate_fusion_se_a_desc_" + s, _p(out), _p(_c(tables[t])), C.c_void_p(ti.data_ptr()), base,
                       nnei * 4, 4, base, nnei * 4, nloc, n_t, M, int(bool(is_sorted)), acc, axis, float(scale),
                       None, 0, None, 0, 0, None, fl, _stream(dev))
        else:
            lib().call("tabulate_fusion_se_a_desc_" + s, _p(out), _p(_c(tables[t])), C.c_void_p(ti.data_ptr()), base,
                       nnei * 4, 4, base, nnei * 4, nloc, n_t, M, int(bool(is_sorted)), acc, axis, float(scale),
                       _p(desc_row), int(mode), _p(desc), int(desc.shape[1]), int(nslice), _p(row_exp), fl, _stream(dev))
    return out, desc, row_exp


def split_i8_rows(x, nslice):
    """fp64 [n, w] -> (int8 [n, nslice*w] balanced base-256 digit slices, most significant first; row_exp int32 [n]):
    x = 2^row_exp * sum_s slice_s * 2^(-7-8s)  (truncated after nslice digits)."""
    dev = _need_cuda(("x", x))
    if x.dtype != torch.float64 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("dpb200: split_i8_rows needs a float64 matrix with unit column stride")
    n, w = x.shape
    out = torch.empty((n, int(nslice) * w), dtype=torch.int8, device=dev)
    ex = torch.empty((n,), dtype=torch.int32, device=dev)
    lib().call("split_i8_rows_f64", _p(out), int(out.shape[1]), _p(ex), _p(x), int(x.stride(0)) if n > 1 else w, n, w,
               int(nslice), _stream(dev))
    return out, ex


def split_i8_combine(acc, row_exp, col_exp, bias=None, idt=None, h=None, activation=True):
    """acc int32 [nslice, n, w] (order sums of the split-integer GEMM) -> fp64.  activation: returns
    (a = tanh(z), y = a*idt + h); otherwise z."""
    dev = _need_cuda(("acc", acc), ("row_exp", row_exp), ("col_exp", col_exp), ("bias", bias), ("idt", idt), ("h", h))
    if acc.dtype != torch.int32 or acc.dim() != 3 or not acc.is_contiguous():
        raise ValueError("dpb200: split_i8_combine needs a contiguous int32 [nslice, n, w] tensor")
    ns, n, w = acc.shape
    a = torch.empty((n, w), dtype=torch.float64, device=dev)
    y = torch.empty((n, w), dtype=torch.float64, device=dev) if activation else None
    lib().call("split_i8_combine_f64", _p(a), _p(y), _p(acc), n * w, ns, _p(row_exp), _p(col_exp), _p(bias), _p(idt),
               _p(h), n, w, int(bool(activation)), _stream(dev))
    return (a, y) if activation else a


def split_tf32(x, copies=3):
    """fp32 [n, w] -> [n, copies*w] = [hi | lo (| hi)], hi = tf32(x), lo = tf32(x - hi)."""
    dev = _need_cuda(("x", x))
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("dpb200: split_tf32 needs a float32 matrix with unit column stride")
    n, w = x.shape
    out = torch.empty((n, int(copies) * w), dtype=torch.float32, device=dev)
    lib().call("split_tf32_f32", _p(out), int(out.shape[1]), _p(x), int(x.stride(0)) if n > 1 else w, n, w, int(copies),
               _stream(dev))
    return out


def mlp_tanh_fwd(z, h=None, idt=None, split3=False):
    """In place a = tanh(z); returns y = a*idt (+ h).  `z` becomes `a` (kept for the backward).
    split3 (fp32): also returns y as the 3xTF32 operand [hi | lo | hi] ([n, 3w])."""
    dev = _need_cuda(("z", z), ("h", h), ("idt", idt))
    s = _suffix(z)
    y = torch.empty_like(z)
    if not split3:
        lib().call("mlp_tanh_fwd_" + s, _p(z), _p(y), _p(h), _p(idt), z.shape[0], z.shape[1], _stream(dev))
        return y
    sp = torch.empty((z.shape[0], 3 * z.shape[1]), dtype=z.dtype, device=dev)
    lib().call("mlp_tanh_fwd_split_" + s, _p(z), _p(y), _p(h), _p(idt), z.shape[0], z.shape[1], _p(sp), _stream(dev))
    return y, sp


def mlp_tanh_bwd(g, a, idt=None, split3=False):
    """t = g * idt * (1 - a^2); g may be a row-strided view (leading stride g.stride(0)).
    split3 (fp32): returns t only as the 3xTF32 operand [hi | lo | hi] ([n, 3w])."""
    dev = _need_cuda(("g", g), ("a", a), ("idt", idt))
    s = _suffix(a)
    if g.stride(1) != 1 and g.shape[0] > 1:
        g = g.contiguous()
    ldg = g.stride(0) if g.shape[0] > 1 else g.shape[1]
    if split3:
        sp = torch.empty((a.shape[0], 3 * a.shape[1]), dtype=a.dtype, device=dev)
        lib().call("mlp_tanh_bwd_split_" + s, None, _p(g), int(ldg), _p(a), _p(idt), a.shape[0], a.shape[1], _p(sp),
                   _stream(dev))
        return sp
    t = torch.empty_like(a)
    lib().call("mlp_tanh_bwd_" + s, _p(t), _p(g), int(ldg), _p(a), _p(idt), a.shape[0], a.shape[1], _stream(dev))
    return t


def fit_gemm_i8(mode, n, N, K, a_slices, a_slice_stride, a_row_stride, row_exp, row_exp_fixed, b_slices, b_k_stride,
                colv, skip=None, t_in=None, out0=None, out1=None, ld_out=0, slices_out=None, ld_slices=0, kp_out=0,
                out_exp=0, nslice=6):
    """One fitting-net GEMM on the int8 tensor cores (dpb200_fit_gemm_i8_f64, csrc/fit_tc.cu): mode 0 forward layer,
    1 backward layer, 2 plain product, 3 plain product stored as float32 (out0 is a float32 matrix).  nslice 6 (fp64
    model) or 4 (fp32 model).  colv [N, 4] = {2^(col_exp-12), add, mul, 0} per output column; fp64
    intermediates are in the row-blocked layout of fit_blocked()."""
    dev = _need_cuda(("a_slices", a_slices), ("b_slices", b_slices), ("colv", colv), ("row_exp", row_exp),
                     ("skip", skip), ("t_in", t_in), ("out0", out0), ("out1", out1), ("slices_out", slices_out))
    lib().call("fit_gemm_i8_f64", int(mode), int(n), int(N), int(K), int(nslice), _p(a_slices), int(a_slice_stride),
               int(a_row_stride), _p(row_exp), int(row_exp_fixed), _p(b_slices), int(b_k_stride), _p(colv), _p(skip),
               _p(t_in), _p(out0), _p(out1), int(ld_out), _p(slices_out), int(ld_slices), int(kp_out), int(out_exp),
               _stream(dev))


def fit_blocked_rows(n):
    return (int(n) + 127) // 128 * 128


def fit_blocked(src, n, N, to_blocked=True):
    """Row-major fp64 [n, N] <-> the row-blocked column-major layout of the tensor-core fitting net:
    element (r, c) at ((r // 128) * N + c) * 128 + r % 128 (flat tensor of fit_blocked_rows(n) * N elements)."""
    dev = _need_cuda(("src", src))
    if to_blocked:
        src = _c(src, torch.float64)
        dst = torch.zeros(fit_blocked_rows(n) * int(N), dtype=torch.float64, device=dev)
        lib().call("fit_blocked_f64", _p(dst), _p(src), int(N), int(n), int(N), 1, _stream(dev))
    else:
        dst = torch.empty((int(n), int(N)), dtype=torch.float64, device=dev)
        lib().call("fit_blocked_f64", _p(dst), _p(src), int(N), int(n), int(N), 0, _stream(dev))
    return dst


def fit_slice_rows(x_blocked, n, N, kp, nslice=6):
    """Blocked fp64 [n, N] -> (int8 [n, nslice*kp] slices with per-row exponent, row_exp int32 [n])."""
    dev = _need_cuda(("x", x_blocked))
    out = torch.empty((int(n), int(nslice) * int(kp)), dtype=torch.int8, device=dev)
    ex = torch.empty((int(n),), dtype=torch.int32, device=dev)
    lib().call("fit_slice_rows_f64", _p(out), int(out.shape[1]), int(kp), _p(ex), _p(x_blocked), int(n), int(N),
               int(nslice), _stream(dev))
    return out, ex


def fit_head(t, y, w_head, idt, b_head, n, N, kp, nslice=6):
    """Energy head + seed of the backward chain: e = y . w_head + b_head; dz = w_head * idt * (1 - t^2) as int8 slices."""
    dev = _need_cuda(("t", t), ("y", y), ("w_head", w_head), ("idt", idt))
    e = torch.empty((int(n),), dtype=torch.float64, device=dev)
    out = torch.empty((int(n), int(nslice) * int(kp)), dtype=torch.int8, device=dev)
    ex = torch.empty((int(n),), dtype=torch.int32, device=dev)
    lib().call("fit_head_f64", _p(e), _p(out), int(out.shape[1]), int(kp), _p(ex), _p(t), _p(y), _p(w_head), _p(idt),
               float(b_head), int(n), int(N), int(nslice), _stream(dev))
    return e, out, ex


def halo_pack(coord, sendlist, shift):
    """sendbuf[k] = coord[sendlist[k]] + shift[k] (dpb200_halo_pack)."""
    dev = _need_cuda(("coord", coord), ("sendlist", sendlist), ("shift", shift))
    s = _suffix(coord)
    n = sendlist.numel()
    out = torch.empty((n, 3), dtype=coord.dtype, device=dev)
    lib().call("halo_pack_" + s, _p(out), _p(_c(coord)), _p(_c(sendlist, torch.int32)), _p(_c(shift, coord.dtype)), n,
               _stream(dev))
    return out


def halo_unpack_add(force, recvbuf, sendlist):
    """force[sendlist[k]] += recvbuf[k] in place (dpb200_halo_unpack_add)."""
    dev = _need_cuda(("force", force), ("recvbuf", recvbuf), ("sendlist", sendlist))
    s = _suffix(force)
    if not force.is_contiguous():
        raise ValueError("dpb200: halo_unpack_add works in place and needs a contiguous force tensor")
    lib().call("halo_unpack_add_" + s, _p(force), _p(_c(recvbuf, force.dtype)), _p(_c(sendlist, torch.int32)),
               sendlist.numel(), _stream(dev))
    return force


# ------------------------------------------------------------------------------------------------
# torch.ops.deepmd.*  (reference operator surface)
# ------------------------------------------------------------------------------------------------
class _TabGradOp(torch.autograd.Function):
    """TabulateFusionSeAGradOp / TabulateFusionSeAttenGradOp
    (source/op/pt/tabulate_multi_device.cc:731-831, 964-1061): differentiable w.r.t. dy only."""

    @staticmethod
    def forward(ctx, table, table_info, em_x, em, two_embed, dy, M, is_sorted):
        ctx.save_for_backward(table, table_info, em_x, em, two_embed)
        ctx.M, ctx.is_sorted = M, is_sorted
        gx, gem, gtwo = tabulate_fusion_se_a_grad(table, table_info, em_x, em, dy, M, two_embed, is_sorted)
        if gtwo is None:
            return gx, gem
        return gx, gem, gtwo

    @staticmethod
    def backward(ctx, *dz):
        table, table_info, em_x, em, two_embed = ctx.saved_tensors
        dz_x = dz[0] if dz[0] is not None else torch.zeros_like(em_x)
        dz_em = dz[1] if dz[1] is not None else torch.zeros_like(em)
        dz_two = dz[2] if len(dz) > 2 else None
        dz_dy = tabulate_fusion_se_a_grad_grad(table, table_info, em_x, em, dz_x, dz_em, ctx.M, two_embed, dz_two,
                                               ctx.is_sorted)
        return None, None, None, None, None, dz_dy, None, None


class _TabOp(torch.autograd.Function):
    """TabulateFusionSeAOp / TabulateFusionSeAttenOp (tabulate_multi_device.cc:881-962, 1063-1151)."""

    @staticmethod
    def forward(ctx, table, table_info, em_x, em, two_embed, M, is_sorted):
        ctx.save_for_backward(table, table_info, em_x, em, two_embed)
        ctx.M, ctx.is_sorted = M, is_sorted
        return tabulate_fusion_se_a(table, table_info, em_x, em, M, two_embed, is_sorted)

    @staticmethod
    def backward(ctx, dy):
        table, table_info, em_x, em, two_embed = ctx.saved_tensors
        res = _TabGradOp.apply(table, table_info, em_x, em, two_embed, dy.contiguous(), ctx.M, ctx.is_sorted)
        gtwo = res[2] if len(res) > 2 else None
        return None, None, res[0], res[1], gtwo, None, None


def _op_tabulate_fusion_se_a(table, table_info, em_x, em, last_layer_size):
    return [_TabOp.apply(table, table_info, em_x, em, None, int(last_layer_size), True)]


def _op_tabulate_fusion_se_atten(table, table_info, em_x, em, two_embed, last_layer_size, is_sorted):
    return [_TabOp.apply(table, table_info, em_x, em, two_embed, int(last_layer_size), bool(is_sorted))]


def decode_mesh(mesh: torch.Tensor, nloc: int):
    """Neighbour-list hand-off encodings of the `mesh` argument (SURVEY.md 8b;
    source/lib/src/prod_env_mat.cc:327-332, source/op/tf/prod_env_mat_multi_device.cc:2813-2828).
    Returns (mode, ilist, numneigh, neigh) with CSR rows, tensors on the CPU or on mesh.device."""
    n = mesh.numel()
    if n == 0:
        return "nopbc", None, None, None
    if n == 6:
        return "pbc", None, None, None
    if n == 16:
        m = mesh.detach().to("cpu", torch.int32).contiguous().numpy()
        inum = int(m[1])
        as_ptr = lambda k: int(np.frombuffer(m[k:k + 2].tobytes(), dtype=np.uint64)[0])
        ilist = np.ctypeslib.as_array(C.cast(as_ptr(4), C.POINTER(C.c_int)), shape=(inum,)).copy()
        numneigh = np.ctypeslib.as_array(C.cast(as_ptr(8), C.POINTER(C.c_int)), shape=(inum,)).copy()
        first = C.cast(as_ptr(12), C.POINTER(C.POINTER(C.c_int)))
        rows = [np.ctypeslib.as_array(first[r], shape=(int(numneigh[r]),)).copy() if numneigh[r] > 0
                else np.zeros(0, np.int32) for r in range(inum)]
        neigh = np.concatenate(rows) if rows else np.zeros(0, np.int32)
        return ("list", torch.from_numpy(ilist.astype(np.int32)), torch.from_numpy(numneigh.astype(np.int32)),
                torch.from_numpy(neigh.astype(np.int32)))
    if n > 16:
        m = mesh.reshape(-1)
        inum = int(m[1].item())
        ilist = m[16:16 + inum]
        numneigh = m[16 + inum:16 + 2 * inum]
        neigh = m[16 + 2 * inum:]
        return "list", ilist, numneigh, neigh
    raise ValueError(f"invalid mesh tensor of length {n} (mixed-type meshes of length 1/7 are not supported)")


def _op_prod_env_mat_a(coord, type, natoms, box, mesh, davg, dstd, rcut_a: float, rcut_r: float, rcut_r_smth: float,
                       sel_a: List[int], sel_r: List[int]):
    """ProdEnvMatA (source/op/tf/prod_env_mat_multi_device.cc:13-30, host logic :451-905)."""
    dev = _need_cuda(("coord", coord), ("type", type), ("davg", davg), ("dstd", dstd))
    if coord.dim() != 2 or type.dim() != 2:
        raise ValueError("Dim of coord and type should be 2")
    nat = natoms.detach().to("cpu").reshape(-1).tolist()
    if len(nat) < 3:
        raise ValueError("number of atoms should be larger than (or equal to) 3")
    nloc, nall = int(nat[0]), int(nat[1])
    ntypes = len(nat) - 2
    if len(sel_a) != ntypes:
        raise ValueError("number of types should match the length of sel array")
    if any(int(s) != 0 for s in sel_r):
        raise NotImplementedError("prod_env_mat_a: radial-only selections (sel_r) are outside the se_a hot path")
    nf = coord.shape[0]
    if coord.shape[1] != nall * 3 or type.shape[1] != nall or type.shape[0] != nf:
        raise ValueError("number of atoms should match")
    sec = [0]
    for s_ in sel_a:
        sec.append(sec[-1] + int(s_))
    nnei = sec[-1]
    if davg.numel() != ntypes * nnei * 4 or dstd.numel() != ntypes * nnei * 4:
        raise ValueError("number of avg / std should be ntype * ndescrpt")
    mode, ilist, numneigh, neigh = decode_mesh(mesh, nloc)
    dt = coord.dtype
    if mode == "list":
        if nf != 1:
            raise ValueError("a neighbour list passed through mesh implies a single frame")
        ilist, numneigh, neigh = (t.to(dev, torch.int32).contiguous() for t in (ilist, numneigh, neigh))
        if ilist.numel() != nloc:
            raise ValueError(f"neighbour list has {ilist.numel()} centre atoms but natoms[0] = {nloc}")
        if neigh.numel() == 0:
            neigh = torch.zeros(1, dtype=torch.int32, device=dev)
        fn = csr_row_pointers(neigh, numneigh)
        mx = int(numneigh.max().item()) if numneigh.numel() else 0
        if mx > MAX_NBOR_SIZE:
            raise ValueError(f"Neighbor row is larger than the maximum supported size {MAX_NBOR_SIZE}")
        em, dv, rij, nl = prod_env_mat_a(coord.reshape(-1), type.reshape(-1), numneigh, None, davg, dstd, nloc, nall,
                                         rcut_r, rcut_r_smth, sec, ilist=ilist, firstneigh=fn, max_nbor_size=mx)
        return (em.reshape(1, -1), dv.reshape(1, -1), rij.reshape(1, -1), nl.reshape(1, -1))
    outs = ([], [], [], [])
    for f in range(nf):
        c = coord[f].reshape(-1, 3).to(dt).contiguous().clone()
        t = type[f].to(torch.int32).contiguous()
        if mode == "pbc":
            if nall != nloc:
                raise ValueError("PBC mesh (length 6) expects nall == nloc; ghosts are generated by the op")
            normalize_coord(c, box[f])
            ext_c, ext_t, mapping = copy_coord(c, t, box[f], rcut_r)
        else:
            ext_c, ext_t, mapping = c, t, None
        n_ext = ext_t.numel()
        numneigh_f, rows_f = build_nlist(ext_c, nloc, rcut_r, ext_t)
        em, dv, rij, nl = prod_env_mat_a(ext_c.reshape(-1), ext_t, numneigh_f, rows_f, davg, dstd, nloc, n_ext, rcut_r,
                                         rcut_r_smth, sec)
        if mapping is not None:
            use_nlist_map(nl, mapping)
        for o, v in zip(outs, (em, dv, rij, nl)):
            o.append(v.reshape(1, -1))
    return tuple(torch.cat(o, 0) for o in outs)


def _op_prod_force_se_a(net_deriv, in_deriv, nlist, natoms, n_a_sel: int, n_r_sel: int):
    """ProdForceSeA (source/op/tf/prod_force_multi_device.cc:6-14, 50-167)."""
    nat = natoms.detach().to("cpu").reshape(-1).tolist()
    nloc, nall = int(nat[0]), int(nat[1])
    nnei = int(n_a_sel) + int(n_r_sel)
    if net_deriv.dim() != 2 or in_deriv.dim() != 2 or nlist.dim() != 2:
        raise ValueError("Dim of net deriv, input deriv and nlist should be 2")
    nf = net_deriv.shape[0]
    if in_deriv.shape[0] != nf or nlist.shape[0] != nf:
        raise ValueError("number of samples should match")
    if nloc * nnei * 4 != net_deriv.shape[1] or nloc * nnei * 12 != in_deriv.shape[1] or nloc * nnei != nlist.shape[1]:
        raise ValueError("number of descriptors should match")
    return _ForceOp.apply(net_deriv, in_deriv, nlist, nloc, nall, nnei)


def _op_prod_virial_se_a(net_deriv, in_deriv, rij, nlist, natoms, n_a_sel: int, n_r_sel: int):
    """ProdVirialSeA (source/op/tf/prod_virial_multi_device.cc:5-15, 40-140)."""
    nat = natoms.detach().to("cpu").reshape(-1).tolist()
    nloc, nall = int(nat[0]), int(nat[1])
    nnei = int(n_a_sel) + int(n_r_sel)
    if net_deriv.dim() != 2 or in_deriv.dim() != 2 or rij.dim() != 2 or nlist.dim() != 2:
        raise ValueError("Dim of net deriv, input deriv, rij and nlist should be 2")
    nf = net_deriv.shape[0]
    if nloc * nnei * 4 != net_deriv.shape[1] or nloc * nnei * 12 != in_deriv.shape[1] or \
            nloc * nnei * 3 != rij.shape[1] or nloc * nnei != nlist.shape[1]:
        raise ValueError("number of descriptors should match")
    return _VirialOp.apply(net_deriv, in_deriv, rij, nlist, nloc, nall, nnei)


def _natoms(natoms):
    nat = natoms.detach().to("cpu").reshape(-1).tolist()
    return int(nat[0]), int(nat[1])


def _op_prod_force_se_a_grad(grad, net_deriv, in_deriv, nlist, natoms, n_a_sel: int, n_r_sel: int):
    """ProdForceSeAGrad (source/op/tf/prod_force_grad_multi_device.cc:5-14, 36-160): grad [nf, nloc*3]."""
    nloc, _ = _natoms(natoms)
    nnei = int(n_a_sel) + int(n_r_sel)
    if grad.dim() != 2 or net_deriv.dim() != 2 or in_deriv.dim() != 2 or nlist.dim() != 2:
        raise ValueError("Dim of grad, net deriv, input deriv and nlist should be 2")
    nf = net_deriv.shape[0]
    if grad.shape[0] != nf or in_deriv.shape[0] != nf or nlist.shape[0] != nf:
        raise ValueError("number of frames should match")
    if grad.shape[1] != nloc * 3:
        raise ValueError("input grad shape should be 3 x natoms")
    if nloc * nnei * 12 != in_deriv.shape[1] or nloc * nnei != nlist.shape[1]:
        raise ValueError("number of descriptors should match")
    return prod_force_grad_a(grad, in_deriv, nlist, nloc, nnei, nf).reshape(nf, -1)


def _op_prod_virial_se_a_grad(grad, net_deriv, in_deriv, rij, nlist, natoms, n_a_sel: int, n_r_sel: int):
    """ProdVirialSeAGrad (source/op/tf/prod_virial_grad_multi_device.cc:5-15, 37-170): grad [nf, 9]."""
    nloc, _ = _natoms(natoms)
    nnei = int(n_a_sel) + int(n_r_sel)
    if grad.dim() != 2 or net_deriv.dim() != 2 or in_deriv.dim() != 2 or rij.dim() != 2 or nlist.dim() != 2:
        raise ValueError("Dim of grad, net deriv, input deriv, rij and nlist should be 2")
    nf = net_deriv.shape[0]
    if grad.shape[0] != nf or grad.shape[1] != 9:
        raise ValueError("input grad shape should be 3 x 3")
    if nloc * nnei * 12 != in_deriv.shape[1] or nloc * nnei * 3 != rij.shape[1] or nloc * nnei != nlist.shape[1]:
        raise ValueError("number of descriptors should match")
    return torch.cat([prod_virial_grad_a(grad[f], in_deriv[f], rij[f], nlist[f], nloc, nnei).reshape(1, -1)
                      for f in range(nf)], 0)


class _ForceOp(torch.autograd.Function):
    """prod_force_se_a, differentiable in net_deriv (deepmd/tf/op/_prod_force_se_a_grad.py)."""

    @staticmethod
    def forward(ctx, net_deriv, in_deriv, nlist, nloc, nall, nnei):
        ctx.save_for_backward(in_deriv, nlist)
        ctx.dims = (nloc, nall, nnei, net_deriv.shape[0])
        return prod_force_a(net_deriv, in_deriv, nlist, nloc, nall, nnei, net_deriv.shape[0])

    @staticmethod
    def backward(ctx, g):
        in_deriv, nlist = ctx.saved_tensors
        nloc, nall, nnei, nf = ctx.dims
        gn = prod_force_grad_a(g.contiguous(), in_deriv, nlist, nloc, nnei, nf, ngrad=nall)
        return gn.reshape(nf, -1), None, None, None, None, None


class _VirialOp(torch.autograd.Function):
    """prod_virial_se_a, differentiable in net_deriv through the virial output (_prod_virial_se_a_grad.py)."""

    @staticmethod
    def forward(ctx, net_deriv, in_deriv, rij, nlist, nloc, nall, nnei):
        ctx.save_for_backward(in_deriv, rij, nlist)
        ctx.dims = (nloc, nnei, net_deriv.shape[0])
        vs, avs = [], []
        for f in range(net_deriv.shape[0]):
            v, av = prod_virial_a(net_deriv[f], in_deriv[f], rij[f], nlist[f], nloc, nall, nnei)
            vs.append(v.reshape(1, 9))
            avs.append(av.reshape(1, -1))
        av = torch.cat(avs, 0)
        ctx.mark_non_differentiable(av)
        return torch.cat(vs, 0), av

    @staticmethod
    def backward(ctx, gv, _gav):
        in_deriv, rij, nlist = ctx.saved_tensors
        nloc, nnei, nf = ctx.dims
        gn = torch.cat([prod_virial_grad_a(gv[f].contiguous(), in_deriv[f], rij[f], nlist[f], nloc, nnei).reshape(1, -1)
                        for f in range(nf)], 0)
        return gn, None, None, None, None, None, None


_REGISTERED = False
_LIBRARY = None


def register_torch_ops():
    """Define torch.ops.deepmd.* once per process (TORCH_LIBRARY_FRAGMENT(deepmd, m) in the
    reference, tabulate_multi_device.cc:1682-1696)."""
    global _REGISTERED, _LIBRARY
    if _REGISTERED:
        return
    L = torch.library.Library("deepmd", "FRAGMENT")
    L.define("tabulate_fusion_se_a(Tensor table, Tensor table_info, Tensor em_x, Tensor em, int last_layer_size) "
             "-> Tensor[]")
    L.define("tabulate_fusion_se_atten(Tensor table, Tensor table_info, Tensor em_x, Tensor em, Tensor two_embed, "
             "int last_layer_size, bool is_sorted) -> Tensor[]")
    L.define("prod_env_mat_a(Tensor coord, Tensor type, Tensor natoms, Tensor box, Tensor mesh, Tensor davg, "
             "Tensor dstd, float rcut_a, float rcut_r, float rcut_r_smth, int[] sel_a, int[] sel_r) "
             "-> (Tensor, Tensor, Tensor, Tensor)")
    L.define("prod_force_se_a(Tensor net_deriv, Tensor in_deriv, Tensor nlist, Tensor natoms, int n_a_sel, "
             "int n_r_sel) -> Tensor")
    L.define("prod_virial_se_a(Tensor net_deriv, Tensor in_deriv, Tensor rij, Tensor nlist, Tensor natoms, "
             "int n_a_sel, int n_r_sel) -> (Tensor, Tensor)")
    L.impl("tabulate_fusion_se_a", _op_tabulate_fusion_se_a, "CompositeImplicitAutograd")
    L.impl("tabulate_fusion_se_atten", _op_tabulate_fusion_se_atten, "CompositeImplicitAutograd")
    L.impl("prod_env_mat_a", _op_prod_env_mat_a, "CompositeExplicitAutograd")
    L.define("prod_force_se_a_grad(Tensor grad, Tensor net_deriv, Tensor in_deriv, Tensor nlist, Tensor natoms, "
             "int n_a_sel, int n_r_sel) -> Tensor")
    L.define("prod_virial_se_a_grad(Tensor grad, Tensor net_deriv, Tensor in_deriv, Tensor rij, Tensor nlist, "
             "Tensor natoms, int n_a_sel, int n_r_sel) -> Tensor")
    # (CompositeImplicitAutograd: the autograd.Function inside provides the derivative w.r.t. net_deriv)
    L.impl("prod_force_se_a", _op_prod_force_se_a, "CompositeImplicitAutograd")
    L.impl("prod_virial_se_a", _op_prod_virial_se_a, "CompositeImplicitAutograd")
    L.impl("prod_force_se_a_grad", _op_prod_force_se_a_grad, "CompositeExplicitAutograd")
    L.impl("prod_virial_se_a_grad", _op_prod_virial_se_a_grad, "CompositeExplicitAutograd")
    _LIBRARY = L
    _REGISTERED = True
