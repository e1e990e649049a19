"""ctypes binding of the dpb200 C ABI (include/dpb200.h).

This is the only place where Python touches the native library.  There is no CPU fallback: if
``lib/libdpb200.so`` is missing the import fails, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdpb200.so")

OK = 0
ERR_INVALID = -1
ERR_CUDA = -2
ERR_OOM = -3
ERR_NLIST_CAPACITY = -4
MAX_NBOR_SIZE = 4096

_T = {"p": C.c_void_p, "i": C.c_int, "l": C.c_longlong, "f": C.c_float, "z": C.c_size_t, "d": C.c_double}

# name -> argument codes (p pointer, i int, l long long, f float, z size_t); `{s}` = f32 | f64
_SIGS = {
    "prod_env_mat_a_{s}": "pppp ppppppp ii pp iii ff pi pz p",
    "prod_env_mat_a_ex_{s}": "pppp ppppppp ii pp i iii ff pi pz p",
    "format_nlist_{s}": "p pp pppp ii iii f pi pz p",
    "tabulate_fusion_se_a_{s}": "pppppp iiii p",
    "tabulate_fusion_se_a_grad_{s}": "ppp pp ppp p iiii p",
    "tabulate_fusion_se_a_grad_grad_{s}": "p pp ppp ppp iiii p",
    "se_atten_gate_scalars_{s}": "ppp ppp iii ff p",
    "prod_force_virial_a_pair_{s}": "ppp pppp pp iii p",
    "tabulate_fusion_se_a_nd_{s}": "pppppp iiiii p",
    "tabulate_fusion_se_a_grad_nd_{s}": "ppp pp ppp p iiiii p",
    "tabulate_fusion_se_a_grad_grad_nd_{s}": "p pp ppp ppp iiiii p",
    "tabulate_fusion_se_atten_gate_{s}": "ppppp ppp iiiii p",
    "tabulate_fusion_se_atten_gate_grad_{s}": "ppp pp pp ppp p iiiii p",
    "tabulate_fusion_se_atten_gate_desc_{s}": "ppppp ppp iiii i d p i p l l i p i i p",
    "tabulate_fusion_se_a_ex_{s}": "ppp pli pl p iiiii p",
    "tabulate_fusion_se_a_grad_ex_{s}": "ppp pp pli pl p p iiii p",
    "prod_force_a_{s}": "pppp iiii p",
    "prod_virial_a_{s}": "pppppp iii p",
    "prod_force_virial_a_{s}": "ppppppp iii p",
    "prod_force_virial_a_ex_{s}": "ppppppp iiiii p",
    "prod_force_grad_a_{s}": "pppp iii p",
    "prod_force_grad_a_ex_{s}": "pppp iiii p",
    "prod_virial_grad_a_{s}": "ppppp ii p",
    "normalize_coord_{s}": "pip p",
    "copy_coord_{s}": "pppp pp ii f p pz p",
    "copy_coord_cells_{s}": "pppp pp ii pp p pz p",
    "build_nlist_{s}": "ppp p iii f p pz p",
    "se_a_descriptor_{s}": "ppp l ii d p",
    "se_a_descriptor_grad_{s}": "pppp l ii d p",
    "mlp_tanh_fwd_{s}": "pppp l i p",
    "mlp_tanh_bwd_{s}": "pp l pp l i p",
    "mlp_tanh_fwd_split_{s}": "pppp l i p p",
    "mlp_tanh_bwd_split_{s}": "pp l pp l i p p",
    "tabulate_fusion_se_a_desc_{s}": "ppp pli pl iiiii i d p i p l i p i p",
    "tabulate_fusion_se_a_grad_fx_{s}": "pp pp pli pl p iiii i p",
    "se_atten_embed_{s}": "ppp pp p l pp p l i p",
    "se_atten_embed_grad_{s}": "p l p ppp pp p l i p",
    "se_atten_rhat_{s}": "ppp l p",
    "se_atten_rhat_grad_{s}": "pppp l p",
    "attn_qkv_normalize_{s}": "pp l i d i p",
    "attn_qkv_normalize_grad_{s}": "ppp l i d i p",
    "attn_weights_{s}": "pp ppp l ii d i p",
    "attn_weights_grad_{s}": "ppp pp ppp l i d i p",
    "attn_residual_layernorm_{s}": "ppp ppp l i d p",
    "attn_residual_layernorm_grad_{s}": "p pppp l i p",
    "halo_pack_{s}": "pppp i p",
    "halo_unpack_add_{s}": "ppp i p",
}


# entry points that exist for one FPTYPE only
_PLAIN_SIGS = {
    "split_i8_rows_f64": "p l p p l l i i p",
    "split_i8_combine_f64": "pp p l i pp ppp l i i p",
    "split_tf32_f32": "p l p l l i i p",
    "scatter_nlist_rows": "ppp i p ii p",
    "fit_gemm_i8_f64": "i l iii p l l p i p i p p p pp l p l i i p",
    "fit_slice_rows_f64": "p l i p p l i i p",
    "fit_head_f64": "p p l i p pppp d l i i p",
    "fit_blocked_f64": "pp l l i i p",
    "fit_slice_cols_f64": "p l l ii i p p i p l p",
}


class DPB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dpb200 error {code}: {msg}")
        self.code = code


class NlistCapacityError(DPB200Error):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback for the dpb200 hot path.")
        self.cdll = C.CDLL(LIB_PATH)
        self.cdll.dpb200_last_error.restype = C.c_char_p
        self.cdll.dpb200_abi_version.restype = C.c_int
        self.cdll.dpb200_launch_count.restype = C.c_longlong
        self.cdll.dpb200_count_replayed_launches.restype = None
        self.cdll.dpb200_count_replayed_launches.argtypes = [C.c_longlong]
        for s_ in ("f32", "f64"):
            f = getattr(self.cdll, "dpb200_fma_peak_" + s_)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(C.c_double), C.c_void_p]
        for fn, res in (("dpb200_prod_env_mat_a_workspace_bytes", "iiiii"), ("dpb200_copy_coord_workspace_bytes", "i"),
                        ("dpb200_build_nlist_workspace_bytes", "i")):
            f = getattr(self.cdll, fn)
            f.restype = C.c_size_t
            f.argtypes = [_T[c] for c in res]
        for fn, sig in _PLAIN_SIGS.items():
            f = getattr(self.cdll, "dpb200_" + fn)
            f.restype = C.c_int
            f.argtypes = [_T[c] for c in sig.replace(" ", "")]
        self.cdll.dpb200_use_nlist_map.restype = C.c_int
        self.cdll.dpb200_use_nlist_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        for pat, sig in _SIGS.items():
            for s in ("f32", "f64"):
                f = getattr(self.cdll, "dpb200_" + pat.format(s=s))
                f.restype = C.c_int
                f.argtypes = [_T[c] for c in sig.replace(" ", "")]

    def exported(self):
        names = ["dpb200_last_error", "dpb200_abi_version", "dpb200_launch_count", "dpb200_count_replayed_launches",
                 "dpb200_fma_peak_f32",
                 "dpb200_fma_peak_f64", "dpb200_prod_env_mat_a_workspace_bytes",
                 "dpb200_copy_coord_workspace_bytes", "dpb200_build_nlist_workspace_bytes", "dpb200_use_nlist_map"]
        for pat in _SIGS:
            for s in ("f32", "f64"):
                names.append("dpb200_" + pat.format(s=s))
        names.extend("dpb200_" + fn for fn in _PLAIN_SIGS)
        return names

    def launch_count(self) -> int:
        return int(self.cdll.dpb200_launch_count())

    def fma_peak(self, suffix: str, stream) -> float:
        v = C.c_double(0.0)
        self.call("fma_peak_" + suffix, C.byref(v), stream)
        return v.value

    def last_error(self) -> str:
        return (self.cdll.dpb200_last_error() or b"").decode()

    def call(self, name, *args):
        """Call dpb200_<name>; returns the (non-negative) status, raises on a negative one."""
        rc = getattr(self.cdll, "dpb200_" + name)(*args)
        if rc < 0:
            msg = self.last_error()
            if rc == ERR_NLIST_CAPACITY:
                raise NlistCapacityError(rc, msg)
            if rc == ERR_INVALID:
                raise ValueError(f"dpb200_{name}: {msg}")
            if rc == ERR_OOM:
                raise MemoryError(f"dpb200_{name}: {msg}")
            raise DPB200Error(rc, f"dpb200_{name}: {msg}")
        return rc


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
