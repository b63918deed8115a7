"""ctypes binding of liblitridge.so (the C ABI declared in include/litridge.h).

The library is the product: there is no Python/CPU fallback.  If it is missing, stale or fails to
load, importing the device layer raises -- loudly -- instead of silently computing elsewhere.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblitridge.so")
ABI_VERSION = 6

_vp, _l, _i, _f, _d, _sz = C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_double, C.c_size_t
_psz, _pi, _pl = C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_long)

# name -> argtypes (restype is int unless listed in _RESTYPES); order mirrors include/litridge.h
PROTOTYPES = {
    "lit_last_error": [],
    "lit_abi_version": [],
    "lit_device_info": [_pi, _pi, _pi, _psz, _psz],
    "lit_gemm_set_sm_limit": [_i],
    "lit_gemm_tf32x3_nt": [_vp, _vp, _l, _vp, _vp, _l, _i, _i, _i, _f, _vp, _l, _f, _vp, _vp, _l, _i, _vp],
    "lit_gemm_tf32x3_nt_corr": [_vp, _vp, _l, _vp, _vp, _l, _i, _i, _i, _i, _vp, _l, _vp, _vp, _l, _i, _vp],
    "lit_gemm_f16x3_nt_corr": [_vp, _vp, _l, _vp, _vp, _l, _i, _i, _i, _i, _vp, _l, _vp, _vp, _l, _i, _vp],
    "lit_gemm_f16x3_nt": [_vp, _vp, _l, _vp, _vp, _l, _i, _i, _i, _f, _vp, _l, _f, _vp, _vp, _l, _vp, _vp, _i, _vp],
    "lit_gemm_corr_series": [_i, _vp, _vp, _l, _vp, _vp, _l, _i, _i, _i, _i, _i, _vp, _l, _vp, _vp, _vp, _l, _i, _vp],
    "lit_split_f16": [_vp, _vp, _l, _l, _l, _l, _vp, _vp, _l, _vp, _vp, _vp],
    "lit_gather_col_reduce": [_vp, _l, _vp, _l, _l, _vp, _vp, _vp],
    "lit_row_absmax": [_vp, _l, _l, _l, _vp, _vp],
    "lit_f16_bound_scales": [_vp, _vp, _vp, _l, _l, _vp, _vp, _vp],
    "lit_gather_rows_transpose_f16": [_vp, _l, _vp, _l, _l, _vp, _vp, _vp, _l, _vp],
    "lit_gemm_f16x3_nt_pairout": [_vp, _vp, _l, _vp, _vp, _l, _i, _i, _i, _f, _vp, _l, _f, _vp, _l, _vp, _vp, _vp, _vp,
                                  _vp, _l, _i, _vp],
    "lit_convert_f64_to_f32": [_vp, _vp, _sz, _vp],
    "lit_convert_f32_to_f64": [_vp, _vp, _sz, _vp],
    "lit_split_tf32": [_vp, _l, _l, _l, _vp, _vp, _l, _vp],
    "lit_transpose_f32": [_vp, _l, _l, _l, _vp, _vp, _l, _vp],
    "lit_gather_rows_f32": [_vp, _l, _vp, _l, _l, _vp, _vp, _l, _l, _vp],
    "lit_gather_rows_transpose_split": [_vp, _l, _vp, _l, _l, _vp, _vp, _l, _vp],
    "lit_axpy_f32": [_f, _vp, _vp, _l, _vp, _l, _l, _l, _vp],
    "lit_fill_f32": [_vp, _sz, _f, _vp],
    "lit_memcpy_2d": [_vp, _sz, _vp, _sz, _sz, _sz, _i, _vp],
    "lit_host_pointer_kind": [_vp],
    "lit_col_stats": [_vp, _l, _vp, _l, _l, _i, _vp, _vp, _vp, _vp],
    "lit_gather_normalize_rows": [_vp, _l, _vp, _l, _l, _vp, _vp, _i, _f, _vp, _vp, _l, _l, _vp],
    "lit_syevd_workspace": [_i, _i, _i, _psz, _psz],
    "lit_syevd": [_vp, _i, _l, _i, _i, _vp, _vp, _sz, _vp, _sz, _vp, _vp],
    "lit_build_alpha_stack": [_vp, _l, _l, _l, _i, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _l, _vp],
    "lit_scale_rows_by_alpha": [_vp, _vp, _l, _l, _i, _vp, _vp, _i, _f, _vp, _vp, _l, _vp],
    "lit_corr_finalize": [_vp, _vp, _l, _i, _i, _l, _l, _f, _i, _i, _vp, _vp, _l, _vp],
    "lit_corr_finalize_scaled": [_vp, _vp, _l, _i, _i, _l, _l, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _l, _vp],
    "lit_corr_finalize_series": [_vp, _l, _i, _l, _l, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _l, _vp],
    "lit_argmax_alpha": [_vp, _l, _i, _l, _i, _vp, _vp, _vp, _vp, _vp],
    "lit_pearson_finalize": [_vp, _vp, _l, _i, _l, _l, _i, _vp, _vp, _vp],
    "lit_bh_workspace": [_l, _psz],
    "lit_bh_fdr": [_vp, _l, _d, _vp, _vp, _vp, _vp, _sz, _vp],
    "lit_fisher_combine": [_vp, _l, _i, _l, _i, _vp, _vp],
    "lit_fir_make_delayed": [_vp, _i, _l, _l, _l, _vp, _i, _i, _vp, _l, _vp],
    "lit_fir_zscore_rows": [_vp, _i, _l, _l, _l, _vp, _i, _i, _l, _l, _i, _vp, _l, _vp],
    "lit_lanczos_downsample": [_vp, _i, _l, _l, _l, _vp, _vp, _l, _d, _d, _i, _vp, _vp, _vp, _l, _vp],
    "lit_lanczos_lambda_max": [_vp, _l, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "lit_lanczos_lambda_max_batched": [_vp, _i, _l, _i, _i, _vp, _vp, _vp, _vp],
    "lit_gemm_tf32x3_nt_batched": [_vp, _vp, _l, _l, _vp, _vp, _l, _l, _i, _i, _i, _f, _vp, _l, _l, _f, _vp, _vp, _l, _l,
                                   _i, _i, _vp],
    "lit_gemm_tf32x3_nt_grouped": [_vp, _vp, _l, _vp, _vp, _l, _l, _i, _i, _i, _i, _vp, _vp, _vp, _l, _vp],
    "lit_group_plan": [_vp, _l, _i, _i, _l, _vp, _vp, _vp, _vp],
    "lit_spd_solve_workspace": [_i, _i, _i, _psz, _psz, _psz, _pl, _pl],
    "lit_spd_solve_batched": [_i, _i, _i, _vp, _l, _vp, _l, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "lit_spd_probe_residual": [_i, _i, _vp, _l, _vp, _l, _vp, _vp, _vp, _l, _l, _vp, _vp, _vp],
    "lit_cheb_update": [_vp, _vp, _vp, _vp, _vp, _vp, _l, _l, _l, _f, _f, _f, _i, _vp],
    "lit_poly_combine": [_vp, _vp, _i, _l, _l, _l, _l, _vp, _vp, _i, _vp, _vp, _l, _vp],
    "lit_series_stack": [_vp, _vp, _l, _l, _l, _vp, _l, _vp, _vp, _l, _vp],
    "lit_sinc_downsample": [_vp, _i, _l, _l, _l, _vp, _vp, _l, _d, _d, _i, _i, _vp, _vp, _vp, _l, _vp],
    "lit_csr_rows_apply": [_vp, _i, _l, _l, _vp, _vp, _vp, _l, _i, _vp, _l, _vp],
    "lit_gabor_downsample": [_vp, _i, _l, _l, _l, _vp, _vp, _l, _vp, _i, _d, _vp, _l, _vp],
}
_RESTYPES = {"lit_last_error": C.c_char_p}

GEMM_AUTO, GEMM_1CTA_N256, GEMM_1CTA_N128, GEMM_2CTA_N256 = 0, 1, 2, 3


class LitRidgeError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load liblitridge.so (once) and attach prototypes.  Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C litcoder_core_b200/csrc`.  litcoder_core_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # stale build
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    got = lib.lit_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI version {got}, expected {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().lit_last_error()
        raise LitRidgeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")
