"""ctypes binding of ``lib/libb200bo.so`` — the C ABI declared in include/b200bo.h.

Loading never needs a GPU (symbols only); calling does.  A missing library is a hard error:
the product has no CPU path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int32, c_longlong, c_size_t, c_uint64, c_void_p

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("B200BO_LIB_PATH") or os.path.join(PKG_DIR, "lib", "libb200bo.so")   # override: kernel A/B experiments

P = c_void_p   # every device / host pointer crosses the ABI as void*

# name -> (restype, argtypes); mirrors include/b200bo.h one to one
PROTOTYPES = {
    "b200bo_version": (c_char_p, []),
    "b200bo_last_error_string": (c_char_p, []),
    "b200bo_device_info": (c_int, [P, P, P]),
    "b200bo_dpp_logdet_score": (c_int, [P, c_int, P, c_longlong, c_int, c_double, P, P]),
    "b200bo_dpp_table": (c_int, [c_int, c_int, c_double, c_double, P, P]),
    "b200bo_dpp_logdet_score_lattice": (c_int, [P, c_int, c_int, P, c_longlong, c_int, P, P]),
    "b200bo_dpp_stats_workspace_bytes": (c_size_t, [c_longlong]),
    "b200bo_dpp_rank_workspace_bytes": (c_size_t, [c_longlong]),
    "b200bo_dpp_rank": (c_int, [P, c_longlong, P, c_int, P, c_int, P, P, P, P, P, c_size_t, P]),
    "b200bo_dpp_count_less": (c_int, [P, c_longlong, c_double, P, P]),
    "b200bo_dpp_window_pick": (c_int, [P, c_longlong, c_int, c_int, c_int, P, c_int, P, c_int, P, P, P]),
    "b200bo_dpp_logdet_score_multigamma": (c_int, [P, c_int, P, c_longlong, c_int, P, c_int, P, P]),
    "b200bo_dpp_zscore": (c_int, [P, c_longlong, P, P, P, c_size_t, P]),
    "b200bo_unrank_u128": (c_int, [P, P, P, P, c_uint64, c_uint64, c_int, c_int, c_longlong, P, P]),
    "b200bo_mt_seeded_randbelow": (c_int, [P, c_longlong, c_uint64, c_uint64, P, P, P, P]),
    "b200bo_wildcat_surface": (c_int, [P, P, P, P, P, P, c_int, P]),
    "b200bo_fp64_fma_probe": (c_longlong, [P, c_int, c_int, P]),
    "b200bo_smem_lds_probe": (c_longlong, [P, c_int, c_int, P]),
    "b200bo_wildcat_eval": (c_int, [P, P, P, P, P, c_longlong, P]),
    "b200bo_simplex_noise": (c_int, [c_uint64, P, P, P, c_longlong, P]),
    "b200bo_wfrag_doubles": (c_size_t, [c_int]),
    "b200bo_gp_assemble_chol": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "b200bo_gp_chol_append": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "b200bo_gp_zvec": (c_int, [P, P, P, P, P, c_int, c_int, P]),
    "b200bo_grid_acq_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200bo_grid_acq_chunks": (c_int, [c_int, c_int, c_int, c_int]),
    "b200bo_grid_posterior_acq_argmax": (c_int, [P, c_int, P, P, P, P, P, P, P, c_int, c_int, c_double, c_double,
                                                 P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "b200bo_kernel_table": (c_int, [P, c_int, c_int, c_double, c_int, c_int, P, P]),
    "b200bo_grid_posterior_acq_argmax_lattice": (c_int, [c_int, c_int, c_double, c_double, c_double, P, P, P, P, P, P, P, P,
                                                         P, c_int, c_int, c_double, c_double, P, P, P, P, P, P, c_int,
                                                         c_int, c_int, c_int, P, c_size_t, P]),
    "b200bo_kernel_table_signed": (c_int, [P, c_int, c_int, c_double, c_int, c_int, P, P]),
    "b200bo_grid_inc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200bo_grid_posterior_update_acq_argmax": (c_int, [c_int, c_int, c_double, c_double, c_double, P, P, P, P, P, P, P, P, P,
                                                        P, P, P, c_int, c_int, c_int, c_double, c_double, P, P, P, P,
                                                        c_int, c_int, P, c_size_t, P]),
    "b200bo_kernel_table_ysigned": (c_int, [P, c_int, c_int, c_double, c_int, c_int, P, P]),
    "b200bo_bo_fixed_loop_supported": (c_int, [c_int, c_int, c_int]),
    "b200bo_bo_fixed_loop_ctas": (c_int, [c_int]),
    "b200bo_bo_fixed_loop_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "b200bo_bo_fixed_loop": (c_int, [c_int, c_int, c_double, c_double, c_double, P, P, P, P, c_int, P, P, P, P, P, P, P, P,
                                     P, P, P, c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_double, P,
                                     c_size_t, P]),
    "b200bo_mll_value_grad": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "b200bo_lbfgs_state_bytes": (c_size_t, [c_int]),
    "b200bo_lbfgs_init": (c_int, [P, P, P, c_int, P, P, c_int, c_double, P]),
    "b200bo_lbfgs_step": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_double, c_double, c_double, c_int, c_int, P]),
    "b200bo_gp_fit": (c_int, [P, P, P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                              c_int, c_int, c_int, P]),
    "b200bo_gp_fit_ordered": (c_int, [P, P, P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                                      c_int, c_int, c_int, P, P]),
    "b200bo_gp_fit_large_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200bo_gp_fit_large": (c_int, [P, P, P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_double, c_double, c_double,
                                    c_int, c_int, c_int, P, c_size_t, P]),
    "b200bo_mll_value_grad_large": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, P, c_size_t, P]),
    "b200bo_bootstrap_means": (c_int, [P, c_int, c_int, P, c_int, P, P]),
    "b200bo_column_mean_percentiles": (c_int, [P, c_int, c_int, P, c_int, P, P]),
    "b200bo_bo_init": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_double, c_double, P]),
    "b200bo_bo_update": (c_int, [P, c_int, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_double,
                                 c_double, P]),
    "b200bo_bo_finalize": (c_int, [P, P, P, c_int, c_int, P]),
    "b200bo_bo_fit_counts": (c_int, [P, P, P, c_int, c_int, P]),
}

_lib = None


class B200boError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """dlopen the library (once) and attach prototypes. Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200boError(
            f"{LIB_PATH} is missing: build it with __graft_entry__.build() or `python -m b200bo.build`. "
            "There is no CPU fallback.")
    try:
        import torch  # noqa: F401  makes libcudart.so.12 resident so the shared runtime is the one torch uses
    except Exception:  # pragma: no cover
        pass
    try:
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    except OSError:
        for cand in ("/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            if os.path.exists(cand):
                ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
                break
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().b200bo_last_error_string().decode()
        kind = "argument error" if rc < 0 else f"CUDA error {rc}"
        raise B200boError(f"{what}: {kind}: {msg}")


def version() -> str:
    return load().b200bo_version().decode()
