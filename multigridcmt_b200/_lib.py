"""ctypes binding of libmgcmt_b200.so (the C ABI in include/mgcmt_b200.h).

The product path has no CPU fallback: if the library is missing, or no CUDA device is present
when a compute entry point is called, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgcmt_b200.so")

OK = 0
SMOOTH_WJACOBI, SMOOTH_RBGS, SMOOTH_GSLEX = 0, 1, 2

_lib = None


class MgcmtError(RuntimeError):
    pass


_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_LL = C.c_longlong

# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header
SIGNATURES = {
    "mgcmt_abi_version": (_I, []),
    "mgcmt_last_error": (C.c_char_p, []),
    "mgcmt_launch_count": (_LL, []),
    "mgcmt_profile_enable": (_I, [_I]),
    "mgcmt_profile_read": (_I, [C.POINTER(_D), C.POINTER(_LL)]),
    "mgcmt_profile_read_kinds": (_I, [C.POINTER(_D), C.POINTER(_LL)]),
    "mgcmt_hier_create": (_I, [C.POINTER(_P), _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P]),
    "mgcmt_hier_create2": (_I, [C.POINTER(_P), _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "mgcmt_hier_create_slab": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "mgcmt_hier_level_buffers": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "mgcmt_vcycle_rq": (_I, [_P, _D, _I, _I, _I, _D, _P, _P, _I, _P, _P]),
    "mgcmt_vcycle_host_block": (_I, [_P, _I, _P, _I, _I, _I, _D, _P, _P, _P]),
    "mgcmt_vcycle_from": (_I, [_P, _I, _D, _I, _D, _P, _P, _P]),
    "mgcmt_slab_rayleigh": (_I, [_P, _I, _P, _P, _P]),
    "mgcmt_slab_up_rq": (_I, [_P, _I, _D, _D, _P, _P, _P, _P, _P, _P]),
    "mgcmt_hier_destroy": (_I, [_P]),
    "mgcmt_hier_num_levels": (_I, [_P]),
    "mgcmt_hier_level_shape": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I)]),
    "mgcmt_hier_level_coefs": (_I, [_P, _I, _P, _P]),
    "mgcmt_apply": (_I, [_P, _I, _D, _P, _P, _P]),
    "mgcmt_apply_mass": (_I, [_P, _I, _P, _P, _P]),
    "mgcmt_axpby": (_I, [_LL, _D, _P, _D, _P, _P, _P]),
    "mgcmt_residual": (_I, [_P, _I, _D, _P, _P, _P, _P]),
    "mgcmt_smooth": (_I, [_P, _I, _I, _D, _D, _I, _P, _P, _P, _P]),
    "mgcmt_restrict": (_I, [_P, _I, _P, _P, _P]),
    "mgcmt_residual_restrict": (_I, [_P, _I, _D, _P, _P, _P, _P]),
    "mgcmt_prolong": (_I, [_P, _I, _P, _P, _P]),
    "mgcmt_prolong_correct": (_I, [_P, _I, _P, _P, _P]),
    "mgcmt_coarse_solve": (_I, [_P, _D, _P, _P, _P]),
    "mgcmt_vcycle": (_I, [_P, _D, _I, _I, _I, _D, _P, _P, _I, _P]),
    "mgcmt_fused_leg": (_I, [_P, _I, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P]),
    "mgcmt_set_option": (_I, [C.c_char_p, _I]),
    "mgcmt_debug_uni_coefficients": (_I, [_D, _D, _D, _D, C.POINTER(_D)]),
    "mgcmt_debug_leg_rows_per_chunk": (_I, [_I, _I, _I, _I, _I]),
    "mgcmt_debug_slab_phases": (_I, [_I, _I, _I, _P, _P, _I]),
    "mgcmt_debug_poison_shared_memory": (_I, [_P]),
    "mgcmt_dot": (_I, [_LL, _P, _P, _P, _P]),
    "mgcmt_rayleigh": (_I, [_P, _I, _P, _P, _P]),
    "mgcmt_normalize": (_I, [_LL, _P, _P]),
    "mgcmt_gram": (_I, [_LL, _I, _P, _LL, _P, _P]),
    "mgcmt_cholqr_apply": (_I, [_LL, _I, _P, _LL, _P, _P]),
    "mgcmt_scale_inv_norm": (_I, [_LL, _P, _P, _P]),
    "mgcmt_axpy_dev": (_I, [_LL, _P, _D, _P, _P, _P]),
    "mgcmt_eigen_residual": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "mgcmt_ortho_status": (_I, [C.POINTER(_I), _P]),
    "mgcmt_rqmin": (_I, [_P, _I, _I, _P, _I, _P, _LL, _P, _P]),
    "mgcmt_rqmin_work_doubles": (_LL, [_P, _I, _I]),
    "mgcmt_gramschmidt": (_I, [_LL, _I, _P, _I, _P]),
    "mgcmt_band_create": (_I, [_I, _I, _P, _P, _I, _P, C.POINTER(_P)]),
    "mgcmt_band_destroy": (_I, [_P]),
    "mgcmt_band_num_levels": (_I, [_P, C.POINTER(_I)]),
    "mgcmt_band_level_shape": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I)]),
    "mgcmt_band_level_diags": (_I, [_P, _I, _P, _P]),
    "mgcmt_band_apply": (_I, [_P, _I, _D, _P, _P, _P]),
    "mgcmt_band_smooth": (_I, [_P, _I, _I, _I, _D, _D, _P, _P, _P]),
    "mgcmt_band_residual_restrict": (_I, [_P, _I, _D, _P, _P, _P, _P]),
    "mgcmt_band_prolong_correct": (_I, [_P, _I, _P, _P, _P]),
    "mgcmt_band_coarse_solve": (_I, [_P, _D, _P, _P, _P]),
    "mgcmt_band_vcycle": (_I, [_P, _D, _I, _I, _I, _D, _P, _P, _P]),
    "mgcmt_nccl_load": (_I, [C.c_char_p]),
    "mgcmt_nccl_unique_id": (_I, [_P]),
    "mgcmt_nccl_comm_create": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "mgcmt_nccl_comm_destroy": (_I, [_P]),
    "mgcmt_slabblock_create": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _D, _P, C.POINTER(_P)]),
    "mgcmt_slabblock_destroy": (_I, [_P]),
    "mgcmt_slabblock_cycle": (_I, [_P, _P, _P, _P, _P, _P]),
    "mgcmt_slabblock_gram": (_I, [_P, _P, _LL, _P]),
    "mgcmt_slabblock_profile": (_I, [_P, _I]),
    "mgcmt_slabblock_set_smoother": (_I, [_P, _I, _D]),
    "mgcmt_slabblock_profile_read": (_I, [_P, _I, _P, _P, C.POINTER(_I)]),
}


def load():
    """Load the shared library (no GPU needed for this) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MgcmtError(
            "libmgcmt_b200.so is not built (%s). Run `python -m multigridcmt_b200.build`; there is no "
            "CPU fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().mgcmt_last_error()
        raise MgcmtError("libmgcmt_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise MgcmtError("no CUDA device: the multigrid path runs on the GPU only (no CPU fallback)")
    return torch
