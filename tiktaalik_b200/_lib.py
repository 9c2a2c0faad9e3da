"""ctypes loader for libtiktaalik_b200.so (C ABI in include/tiktaalik_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or no CUDA device
is available, every compute entry point raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("TKB_LIB", os.path.join(_HERE, "libtiktaalik_b200.so"))   # TKB_LIB: tuning builds only
_LIB = None

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)

# every symbol include/tiktaalik_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tkb_last_error": (C.c_char_p, []),
    "tkb_version": (C.c_int, []),
    "tkb_make_kernels": (C.c_int, [C.c_int, C.c_int, _dp, C.c_int, C.c_void_p]),
    "tkb_make_kernels_shard": (C.c_int, [C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tkb_make_matrices": (C.c_int, [C.c_int, _dp, C.c_int, C.c_void_p]),
    "tkb_get_nx": (C.c_int, []),
    "tkb_get_nxi": (C.c_int, []),
    "tkb_get_nq2": (C.c_int, []),
    "tkb_get_lnlo": (C.c_int, []),
    "tkb_evomatrix": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p), _i64p, _i64p]),
    "tkb_evomatrix_copy": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tkb_evokernel": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tkb_pixelspace": (C.c_int, [C.c_int, C.c_double, C.c_int, _dp]),
    "tkb_alpha": (C.c_int, [C.c_int, _dp, _dp]),
    "tkb_apply": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                            C.c_void_p, C.c_void_p]),
    "tkb_apply_strided": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
    "tkb_gk": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tkb_free": (C.c_int, []),
    "tkb_qags_info": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "tkb_raw_kernel": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    "tkb_eval_kernels": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tkb_eval_integrand": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tkb_eval_scalar": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tkb_stage_ms": (C.c_int, [_dp]),
    "tkb_launch_count": (C.c_longlong, []),
    "tkb_set_variant": (C.c_int, [C.c_int, C.c_int]),
    "tkb_task_classes": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tkb_evolve_workspace": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int]),
    "tkb_evolve_fixed_flavor": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, _dp, _dp,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p]),
    "tkb_shard_unique_id": (C.c_int, [C.c_char_p]),
    "tkb_shard_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    "tkb_allgather_matrices": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]),
    "tkb_allgather_matrices_strided": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tkb_shard_finalize": (C.c_int, []),
    "tkb_p2p_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "tkb_p2p_handle": (C.c_int, [C.c_int, C.c_char_p]),
    "tkb_p2p_open": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p]),
    "tkb_p2p_push_matrices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tkb_p2p_close": (C.c_int, []),
    "tkb_fp64_peak": (C.c_int, [_dp]),
}


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "tiktaalik_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C tiktaalik_b200/csrc`. There is no CPU fallback." % SO_PATH)
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise RuntimeError("tiktaalik_b200: " + lib().tkb_last_error().decode("utf-8", "replace"))


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("tiktaalik_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr())


class _DevView:
    """Zero-copy window on library-owned device memory (valid until the next make_kernels/make_matrices)."""

    def __init__(self, addr, shape, dtype="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": dtype,
                                         "data": (int(addr), False), "version": 2, "strides": None}


def view_tensor(addr, shape, dtype=torch.float64):
    typestr = {torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevView(addr, shape, typestr), device=require_cuda())
