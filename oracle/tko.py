"""ctypes binding of the CPU oracle (oracle/libtko.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from tiktaalik_b200/.
The oracle is a plain-C restatement of the reference algorithm (see oracle/tko.h);
this module only builds it on demand and exposes numpy-friendly wrappers whose
names follow the reference's Fortran/Python API.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_FAST = False


def use_fast_build(on=True):
    """bench.py's CPU legs only: switch to libtko_fast.so (libm log / pow, -Ofast -march=native: the reference's own
    build flags) for TIMING.  Never used as a checker: its numbers differ from the oracle's in the last bits."""
    global _LIB, _FAST
    if bool(on) != _FAST:
        _FAST = bool(on)
        _LIB = None

KERNEL_NAMES = [
    "K0_qq_pls", "KA0_qG_reg", "KVmA0_qG_reg", "KA0_Gq_reg", "KVmA0_Gq_reg", "KA0_GG_pls", "KVmA0_GG_reg",
    "KV1_NSp_pls", "KV1_NSm_pls", "KV1_NS_pls_nfl",
    "KV1_qq_reg", "KA1_qq_reg", "KV1_qG_reg", "KA1_qG_reg",
    "KV1_Gq_reg", "KA1_Gq_reg", "KV1_Gq_reg_nfl", "KA1_Gq_reg_nfl",
    "KV1_GG_pls", "KA1_GG_pls", "KV1_GG_pls_nfl", "KA1_GG_pls_nfl", "KV1_NSm_pls_nfl"]
CST_NAMES = ["KV1_NSp", "KV1_NSm", "KV1_NS_nfl", "KV1_GG", "KA1_GG", "KV1_GG_nfl", "KA1_GG_nfl"]

F1 = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile libtko.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libtko_fast.so" if _FAST else "libtko.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        d, i = C.c_double, C.c_int
        sigs = {
            "tko_dilog": (d, [d]), "tko_log_x": (d, [d]), "tko_log_cr_x": (d, [d]), "tko_pow_x": (d, [d, d]), "tko_theta_step": (d, [d]), "tko_rho_step": (d, [d, d]),
            "tko_erbl_step": (d, [d]), "tko_abslog": (d, [d]), "tko_log2": (d, [d]), "tko_logprod": (d, [d, d]),
            "tko_kernel_by_index": (d, [i, d, d, d]),
            "tko_cst_integral": (d, [i, d, d, _ip, _ip]),
            "tko_push_forward": (d, [d, d, i, i]), "tko_pull_back": (d, [d, d, i, i]),
            "tko_push_jacob": (d, [d, d, i, i]),
            "tko_initialize_lagrange_weights": (None, [i, i]),
            "tko_interpixel": (d, [i, i, d, d, i]),
            "tko_pixelspace": (None, [i, d, i, _dp]), "tko_grid_node": (d, [i, i, d, i]), "tko_set_variant": (None, [i, i]),
            "tko_qk21": (None, [F1, C.c_void_p, d, d, _dp, _dp, _dp, _dp]),
            "tko_qags": (None, [F1, C.c_void_p, d, d, d, d, _dp, _dp, _ip, _ip, _ip]),
            "tko_get_alpha_qcd": (d, [d]), "tko_get_neff": (i, [d]),
            "tko_make_kernels": (i, [i, i, _dp, i, i]), "tko_make_matrices": (i, [i, _dp, i]),
            "tko_get_nx": (i, []), "tko_get_nxi": (i, []), "tko_get_nQ2": (i, []),
            "tko_evomat": (i, [i, i, _dp]), "tko_kernel": (i, [i, d, i, i, i, _dp]),
            "tko_raw_kernel": (i, [i, i, _dp]), "tko_free_all": (None, []),
            "tko_gk": (d, [i, d, d, d]), "tko_gk_grid": (None, [i, i, i, i, _dp, _dp, _dp, _dp]),
        }
        for name in ["K0_qq_cst", "KA0_GG_cst", "KA0_GG_nfl", "KV1_NSp_cst", "KV1_NSm_cst", "KV1_NSp_cst_nfl",
                     "KV1_NSm_cst_nfl", "KV1_GG_cst", "KA1_GG_cst", "KV1_GG_cst_nfl", "KA1_GG_cst_nfl"]:
            sigs["tko_" + name] = (d, [d, d])
        for name, (res, args) in sigs.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_dp)


# ---- scalar helpers -------------------------------------------------------------------------
def dilog(z):
    return lib().tko_dilog(float(z))


def kernel(name_or_idx, x, y, xi):
    idx = KERNEL_NAMES.index(name_or_idx) if isinstance(name_or_idx, str) else int(name_or_idx)
    return lib().tko_kernel_by_index(idx, float(x), float(y), float(xi))


def cst_integral(name_or_idx, x, xi):
    """plus-prescription integral of one NLO constant + (last, ifail) of its 3 QAGS calls"""
    idx = CST_NAMES.index(name_or_idx) if isinstance(name_or_idx, str) else int(name_or_idx)
    last = (C.c_int * 3)()
    ifail = (C.c_int * 3)()
    v = lib().tko_cst_integral(idx, float(x), float(xi), last, ifail)
    return v, list(last), list(ifail)


def cst(name, x, xi):
    return getattr(lib(), "tko_" + name)(float(x), float(xi))


def qk21(f, a, b):
    cb = F1(lambda y, ctx: float(f(y)))
    r = [C.c_double() for _ in range(4)]
    lib().tko_qk21(cb, None, a, b, *[C.byref(v) for v in r])
    return tuple(v.value for v in r)


def qags(f, a, b, epsabs=1e-12, epsrel=1e-6):
    cb = F1(lambda y, ctx: float(f(y)))
    res, err = C.c_double(), C.c_double()
    neval, ifail, last = C.c_int(), C.c_int(), C.c_int()
    lib().tko_qags(cb, None, a, b, epsabs, epsrel, C.byref(res), C.byref(err),
                   C.byref(neval), C.byref(ifail), C.byref(last))
    return res.value, err.value, neval.value, ifail.value, last.value


def alphaQCD(Q2):
    return np.array([lib().tko_get_alpha_qcd(float(q)) for q in np.atleast_1d(Q2)])


def pixelspace(nx, xi=0.5, grid_type=1):
    x = np.empty(nx)
    lib().tko_pixelspace(nx, float(xi), grid_type, _ptr(x))
    return x


def set_variant(eta_recip_f32=False, x0_recip_f32=False):
    """-Ofast float32-division variants of the node formula / stencil origin; (False, False) is canonical."""
    lib().tko_set_variant(int(bool(eta_recip_f32)), int(bool(x0_recip_f32)))


def grid_nodes(nx, xi, grid_type):
    """the x_i the kernels are evaluated at (convolution.f90:45-46; differs from pixelspace by <=3e-8)"""
    return np.array([lib().tko_grid_node(i, nx, float(xi), grid_type) for i in range(1, nx + 1)])


# ---- the matrices API (tiktaalik/matrices.py names) ------------------------------------------
def initialize_kernels(nx, xi, grid_type=1, mode=1):
    xi = _arr(np.atleast_1d(xi))
    rc = lib().tko_make_kernels(nx, xi.shape[0], _ptr(xi), grid_type, mode)
    if rc:
        raise RuntimeError("tko_make_kernels failed")


def initialize_evolution_matrices(Q2, nlo=False):
    Q2 = _arr(Q2)
    assert Q2.shape[0] >= 2
    rc = lib().tko_make_matrices(Q2.shape[0], _ptr(Q2), int(bool(nlo)))
    if rc:
        raise RuntimeError("tko_make_matrices failed")


def _evomat(kind, nstype=1):
    L = lib()
    nx, nxi, nq = L.tko_get_nx(), L.tko_get_nxi(), L.tko_get_nQ2()
    m = nx * (2 if kind in (1, 3) else 1)
    M = np.empty((m, m, nxi, nq), order="F")
    if L.tko_evomat(kind, nstype, _ptr(M)):
        raise RuntimeError("tko_evomat failed")
    M[np.isnan(M)] = 0
    M[np.isinf(M)] = 0
    return M


def matrix_VNS(ns_type=1):
    return _evomat(0, ns_type)


def matrix_VSG():
    return _evomat(1)


def matrix_ANS(ns_type=1):
    return _evomat(2, ns_type)


def matrix_ASG():
    return _evomat(3)


_KTYPES = {"VQQ": 0, "VQG": 1, "VGQ": 2, "VGG": 3, "AQQ": 4, "AQG": 5, "AGQ": 6, "AGG": 7}


def kernel_matrix(kind, Q2=1.29 ** 2, nfl=4, nlo=False, ns_type=1):
    L = lib()
    nx, nxi = L.tko_get_nx(), L.tko_get_nxi()
    K = np.empty((nx, nx, nxi), order="F")
    if L.tko_kernel(_KTYPES[kind], float(Q2), nfl, int(bool(nlo)), ns_type, _ptr(K)):
        raise RuntimeError("tko_kernel failed")
    return K


RAW_NAMES = ["K_NS_0", "KV_NS_1p", "KV_NS_1m", "KV_SG_0", "KA_SG_0", "KV_SG_1", "KA_SG_1"]


def raw_kernel(which, nfl=4):
    L = lib()
    idx = RAW_NAMES.index(which) if isinstance(which, str) else int(which)
    nx, nxi = L.tko_get_nx(), L.tko_get_nxi()
    m = nx * (2 if idx >= 3 else 1)
    K = np.empty((m, m, nxi), order="F")
    if L.tko_raw_kernel(idx, nfl, _ptr(K)):
        raise RuntimeError("tko_raw_kernel failed")
    return K


# ---- GK model (tiktaalik/model.py names) ----------------------------------------------------
_GK = {"Hu": 0, "Hd": 1, "Hs": 2, "Hg": 3, "Hu_val": 4, "Hd_val": 5, "Hu_sea": 6, "Hd_sea": 7}


def gk(name, x, xi, t):
    x, xi, t = (_arr(np.atleast_1d(v)) for v in (x, xi, t))
    f = np.empty((x.shape[0], xi.shape[0], t.shape[0]), order="F")
    lib().tko_gk_grid(_GK[name], x.shape[0], xi.shape[0], t.shape[0], _ptr(x), _ptr(xi), _ptr(t), _ptr(f))
    return f


def H_singlet(x, xi, t):
    x = _arr(np.atleast_1d(x))
    return (gk("Hu", x, xi, t) - gk("Hu", -x, xi, t) + gk("Hd", x, xi, t) - gk("Hd", -x, xi, t)
            + gk("Hs", x, xi, t) - gk("Hs", -x, xi, t))
