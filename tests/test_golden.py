"""Committed golden vectors (tests/golden/, written by tests/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path reproduces them."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")) if "scalars" not in p)


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_fixtures_exist():
    assert len(CASES) >= 3 and os.path.exists(os.path.join(GOLD, "scalars.npz"))


def test_oracle_reproduces_scalar_fixtures(tko):
    g = _load("scalars")
    assert np.array_equal(np.array([tko.dilog(v) for v in g["z"]]), g["dilog"])
    got = np.array([[tko.kernel(f, *p) for f in range(22)] for p in g["pts"]])
    assert np.array_equal(got, g["kernels"], equal_nan=True)
    assert np.array_equal(tko.alphaQCD(g["alpha_q2"]), g["alpha"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_matrix_fixtures(tko, name):
    g = _load(name)
    tko.initialize_kernels(int(g["nx"]), g["xi"], int(g["grid_type"]), mode=1)
    tko.initialize_evolution_matrices(g["Q2"], nlo=bool(g["nlo"]))
    for nm in tko.RAW_NAMES:
        assert np.array_equal(tko.raw_kernel(nm, 4), g["raw4_" + nm], equal_nan=True), nm
    assert np.array_equal(tko.matrix_VNS(1), g["M_VNS_p"])
    assert np.array_equal(tko.matrix_VSG(), g["M_VSG"])
    assert np.array_equal(tko.matrix_ASG(), g["M_ASG"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_matrix_fixtures(tk, name):
    import torch
    g = _load(name)
    nx, nxi = int(g["nx"]), g["xi"].shape[0]
    tk.matrices.initialize_kernels(nx, g["xi"], grid_type=int(g["grid_type"]))
    tk.matrices.initialize_evolution_matrices(g["Q2"], nlo=bool(g["nlo"]))
    L = tk._lib.lib()
    for which, nm in enumerate(["K_NS_0", "KV_NS_1p", "KV_NS_1m", "KV_SG_0", "KA_SG_0", "KV_SG_1", "KA_SG_1"]):
        addr = C.c_void_p()
        tk._lib.check(L.tkb_raw_kernel(which, 4, C.byref(addr), tk._lib.stream_ptr()))
        torch.cuda.synchronize()
        m = nx * (2 if which >= 3 else 1)
        got = tk._lib.view_tensor(addr.value, (nxi, m, m)).permute(1, 2, 0).cpu().numpy()
        assert np.array_equal(got, g["raw4_" + nm], equal_nan=True), nm          # assembly is bit-exact
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    tk._lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
    last = tk._lib.view_tensor(b.value, (nxi, nx, 7, 3), torch.int32).cpu().numpy()
    ifail = tk._lib.view_tensor(c.value, (nxi, nx, 7, 3), torch.int32).cpu().numpy()
    assert np.array_equal(last, g["qags_last"]) and np.array_equal(ifail, g["qags_ifail"])   # same QUADPACK path
    for got, ref in ((tk.matrices.matrix_VNS(1), g["M_VNS_p"]), (tk.matrices.matrix_VNS(-1), g["M_VNS_m"]),
                     (tk.matrices.matrix_VSG(), g["M_VSG"]), (tk.matrices.matrix_ASG(), g["M_ASG"])):
        got = got.cpu().numpy()
        scale = np.abs(ref).max()
        assert np.all(np.abs(got - ref) <= 1e-9 * np.abs(ref) + 1e-12 * scale)
