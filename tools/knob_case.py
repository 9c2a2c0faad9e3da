"""One small NLO case (nx = 40, 3 xi, grid 2, three Q2 points) run by the library in THIS process -- i.e. under
whatever TKB_* switches the environment carries (they are read once per process) -- and checked against the oracle:
raw kernel matrices and QAGS paths bit for bit, evolution matrices and an application within 1e-9 / 1e-12.
Prints 'OK' or the first mismatch.  Used by tests/test_gpu_knobs.py, one subprocess per switch."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiktaalik_b200 as tk
from oracle import tko


def main():
    nx, gt = 40, 2
    xi = np.array([3e-4, 0.07, 0.62])
    Q2 = np.array([4.0, 12.0, 40.0])
    tk.matrices.initialize_kernels(nx, xi, grid_type=gt)
    tko.initialize_kernels(nx, xi, gt, mode=1)
    L = tk._lib.lib()
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    tk._lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
    res = tk._lib.view_tensor(a.value, (3, nx, 7, 3)).cpu().numpy()
    last = tk._lib.view_tensor(b.value, (3, nx, 7, 3), torch.int32).cpu().numpy()
    for k in range(3):
        xs = tko.grid_nodes(nx, xi[k], gt)
        for i in range(0, nx, 3):
            for cst in range(7):
                v, l, _ = tko.cst_integral(cst, xs[i], xi[k])
                tot = ((0.0 + res[k, i, cst, 0]) + res[k, i, cst, 1]) + res[k, i, cst, 2]
                if list(last[k, i, cst]) != l or not (tot == v or (np.isnan(tot) and np.isnan(v))):
                    return "QAGS constant %d at xi[%d], node %d differs" % (cst, k, i)
    for nfl in (3, 4, 5):
        for which, name in enumerate(tko.RAW_NAMES):
            addr = C.c_void_p()
            tk._lib.check(L.tkb_raw_kernel(which, nfl, C.byref(addr), tk._lib.stream_ptr()))
            torch.cuda.synchronize()
            m = nx * (2 if which >= 3 else 1)
            got = tk._lib.view_tensor(addr.value, (3, m, m)).permute(1, 2, 0).cpu().numpy()
            if not np.array_equal(got, tko.raw_kernel(name, nfl), equal_nan=True):
                return "raw kernel %s nfl=%d not bit-identical" % (name, nfl)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    tko.initialize_evolution_matrices(Q2, nlo=True)
    g = np.random.default_rng(3).random((2 * nx, 3, 5))
    for name, kw in (("matrix_VNS", {"ns_type": 1}), ("matrix_VNS", {"ns_type": -1}), ("matrix_VSG", {}), ("matrix_ASG", {})):
        got = getattr(tk.matrices, name)(**kw)
        ref = getattr(tko, name)(**kw)
        if not np.allclose(got.cpu().numpy(), ref, rtol=1e-9, atol=1e-12):
            return "%s %s outside 1e-9 / 1e-12" % (name, kw)
        gg = g[: ref.shape[0]]
        for S in (1, 5):
            out = tk.matrices.apply_matrix(got, gg[:, :, :S]).cpu().numpy()
            if not np.allclose(out, np.einsum('ijkl,jks->iksl', ref, gg[:, :, :S]), rtol=1e-9, atol=1e-12):
                return "application of %s, S=%d, outside 1e-9 / 1e-12" % (name, S)
    return "OK"


if __name__ == "__main__":
    print(main())
