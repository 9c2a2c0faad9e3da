"""Randomised parity sweep: kernel matrices (bit for bit) and QAGS paths of the library against the oracle on random
(nx, xi list, grid type) configurations.  Usage: python tools/stress_parity.py [cases] [seed]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiktaalik_b200 as tk
from oracle import tko



def run(cases, seed, verbose=True):
  rng = np.random.default_rng(seed)
  L = tk._lib.lib()
  bad = 0
  for case in range(cases):
      nx = int(rng.choice([6, 8, 10, 12, 16, 20, 24, 30, 36]))
      gt = int(rng.choice([1, 2]))
      nxi = int(rng.integers(1, 5))
      xi = np.sort(np.concatenate([10.0 ** rng.uniform(-5, -0.05, nxi - nxi // 2), rng.uniform(0.05, 0.999, nxi // 2)]))
      if gt == 2 and nx % 4:
          nx += 2
      tk.matrices.initialize_kernels(nx, xi, grid_type=gt)
      tko.initialize_kernels(nx, xi, gt, mode=1)
      xi_c = xi.copy()
      xi_c[0] = max(xi_c[0], 1e-6)
      xi_c[-1] = min(xi_c[-1], 0.999999)
      a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
      tk._lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
      last = tk._lib.view_tensor(b.value, (nxi, nx, 7, 3), torch.int32).cpu().numpy()
      ifail = tk._lib.view_tensor(c.value, (nxi, nx, 7, 3), torch.int32).cpu().numpy()
      npath = 0
      for k in range(nxi):
          xs = tko.grid_nodes(nx, xi_c[k], gt)
          for i in range(nx):
              for cst in range(7):
                  _, l, e = tko.cst_integral(cst, xs[i], xi_c[k])
                  if list(last[k, i, cst]) != l or list(ifail[k, i, cst]) != e:
                      npath += 1
      nmat = 0
      for nfl in (3, 4, 5):
          for which, name in enumerate(tko.RAW_NAMES):
              addr = C.c_void_p()
              tk._lib.check(L.tkb_raw_kernel(which, nfl, C.byref(addr), tk._lib.stream_ptr()))
              torch.cuda.synchronize()
              m = nx * (2 if which >= 3 else 1)
              got = tk._lib.view_tensor(addr.value, (nxi, m, m)).permute(1, 2, 0).cpu().numpy()
              if not np.array_equal(got, tko.raw_kernel(name, nfl), equal_nan=True):
                  nmat += 1
      if verbose: print("case %2d: nx=%2d grid=%d xi=%s  path mismatches %d, matrices not bit-identical %d, max last %d"
            % (case, nx, gt, np.array2string(xi, precision=3), npath, nmat, int(last.max())))
      bad += npath + nmat
  return bad


if __name__ == "__main__":
    total = run(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print("TOTAL mismatches:", total)
    sys.exit(1 if total else 0)
