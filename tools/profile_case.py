"""Small fixed workload for ncu captures: nx=160, grid 2, a handful of xi spanning the bench range,
NLO assembly + one RK4 interval, then a few applications.  Prints stage times and QAGS statistics."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiktaalik_b200 as tk
from tiktaalik_b200 import _lib

nxi = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 160
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
xi = np.geomspace(2e-5, 0.5, nxi)
L = _lib.lib()
for _ in range(reps):
    tk.matrices.initialize_kernels(nx, xi, grid_type=2)
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 10.0]), nlo=True)
ms = (C.c_double * 3)()
_lib.check(L.tkb_stage_ms(ms))
a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
_lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
last = _lib.view_tensor(b.value, (nxi, nx, 7, 3), torch.int32).cpu().numpy().astype(np.int64)
ifail = _lib.view_tensor(c.value, (nxi, nx, 7, 3), torch.int32).cpu().numpy()
neval = 42 * last - 21
print("stage ms: K1 %.2f  K2 %.2f  K3K4 %.2f" % tuple(ms))
print("QAGS evals per constant:", neval.sum(axis=(0, 1, 3)).tolist(), "total", int(neval.sum()))
print("mean last per constant/sub:", np.round(last.mean(axis=(0, 1)), 1).tolist())
print("max last:", int(last.max()), " instances at limit:", int((last >= 500).sum()), " ifail!=0:", int((ifail != 0).sum()),
      "of", last.size)
print("evals per xi:", neval.sum(axis=(1, 2, 3)).tolist())
M = tk.matrices.matrix_VSG()
g = torch.randn(2 * nx, nxi, 4, dtype=torch.float64, device="cuda")
out = tk.matrices.apply_matrix(M, g)
torch.cuda.synchronize()
print("launches", L.tkb_launch_count())
