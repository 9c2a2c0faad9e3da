"""Workload for the calibration of the ALGORITHMIC flop count of one QAGS integrand evaluation, per constant.

Runs tkb_eval_integrand (the straight evaluation of K(x,y,xi) - K(y,x,xi) from the reference formulas: four half
kernels in full, nothing skipped, no tables) for each of the seven constants on NPTS points: x = grid nodes of the C3
grid (nx = 160, grid 2) at xi drawn from the benchmark's list, y uniform in (-1, 1) with a third of the points inside
|y| < 1.5 xi.  Run it under
    ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum -k regex:eval_integrand_kernel --csv --log-file <csv> python tools/calibrate_k2_flops.py
and feed the CSV to tools/ncu_summary.py --calibration, which writes profiles/k2_flops_per_eval.json
(flop = dadd + dmul + 2 dfma thread instructions / NPTS)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiktaalik_b200 as tk

NPTS = 1 << 20
rng = np.random.default_rng(2025)
xis = np.geomspace(2e-5, 0.5, 256)
xi = xis[rng.integers(0, 256, NPTS)]
i = rng.integers(1, 161, NPTS)
eta = (2 * i - 1) / 160.0 - 1.0
x = np.where(eta < -0.5, -(xi * xi) ** (1 + eta), np.where(eta <= 0.5, 2 * xi * eta, (xi * xi) ** (1 - eta)))
y = rng.uniform(-1, 1, NPTS)
third = NPTS // 3
y[:third] *= 1.5 * xi[:third]
L = tk._lib.lib()
xd, yd, xid = (torch.as_tensor(a).cuda() for a in (x, y, xi))
out = torch.empty(NPTS, dtype=torch.float64, device="cuda")
for c in range(7):
    tk._lib.check(L.tkb_eval_integrand(c, NPTS, tk._lib.ptr(xd), tk._lib.ptr(yd), tk._lib.ptr(xid), tk._lib.ptr(out),
                                       tk._lib.stream_ptr()))
torch.cuda.synchronize()
print("NPTS", NPTS, "finite fraction of the last constant", float(torch.isfinite(out).double().mean()))
