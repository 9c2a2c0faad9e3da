"""Application benchmark (config C4 of BASELINE.json): stored NLO matrices nx=160, 128 xi, 16 Q2 applied to
S synthetic GPD samples.  Prints, per S and matrix type, time, algorithmic GB/s against the measured HBM
peak (MEASURED_PEAKS.json) and TFLOP/s against the measured DFMA peak, plus evolved GPDs/s."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiktaalik_b200 as tk
from tiktaalik_b200 import _lib

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 160
nxi = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 16
Slist = [int(s) for s in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 4, 8, 32, 256, 1024]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    hbm = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "measured"
except Exception:
    hbm, src = 6650.0, "fallback"
L = _lib.lib()
peak = C.c_double()
_lib.check(L.tkb_fp64_peak(C.byref(peak)))
xi = np.geomspace(1e-3, 0.9, nxi)
tk.matrices.initialize_kernels(nx, xi, grid_type=2)
tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 10.0, nq), nlo=True)
tk.f90wrap.matrices.dummy.zero_copy = True
mats = {"NS": tk.matrices.matrix_VNS(), "SG": tk.matrices.matrix_VSG()}
gen = torch.Generator(device="cuda").manual_seed(1234)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
print("HBM peak %.1f GB/s (%s), FP64 DFMA peak %.2f TFLOP/s (measured here)" % (hbm, src, peak.value))
for name, M in mats.items():
    m = M.shape[0]
    for S in Slist:
        g = torch.rand((m, nxi, S), generator=gen, dtype=torch.float64, device="cuda")
        out = tk.matrices.apply_matrix(M, g)
        reps = 5
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = tk.matrices.apply_matrix(M, g)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts)) * 1e-3
        byt = 8.0 * (m * m * nxi * nq + m * nxi * S + m * nxi * S * nq)
        flop = 2.0 * m * m * nxi * nq * S
        print("%s m=%d S=%5d  %8.3f ms  %8.1f GB/s (%.2f of HBM)  %6.2f TFLOP/s (%.2f of FP64)  %.3e evolved GPDs/s"
              % (name, m, S, t * 1e3, byt / t / 1e9, byt / t / 1e9 / hbm, flop / t / 1e12, flop / t / 1e12 / peak.value, S / t))
# config C4 at its full sample count: 10^5 samples, applied in blocks that are read in place, results rotating
# through two buffers (the 262 / 1049 GB result never exists as a whole); one timed pass over all samples
S_total = int(os.environ.get("TKB_C4_SAMPLES", "100000"))
for name, M in mats.items():
    m = M.shape[0]
    g_all = torch.rand((m, nxi, S_total), generator=gen, dtype=torch.float64, device="cuda")
    for chunk in (1024, 2048):
        for _ in tk.matrices.apply_matrix_chunks(M, g_all[:, :, : 2 * chunk], chunk=chunk):
            pass
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in tk.matrices.apply_matrix_chunks(M, g_all, chunk=chunk):
            pass
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        flop = 2.0 * m * m * nxi * nq * S_total
        print("%s m=%d S=%d in blocks of %d: %.3f s  %.3e evolved GPDs/s  %.2f TFLOP/s (%.2f of FP64)"
              % (name, m, S_total, chunk, t, S_total / t, flop / t / 1e12, flop / t / 1e12 / peak.value))
    del g_all
ms = (C.c_double * 3)()
_lib.check(L.tkb_stage_ms(ms))
flop_step = 144.0 * nx ** 3 * nxi * (nq - 1)
print("stepping: %.2f ms for %d intervals -> %.2f TFLOP/s (%.2f of FP64 peak), incl. composition" % (ms[2], nq - 1, flop_step / (ms[2] * 1e-3) / 1e12, flop_step / (ms[2] * 1e-3) / 1e12 / peak.value))
