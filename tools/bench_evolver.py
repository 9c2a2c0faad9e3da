"""Timing of the Evolver application pipeline: evolveGPDs on a GK-model-shaped batch.
Prints the time per call, the share spent inside the application kernels (tkb_apply) and evolved GPDs/s."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiktaalik_b200 as tk

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 160
nxi = int(sys.argv[2]) if len(sys.argv) > 2 else 32
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 16
nQ2 = int(sys.argv[4]) if len(sys.argv) > 4 else 16
xi = np.geomspace(1e-3, 0.9, nxi)
t0 = time.perf_counter()
ev = tk.evolution.Evolver(nx=nx, xi=xi, nQ2=nQ2, Q2i=1.0, Q2f=100.0, grid_type=2)      # crosses mc2 and mb2
torch.cuda.synchronize()
print("Evolver construction (kernels + 24 matrix sets over 3 segments): %.2f s" % (time.perf_counter() - t0))
gen = torch.Generator(device="cuda").manual_seed(3)
gpd = torch.rand((4, nx, nxi, nt), generator=gen, dtype=torch.float64, device="cuda")     # g, u, d, s
for nlo in (False, True):
    out = ev.evolveGPDs(gpd, gpd_type='V', nlo=nlo)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = ev.evolveGPDs(gpd, gpd_type='V', nlo=nlo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        ev.evolveGPDs(gpd, gpd_type='V', nlo=nlo)
        torch.cuda.synchronize()
    ka = prof.key_averages()
    tot = sum(k.device_time_total for k in ka) / 1e3
    ours = sum(k.device_time_total for k in ka if "tkb::" in k.key) / 1e3
    print("evolveGPDs nlo=%s: out %s, %.3f ms per call (device kernels %.3f ms, of which application kernels %.3f ms); "
          "%.3g evolved flavour-GPD samples/s" % (nlo, tuple(out.shape), ms, tot, ours, 4 * nt * nxi / (ms * 1e-3)))
