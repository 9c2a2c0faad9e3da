#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: evolution-matrix sets per second
(xi points / s) at nx=160, NLO (BASELINE.json metric), one process per GPU.

A "step" is one full pass of the hot path over the workload: tkb_make_kernels (fixed quadrature K1,
QAGS constants K2) + tkb_make_matrices (composition K3, RK4 stepping K4) for all xi of the rank.
  value   device-timed (CUDA events on the launching stream, barrier + synchronize on both sides,
          max over ranks): xi points of all ranks / time.  The only inputs are the xi and Q2
          lists (a few KB of host parameters); everything else is produced in HBM.
  e2e     the same metric through the public Python API (tiktaalik_b200.matrices) with the xi / Q2
          host arrays going in and the four matrix sets read back into pinned host memory
          every step (what the reference's f2py getters deliver: host arrays).
  roofline  dominant kernel = K2 (QAGS), FP64-pipe bound: algorithmic flops (integrand evaluations
          of the executed QUADPACK path x reference-formula flops per evaluation)
          / K2's CUDA-event duration, against the DFMA peak measured in the same run.
  cpu_baseline  the oracle in its reference-shaped unfactored loop nest (OpenMP like matevo.f90:288)
          on a bounded sample of the same workload, all host cores.

`--impl reference` times the reference's CPU algorithm (oracle port; the Fortran cannot be built in
this image) on the host cores for the same metric and config.
N > 1 (torchrun): weak scaling, 256 xi per rank taken from one global log-spaced list, no data-path
collective (every xi is independent); the matrices stay xi-sharded.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NX = 160
NXI_PER_RANK = 256
GRID_TYPE = 2
Q2 = np.array([4.0, 10.0])
XI_MIN, XI_MAX = 2e-5, 0.5
WORKLOAD = "C3: NLO evolution-matrix sets, nx=160, grid_type=2, 256 xi per GPU log-spaced 2e-5..0.5, Q2 4->10 GeV^2"


def xi_for_rank(rank, world):
    xi = np.geomspace(XI_MIN, XI_MAX, NXI_PER_RANK * world)
    return np.ascontiguousarray(xi[rank::world])       # interleaved: every rank sees the same cost mix


def load_json(path, default):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0]))
                smax.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_step(xi_sample, threads):
    """One pass of the reference-shaped CPU algorithm (oracle mode 0) over xi_sample; returns seconds."""
    from oracle import tko
    os.environ["OMP_NUM_THREADS"] = str(threads)
    t0 = time.perf_counter()
    tko.initialize_kernels(NX, xi_sample, GRID_TYPE, mode=0)
    tko.initialize_evolution_matrices(Q2, nlo=True)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    xi_all = np.geomspace(XI_MIN, XI_MAX, NXI_PER_RANK * world)
    # four xi spanning the range per step (QAGS cost varies with xi): ~7 s of CPU per step on 16 threads
    sample = xi_all[[len(xi_all) // 8, (3 * len(xi_all)) // 8, (5 * len(xi_all)) // 8, (7 * len(xi_all)) // 8]]
    for _ in range(args.warmup):
        cpu_reference_step(sample, threads)
    times = [cpu_reference_step(sample, threads) for _ in range(args.steps)]
    t = float(np.sum(times))
    value = args.steps * sample.shape[0] / t
    sample_txt = "4 of %d xi (%s) per step, full nx=160 NLO assembly + RK4" % (
        len(xi_all), ", ".join("%.2e" % v for v in sample))
    line = {
        "impl": "reference", "metric": "evolution-matrix sets/sec (xi pts/s) at nx=160 NLO", "value": value,
        "unit": "xi pts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nx": NX, "grid_type": GRID_TYPE, "nlo": True, "Q2": Q2.tolist(),
                   "note": "reference CPU algorithm (oracle port of the Fortran, unfactored loop nest, OpenMP)"},
        "cpu_baseline": {"value": value, "unit": "xi pts/s", "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "xi pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import tiktaalik_b200 as tk
    from tiktaalik_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    xi = xi_for_rank(rank, world)
    xi_p = xi.ctypes.data_as(_lib._dp)
    q2_p = Q2.ctypes.data_as(_lib._dp)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.zero_()
        _lib.check(L.tkb_make_kernels_shard(NX, xi.shape[0], xi_p, GRID_TYPE, 0, 0, _lib.stream_ptr()))
        _lib.check(L.tkb_make_matrices(Q2.shape[0], q2_p, 1, _lib.stream_ptr()))

    # FP64 roofline denominator, measured here (MEASURED_PEAKS.json has no FP64 entry)
    import ctypes as C
    peak = C.c_double()
    _lib.check(L.tkb_fp64_peak(C.byref(peak)))

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.tkb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = np.zeros(3)
    ms3 = (C.c_double * 3)()
    ev0.record()
    for _ in range(args.steps):
        step()
        _lib.check(L.tkb_stage_ms(ms3))
        stage += np.array(list(ms3))
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = L.tkb_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    stage /= args.steps
    total_xi = NXI_PER_RANK * world
    value = total_xi * args.steps / (ms * 1e-3)

    # algorithmic work of K2 on this rank: integrand evaluations of the executed QUADPACK path
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
    last = _lib.view_tensor(b.value, (xi.shape[0], NX, 7, 3), torch.int32).to(torch.int64)
    neval_c = (42 * last - 21).sum(dim=(0, 1, 3)).cpu().numpy().astype(np.float64)        # per constant
    fpe = load_json(os.path.join(ROOT, "profiles", "k2_flops_per_eval.json"), {})
    flops_per_eval = np.array(fpe.get("flops_per_eval", [4500.0, 4500.0, 250.0, 5200.0, 4200.0, 600.0, 500.0]))
    k2_flops = float((neval_c * flops_per_eval).sum())
    k2_tflops = k2_flops / (stage[1] * 1e-3) / 1e12 if stage[1] > 0 else 0.0

    # ---- e2e through the public API with host buffers ----
    f90 = tk.f90wrap.matrices.dummy
    f90.zero_copy = True
    names = [("matrix_VNS", {"ns_type": 1}), ("matrix_VNS", {"ns_type": -1}), ("matrix_VSG", {}), ("matrix_ASG", {})]
    host = None
    d2h_bytes = 0

    # The device->host read of step k (1.05 GB over PCIe, ~18 ms) runs on a copy stream under the kernel assembly of
    # step k+1, which does not touch the matrices; the stepping of step k+1 waits for it.  Every step's matrices
    # reach the pinned host buffers inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    copy_done = None

    def e2e_step():
        nonlocal host, d2h_bytes, copy_done
        tk.shard.f90src.make_kernels_shard(NX, xi, GRID_TYPE, 0, 0)
        if copy_done is not None:
            torch.cuda.current_stream().wait_event(copy_done)      # the stepping overwrites what is being read
        tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
        mats = [getattr(tk.matrices, n)(**kw).permute(2, 3, 0, 1) for n, kw in names]     # storage order, contiguous
        if host is None:
            host = [torch.empty(tuple(m.shape), dtype=torch.float64).pin_memory() for m in mats]
            d2h_bytes = sum(m.numel() * 8 for m in mats)
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            for h, m in zip(host, mats):
                h.copy_(m, non_blocking=True)
            copy_done = torch.cuda.Event()
            copy_done.record(copy_stream)

    e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()                                       # the last step's read included
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = total_xi * args.steps / float(t_e.item())
    f90.zero_copy = False

    # ---- second half of the BASELINE metric: evolved GPDs/s (config C4: stored NLO matrices nx=160, 128 xi,
    # 16 Q2 applied to S synthetic samples), rank-local, device-timed; HBM-bound at S=1, FP64-bound at S=1024 ----
    application = None
    try:
        hbm = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}).get("hbm_gbs")
        hbm_src = "measured (MEASURED_PEAKS.json)" if hbm else "fallback (B200_PROFILING.md)"
        hbm = hbm or 6650.0
        xi4 = np.geomspace(1e-3, 0.9, 128)
        tk.matrices.initialize_kernels(NX, xi4, grid_type=2)
        tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 10.0, 16), nlo=True)
        f90.zero_copy = True
        Msg = tk.matrices.matrix_VSG()
        f90.zero_copy = False
        gen = torch.Generator(device=dev).manual_seed(1234)
        application = {"workload": "C4: V-type singlet NLO matrices nx=160 (m=320), 128 xi, 16 Q2, S samples"}
        for S in (1, 1024):
            g = torch.rand((2 * NX, 128, S), generator=gen, dtype=torch.float64, device=dev)
            tk.matrices.apply_matrix(Msg, g)
            ts = []
            for _ in range(5):
                flush.zero_()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                tk.matrices.apply_matrix(Msg, g)
                a1.record()
                torch.cuda.synchronize()
                ts.append(a0.elapsed_time(a1))
            t = float(np.median(ts)) * 1e-3
            m = 2 * NX
            byt = 8.0 * (m * m * 128 * 16 + m * 128 * S + m * 128 * S * 16)
            flop = 2.0 * m * m * 128 * 16 * S
            application["S=%d" % S] = {"ms": t * 1e3, "evolved_gpds_per_s": S / t, "GBps": byt / t / 1e9,
                                       "frac_of_hbm": byt / t / 1e9 / hbm, "TFLOPs": flop / t / 1e12,
                                       "frac_of_fp64": flop / t / 1e12 / peak.value}
        # the configuration at its full sample count: 10^5 samples read in place in blocks of 1024, results rotating
        # through two buffers (the 1 TB result never exists as a whole); one timed pass over all samples
        S_total = 100000
        g_all = torch.rand((2 * NX, 128, S_total), generator=gen, dtype=torch.float64, device=dev)
        for _ in tk.matrices.apply_matrix_chunks(Msg, g_all[:, :, :2048], chunk=1024):
            pass
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        nblk = 0
        for _ in tk.matrices.apply_matrix_chunks(Msg, g_all, chunk=1024):
            nblk += 1
        a1.record()
        torch.cuda.synchronize()
        t = a0.elapsed_time(a1) * 1e-3
        flop = 2.0 * m * m * 128 * 16 * S_total
        application["S=%d" % S_total] = {"ms": t * 1e3, "blocks": nblk, "block_samples": 1024, "evolved_gpds_per_s": S_total / t,
                                         "TFLOPs": flop / t / 1e12, "frac_of_fp64": flop / t / 1e12 / peak.value,
                                         "input_bytes": int(g_all.numel() * 8)}
        del g_all
        application["hbm_peak_GBps"] = hbm
        application["hbm_peak_source"] = hbm_src
    except Exception as exc:      # the headline line must still be printed
        application = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "evolution-matrix sets/sec (xi pts/s) at nx=160 NLO", "value": value, "unit": "xi pts/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "nx": NX, "grid_type": GRID_TYPE, "nlo": True, "Q2": Q2.tolist(),
                       "xi_per_gpu": NXI_PER_RANK, "l2": "256 MB buffer rewritten between steps; outputs (2.2 GB) exceed L2",
                       "stage_ms": {"K1_fixed_quadrature_span_overlapped_with_K2": stage[0],
                                    "assembly_K1_and_K2": stage[1], "K3K4_compose_rk4": stage[2]}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "xi pts/s", "h2d_bytes_per_step": int(xi.nbytes + Q2.nbytes),
                    "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": int(launches),
            "application": application,
            "roofline": {"bound": "fp64", "kernel": "K2, the QAGS constants: assemble_cst_pipe_kernel<KID> x 7 launches; duration = the assembly span, which also contains the overlapped fixed-quadrature kernel K1",
                         "achieved": k2_tflops, "peak": peak.value, "unit": "TFLOP/s",
                         "frac": k2_tflops / peak.value if peak.value else None, "traffic": fpe.get("traffic_bytes"),
                         "traffic_source": fpe.get("traffic_source"),
                         "peak_source": "DFMA microbenchmark measured in this run (tkb_fp64_peak); MEASURED_PEAKS.json "
                                        "has HBM and bf16 only",
                         "share_of_step": stage[1] / (ms / args.steps),
                         "k2_integrand_evaluations_per_s": float(neval_c.sum()) / (stage[1] * 1e-3) if stage[1] > 0 else None,
                         "k1_kernel_evaluations_per_step": 2898.0 * NX * xi.shape[0],
                         "flops_per_eval_source": fpe.get("source", "estimate (no ncu capture yet)")},
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            xi_all = np.geomspace(XI_MIN, XI_MAX, NXI_PER_RANK * world)
            sample = xi_all[np.linspace(len(xi_all) // 16, len(xi_all) - 1 - len(xi_all) // 16, 8).astype(int)]
            t = cpu_reference_step(sample, threads)
            line["cpu_baseline"] = {"value": sample.shape[0] / t, "unit": "xi pts/s", "cores": threads, "kind": "port",
                                    "sample": "8 of %d xi (strided over the range, %.2e .. %.2e), full nx=160 NLO "
                                              "assembly + RK4, reference-shaped unfactored loop nest, %.1f s"
                                              % (len(xi_all), sample[0], sample[-1], t)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
