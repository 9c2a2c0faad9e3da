#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: evolution-matrix sets per second
(xi points / s) at nx=160, NLO (BASELINE.json metric), one process per GPU.

A "step" is one full pass of the hot path over the workload: tkb_make_kernels (fixed quadrature K1,
QAGS constants K2) + tkb_make_matrices (composition K3, RK4 stepping K4) for all xi of the rank and,
with more than one GPU, the all-gather of the four matrix families along xi (the one collective of
the path).
  value   device-timed (CUDA events on the launching stream, barrier + synchronize on both sides,
          max over ranks): xi points of all ranks / time.  The only inputs are the xi and Q2
          lists (a few KB of host parameters); everything else is produced in HBM.
  N > 1   weak scaling: 256 xi per rank (c3) taken by stride from one global log-spaced list (rank r holds
          xi[r::N]: every rank the same cost mix; tiktaalik_b200.shard uses the same partition).  The gather of
          step k runs on a side stream under the kernel assembly of step k+1, which does not touch the matrices;
          the stepping of step k+1 waits for it, the last gather is waited for inside the timed region.  Default
          form: copy-engine pushes into every rank's receive buffer (tkb_p2p_push_matrices, CUDA IPC peer memory
          over NVLink: no SMs, so it really runs under the QAGS kernels, which own every SM); --gather nccl: per-xi
          NCCL all-gathers (tkb_allgather_matrices_strided).  Either way the result is in the global list's order.
          `allgather` reports its own duration and the bytes a rank receives.
  e2e     the same metric through the public Python API (tiktaalik_b200.matrices) with the xi / Q2
          host arrays going in and the four matrix sets read back into pinned host memory
          every step (what the reference's f2py getters deliver: host arrays); pipelined over steps
          (the read of step k under the assembly of step k+1, zero-copy views).  `e2e.single_call_ms` is the
          latency of ONE such call through the default (cloning) getters with nothing overlapped.
  roofline  dominant kernel = K2 (QAGS), FP64-pipe bound: algorithmic flops (integrand evaluations
          of the executed QUADPACK path x the per-constant flops of one straight evaluation of the
          reference formulas, profiles/k2_flops_per_eval.json) / K2's CUDA-event duration, against the
          larger of the two FP64 peaks measured in this run (DFMA microbenchmark, cuBLAS DGEMM 8192^3).
  cpu_baseline  the oracle in its reference-shaped unfactored loop nest (OpenMP like matevo.f90:288)
          on a bounded sample of the same workload (8 xi strided over the whole list, both ends included),
          all host cores.  `--impl reference` times the same thing with the same sampling rule.

`--workload c5`: the north star's target shape (nx=320, grid 2, 512 xi per rank of geomspace(2e-5, 0.9), Q2 4->10,
NLO), the rank's xi processed in 4 chunks of 128, each chunk's gather under the next chunk's assembly.
"""
import argparse
import ctypes as C
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

GRID_TYPE = 2
Q2 = np.array([4.0, 10.0])
WORKLOADS = {
    "c3": {"nx": 160, "xi_per_rank": 256, "xi_min": 2e-5, "xi_max": 0.5, "chunks": 1,
           "name": "C3: NLO evolution-matrix sets, nx=160, grid_type=2, 256 xi per GPU log-spaced 2e-5..0.5, Q2 4->10 GeV^2"},
    "c5": {"nx": 320, "xi_per_rank": 512, "xi_min": 2e-5, "xi_max": 0.9, "chunks": 4,
           "name": "C5: NLO evolution-matrix sets, nx=320, grid_type=2, 512 xi per GPU log-spaced 2e-5..0.9 (4096 xi on 8 GPUs), "
                   "Q2 4->10 GeV^2, 4 chunks of 128 xi per rank"},
}
FAMILIES = [(0, 1), (0, -1), (1, 1), (3, 1)]            # (type, nstype): VNS+, VNS-, VSG, ASG   (TKB_VNS = 0, VSG = 1, ASG = 3)
CPU_KIND_NOTE = ("restated C port of the Fortran built like the reference itself (oracle/libtko_fast.so: -Ofast -march=x86-64-v3 "
                 "-fopenmp, libm log/pow; f90src/CMakeLists.txt:12-14), OpenMP over x rows; timing only -- the bit-exact checker "
                 "build (own log/pow, -O3 -ffp-contract=off) is about 1.9x slower")


def global_xi(wl, world):
    return np.geomspace(wl["xi_min"], wl["xi_max"], wl["xi_per_rank"] * world)


def xi_for_rank(wl, rank, world):
    return np.ascontiguousarray(global_xi(wl, world)[rank::world])       # strided: every rank sees the same cost mix


def cpu_sample(xi_all):
    """the one sampling rule of both CPU legs: 8 xi strided over the whole list, first (smallest, the most expensive:
    QAGS reaches limit=500 there) and last included"""
    return xi_all[np.linspace(0, len(xi_all) - 1, 8).round().astype(int)]


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


def cpu_reference_step(nx, xi_sample, threads, fast=True):
    """One pass of the reference-shaped CPU algorithm (oracle mode 0) over xi_sample; returns seconds.
    fast: the timing-only build with the reference's own compiler flags and libm (oracle/tko_fp.h, TKO_LIBM)."""
    from oracle import tko
    tko.use_fast_build(fast)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    try:        # torchrun exports OMP_NUM_THREADS=1 and libgomp has read it by now: set the thread count directly
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    except OSError:
        pass
    t0 = time.perf_counter()
    tko.initialize_kernels(nx, xi_sample, GRID_TYPE, mode=0)
    tko.initialize_evolution_matrices(Q2, nlo=True)
    return time.perf_counter() - t0


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    xi_all = global_xi(wl, world)
    sample = cpu_sample(xi_all)
    for _ in range(args.warmup):
        cpu_reference_step(wl["nx"], sample, threads)
    times = [cpu_reference_step(wl["nx"], sample, threads) for _ in range(args.steps)]
    t = float(np.sum(times))
    value = args.steps * sample.shape[0] / t
    sample_txt = "8 of %d xi per step (strided over the whole list, %.2e .. %.2e, both ends included), full nx=%d NLO assembly + RK4" % (
        len(xi_all), sample[0], sample[-1], wl["nx"])
    line = {
        "impl": "reference", "metric": "evolution-matrix sets/sec (xi pts/s) at nx=%d NLO" % wl["nx"], "value": value,
        "unit": "xi pts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "nx": wl["nx"], "grid_type": GRID_TYPE, "nlo": True, "Q2": Q2.tolist(),
                   "note": "reference CPU algorithm (oracle port of the Fortran, unfactored loop nest, OpenMP); the Fortran "
                           "itself cannot be built: no Fortran compiler in this image or on the GPU box"},
        "cpu_baseline": {"value": value, "unit": "xi pts/s", "cores": threads, "kind": "port", "kind_note": CPU_KIND_NOTE,
                         "sample": sample_txt},
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
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: the exchange by copy-engine pushes into peer memory (default) or by NCCL all-gathers")
    ap.add_argument("--verify", action="store_true",
                    help="N > 1: after the timed region check the gathered arrays (own slabs bit for bit, foreign slabs by the "
                         "owners' checksums) for the last chunk of the last step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the e2e, application and CPU legs (profiling runs)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    NX = wl["nx"]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
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
    xi = xi_for_rank(wl, rank, world)
    nchunk = wl["chunks"]
    chunk_xi = [np.ascontiguousarray(c) for c in np.array_split(xi, nchunk)]
    assert len({c.shape[0] for c in chunk_xi}) == 1
    per_chunk = chunk_xi[0].shape[0]
    q2_p = Q2.ctypes.data_as(_lib._dp)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)      # > 126 MB L2
    main_stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the collective (N > 1): gathered arrays [xi_total][Q2][m][m] in the order of the global list ----
    gather_stream = gathered = None
    gather_done = None
    gather_ev = []
    use_p2p = args.gather == "p2p"
    if world > 1:
        gather_stream = torch.cuda.Stream(device=dev)
        if use_p2p:
            gathered = tk.shard.init_p2p(NX, xi.shape[0] * world, Q2.shape[0], strict=False)
            if gathered is None:                 # CUDA IPC between the ranks is not available on this box: NCCL instead
                use_p2p = False
        if not use_p2p:
            tk.shard.init_capi()
            gathered = [torch.empty((xi.shape[0] * world, Q2.shape[0], NX * (1 if t == 0 else 2), NX * (1 if t == 0 else 2)),
                                    dtype=torch.float64, device=dev) for t, _ in FAMILIES]
    recv_bytes = sum(g.numel() * 8 for g in gathered) * (world - 1) // world if world > 1 else 0

    def enqueue_exchange(k_first):
        if use_p2p:
            tk.shard.push_matrices_p2p(k_first, stream=gather_stream)
        else:
            for (t, ns), out in zip(FAMILIES, gathered):
                _lib.check(L.tkb_allgather_matrices_strided(t, ns, k_first, C.c_void_p(out.data_ptr()),
                                                            C.c_void_p(gather_stream.cuda_stream)))

    def launch_gather(k_first, timed):
        """all four families of the rank's current matrices -> their places in the global order, on the side stream"""
        nonlocal gather_done
        gather_stream.wait_stream(main_stream)
        with torch.cuda.stream(gather_stream):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(gather_stream)
            enqueue_exchange(k_first)
            e1.record(gather_stream)
        gather_done = e1
        if timed:
            gather_ev.append((e0, e1))

    stage_acc = np.zeros(3)
    ms3 = (C.c_double * 3)()

    def step(timed=False):
        flush.zero_()
        for c, cx in enumerate(chunk_xi):
            _lib.check(L.tkb_make_kernels_shard(NX, cx.shape[0], cx.ctypes.data_as(_lib._dp), GRID_TYPE, 0, 0, _lib.stream_ptr()))
            if gather_done is not None:
                main_stream.wait_event(gather_done)          # the stepping rewrites the matrices the gather reads
            _lib.check(L.tkb_make_matrices(Q2.shape[0], q2_p, 1, _lib.stream_ptr()))
            if timed:
                _lib.check(L.tkb_stage_ms(ms3))
                stage_acc[:] += np.array(list(ms3))
            if world > 1:
                launch_gather(c * per_chunk, timed)

    # FP64 roofline denominators, measured here (MEASURED_PEAKS.json has no FP64 entry): a DFMA microbenchmark and
    # cuBLAS DGEMM 8192^3 (the FP64 tensor-core path); fractions are quoted against the larger one
    peak = C.c_double()
    _lib.check(L.tkb_fp64_peak(C.byref(peak)))
    a8 = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b8 = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(a8, b8)
    dg = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a8, b8)
        e1.record()
        torch.cuda.synchronize()
        dg.append(2.0 * 8192.0 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    dgemm_tflops = max(dg)
    del a8, b8
    fp64_peak = max(peak.value, dgemm_tflops)

    for _ in range(args.warmup):
        step()
    if gather_done is not None:
        main_stream.wait_event(gather_done)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.tkb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(timed=True)
    if gather_done is not None:
        main_stream.wait_event(gather_done)                  # the last gather belongs to the timed region
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = L.tkb_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    stage = stage_acc / args.steps                           # per step (summed over the chunks of a step)
    total_xi = wl["xi_per_rank"] * world
    value = total_xi * args.steps / (ms * 1e-3)

    allgather = None
    if world > 1:
        g_ms = float(np.mean([a.elapsed_time(b) for a, b in gather_ev])) * nchunk      # per step, overlapped with assembly
        # the same exchange with nothing else running: 3 repetitions, the best one
        alone = []
        for _ in range(3):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(gather_stream):
                e0.record(gather_stream)
                enqueue_exchange((nchunk - 1) * per_chunk)
                e1.record(gather_stream)
            torch.cuda.synchronize()
            alone.append(e0.elapsed_time(e1))
        t_al = torch.tensor([min(alone)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_al, op=dist.ReduceOp.MAX)
        chunk_recv = recv_bytes / nchunk
        allgather = {"collective": ("copy-engine pushes into every rank's receive buffer over NVLink (tkb_p2p_push_matrices: one strided "
                                    "cudaMemcpy2DAsync per peer and matrix family, CUDA IPC peer memory, no SMs)" if use_p2p else
                                    "ncclAllGather per xi, one NCCL group per matrix family (tkb_allgather_matrices_strided)")
                                   + "; 4 families, result in the order of the global xi list"
                                   + ("" if use_p2p or args.gather == "nccl" else " (fallback: CUDA IPC peer mapping failed on this box)"),
                     "ms_per_step_overlapped_with_assembly": g_ms, "bytes_received_per_rank_per_step": int(recv_bytes),
                     "alone_ms_per_chunk": float(t_al.item()), "alone_GBps_per_rank": chunk_recv / (float(t_al.item()) * 1e-3) / 1e9,
                     "chunks_per_step": nchunk, "inside_timed_region": True}

    verify = None
    if world > 1 and args.verify:
        # the last chunk's matrices are still in the library: every rank's slabs must sit at (k_first + k) * world + rank
        # of every rank's gathered array
        barrier()
        ok = True
        k_first = (nchunk - 1) * per_chunk
        f90 = tk.f90wrap.matrices.dummy
        f90.zero_copy = True
        local = [tk.matrices.matrix_VNS(1), tk.matrices.matrix_VNS(-1), tk.matrices.matrix_VSG(), tk.matrices.matrix_ASG()]
        f90.zero_copy = False
        for loc, out in zip(local, gathered):
            store = out.permute(2, 3, 0, 1) if use_p2p else out                  # [xi][Q2][m][m]
            blk = store[k_first * world:(k_first + per_chunk) * world].reshape(per_chunk, world, *store.shape[1:])
            ok = ok and torch.equal(blk[:, rank], loc.permute(2, 3, 0, 1))
            mine = loc.nan_to_num().sum().reshape(1)
            owners = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(owners, mine)
            for r in range(world):
                ok = ok and bool(torch.equal(blk[:, r].nan_to_num().sum().reshape(1), owners[r]))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verify = "OK" if int(flag.item()) == 1 else "MISMATCH"

    # algorithmic work of K2 on this rank: integrand evaluations of the executed QUADPACK path (of the last chunk,
    # scaled to the step: the chunks are strided samples of the same list)
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
    last = _lib.view_tensor(b.value, (per_chunk, NX, 7, 3), torch.int32).to(torch.int64)
    neval_c = (42 * last - 21).sum(dim=(0, 1, 3)).cpu().numpy().astype(np.float64) * nchunk        # per constant
    fpe = load_json(os.path.join(ROOT, "profiles", "k2_flops_per_eval.json"), {})
    flops_per_eval = np.array(fpe.get("flops_per_eval", [3450.0] * 7))
    k2_flops = float((neval_c * flops_per_eval).sum())
    k2_tflops = k2_flops / (stage[1] * 1e-3) / 1e12 if stage[1] > 0 else 0.0
    exe = fpe.get("executed_flops_per_eval")
    executed_over_algorithmic = (float((neval_c * np.array(exe)).sum()) / k2_flops) if exe and k2_flops > 0 else None

    e2e = None
    application = None
    if not args.no_extras and args.workload == "c3":
        # ---- e2e through the public API with host buffers ----
        f90 = tk.f90wrap.matrices.dummy
        names = [("matrix_VNS", {"ns_type": 1}), ("matrix_VNS", {"ns_type": -1}), ("matrix_VSG", {}), ("matrix_ASG", {})]
        # (1) one call, nothing overlapped, default getters (private copies): build, four getters, read into host buffers
        # (pinned and allocated beforehand, as a caller that does this repeatedly would; a first .cpu() into pageable
        # memory costs ~0.5 s more for the 1.05 GB)
        tk.matrices.initialize_kernels(NX, xi, grid_type=GRID_TYPE)
        tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
        host1 = [torch.empty(tuple(getattr(tk.matrices, n)(**kw).permute(2, 3, 0, 1).shape), dtype=torch.float64).pin_memory()
                 for n, kw in names]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tk.shard.f90src.make_kernels_shard(NX, xi, GRID_TYPE, 0, 0)
        tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
        for h, (n, kw) in zip(host1, names):
            h.copy_(getattr(tk.matrices, n)(**kw).permute(2, 3, 0, 1), non_blocking=True)
        torch.cuda.synchronize()
        single_call_ms = (time.perf_counter() - t0) * 1e3
        del host1
        # (2) throughput: the device->host read of step k (1.05 GB over PCIe, ~18 ms) runs on a copy stream under the
        # kernel assembly of step k+1, which does not touch the matrices; the stepping of step k+1 waits for it.
        # Every step's matrices reach the pinned host buffers inside the timed region.
        f90.zero_copy = True
        host = None
        d2h_bytes = 0
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
        f90.zero_copy = False
        e2e = {"value": total_xi * args.steps / float(t_e.item()), "unit": "xi pts/s",
               "h2d_bytes_per_step": int(xi.nbytes + Q2.nbytes), "d2h_bytes_per_step": int(d2h_bytes),
               "mode": "pipelined over steps: the read of step k on a copy stream under the assembly of step k+1 (zero-copy views); "
                       "every rank delivers its own xi share to its host",
               "single_call_ms": single_call_ms,
               "single_call_note": "latency of ONE call: build + four default (cloning) getters + copies into preallocated pinned host buffers, nothing overlapped"}

        # ---- second half of the BASELINE metric: evolved GPDs/s (config C4: stored NLO matrices nx=160, 128 xi,
        # 16 Q2 applied to S synthetic samples), rank-local, device-timed; HBM-bound at S=1, FP64-bound at S=1024 ----
        try:
            hbm = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}).get("hbm_gbs")
            hbm_src = "measured (MEASURED_PEAKS.json)" if hbm else "fallback (B200_PROFILING.md)"
            hbm = hbm or 6650.0
            xi4 = np.geomspace(1e-3, 0.9, 128)
            tk.matrices.initialize_kernels(NX, xi4, grid_type=2)
            tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 10.0, 16), nlo=True)      # allocations, composition
            tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 10.0, 16), nlo=True)      # the timed one
            stage4 = (C.c_double * 3)()
            _lib.check(L.tkb_stage_ms(stage4))
            f90.zero_copy = True
            Msg = tk.matrices.matrix_VSG()
            f90.zero_copy = False
            gen = torch.Generator(device=dev).manual_seed(1234)
            application = {"workload": "C4: V-type singlet NLO matrices nx=160 (m=320), 128 xi, 16 Q2, S samples",
                           "rk4_stepping_ms_15_intervals_128_xi": stage4[2],
                           "rk4_stepping_TFLOPs": 144.0 * NX ** 3 * 128 * 15 / (stage4[2] * 1e-3) / 1e12 if stage4[2] > 0 else None}
            if application["rk4_stepping_TFLOPs"]:
                application["rk4_stepping_frac_of_fp64"] = application["rk4_stepping_TFLOPs"] / fp64_peak
            m = 2 * NX
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
                byt = 8.0 * (m * m * 128 * 16 + m * 128 * S + m * 128 * S * 16)
                flop = 2.0 * m * m * 128 * 16 * S
                application["S=%d" % S] = {"ms": t * 1e3, "evolved_gpds_per_s": S / t, "GBps": byt / t / 1e9,
                                           "frac_of_hbm": byt / t / 1e9 / hbm, "TFLOPs": flop / t / 1e12,
                                           "frac_of_fp64": flop / t / 1e12 / fp64_peak}
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
                                             "TFLOPs": flop / t / 1e12, "frac_of_fp64": flop / t / 1e12 / fp64_peak,
                                             "input_bytes": int(g_all.numel() * 8)}
            del g_all
            application["hbm_peak_GBps"] = hbm
            application["hbm_peak_source"] = hbm_src
        except Exception as exc:      # the headline line must still be printed
            application = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "evolution-matrix sets/sec (xi pts/s) at nx=%d NLO" % NX, "value": value, "unit": "xi pts/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "nx": NX, "grid_type": GRID_TYPE, "nlo": True, "Q2": Q2.tolist(),
                       "xi_per_gpu": wl["xi_per_rank"], "xi_partition": "strided (rank r holds xi[r::n_gpus] of the global list)",
                       "l2": "256 MB buffer rewritten between steps; outputs (2.2 GB) exceed L2",
                       "stage_ms": {"K1_fixed_quadrature_span_overlapped_with_K2": stage[0],
                                    "assembly_K1_and_K2": stage[1], "K3K4_compose_rk4": stage[2]}},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "K2, the QAGS constants: six launches (assemble_cst_pair_kernel for KV1_NSp + KV1_NSm, assemble_cst_pipe_kernel<KID> for the other five); duration = the assembly span, which also contains the overlapped fixed-quadrature kernel K1",
                         "achieved": k2_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": k2_tflops / fp64_peak if fp64_peak else None, "traffic": fpe.get("traffic_bytes"),
                         "traffic_source": fpe.get("traffic_source"),
                         "peak_source": "measured in this run: max(DFMA microbenchmark tkb_fp64_peak, cuBLAS DGEMM 8192^3 via torch.matmul); "
                                        "MEASURED_PEAKS.json has HBM and bf16 only",
                         "peaks_TFLOPs": {"dfma_microbenchmark": peak.value, "cublas_dgemm_8192": dgemm_tflops},
                         "algorithmic_flops_per_eval": flops_per_eval.tolist(),
                         "integrand_evaluations_per_step": neval_c.tolist(),
                         "fp64_pipe_active_ncu": (fpe.get("fp64_pipe_active_ncu_time_weighted_pct") or 0.0) / 100.0 or None,
                         "fp64_pipe_active_ncu_per_constant_pct": fpe.get("fp64_pipe_active_ncu_pct"),
                         "executed_over_algorithmic": executed_over_algorithmic,
                         "ncu_source": fpe.get("executed_source"),
                         "share_of_step": stage[1] / (ms / args.steps),
                         "k2_integrand_evaluations_per_s": float(neval_c.sum()) / (stage[1] * 1e-3) if stage[1] > 0 else None,
                         "k1_kernel_evaluations_per_step": 2898.0 * NX * xi.shape[0],
                         "flops_per_eval_source": fpe.get("definition", "estimate (no ncu capture yet)")},
        }
        if e2e is not None:
            line["e2e"] = e2e
        else:
            line["e2e"] = {"value": None, "unit": "xi pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": "not measured in this mode"}
        if allgather is not None:
            if verify is not None:
                allgather["verified_last_chunk"] = verify
            line["allgather"] = allgather
        if application is not None:
            line["application"] = application
            for S in (1, 1024, 100000):
                k = application.get("S=%d" % S)
                if k:
                    line["evolved_gpds_per_s_S%d" % S] = k["evolved_gpds_per_s"]
                    if S == 1:
                        line["apply_frac_of_hbm_S1"] = k["frac_of_hbm"]
                    else:
                        line["apply_frac_of_fp64_S%d" % S] = k["frac_of_fp64"]
        if not args.no_cpu_baseline and not args.no_extras:
            threads = os.cpu_count() or 1
            xi_all = global_xi(wl, world)
            sample = cpu_sample(xi_all)
            t = cpu_reference_step(NX, sample, threads)
            t_exact = cpu_reference_step(NX, sample, threads, fast=False)
            line["cpu_baseline"] = {"value": sample.shape[0] / t, "unit": "xi pts/s", "cores": threads, "kind": "port",
                                    "kind_note": CPU_KIND_NOTE, "bit_exact_checker_build_value": sample.shape[0] / t_exact,
                                    "sample": "8 of %d xi (strided over the whole list, %.2e .. %.2e, both ends included), full nx=%d NLO "
                                              "assembly + RK4, reference-shaped unfactored loop nest, %.1f s"
                                              % (len(xi_all), sample[0], sample[-1], NX, t)}
        print(json.dumps(line))
    if world > 1:
        barrier()
        if use_p2p:
            tk.shard.close_p2p()
        else:
            tk._lib.lib().tkb_shard_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
