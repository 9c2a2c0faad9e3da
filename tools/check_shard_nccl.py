"""torchrun check (N GPUs of one node): xi-sharded assembly + stepping + NCCL all-gather equals the
unsharded result computed on every rank.  Run:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_shard_nccl.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import tiktaalik_b200 as tk

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx, gt = 24, 2
xi = np.geomspace(1e-7, 0.9999999, 2 * world + 1)          # uneven split, both ends clamped
Q2 = tk.matrices.Q2space(4.0, 30.0, 4)
# reference: the whole list on this GPU
tk.matrices.initialize_kernels(nx, xi, grid_type=gt)
tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
full = {n: getattr(tk.matrices, n)().clone() for n in ("matrix_VNS", "matrix_ANS", "matrix_VSG", "matrix_ASG")}
# sharded
mine = tk.shard.initialize_kernels_sharded(nx, xi, grid_type=gt, rank=rank, world_size=world)
tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
ok = True
for n, ref in full.items():
    got = tk.shard.gather_matrix(getattr(tk.matrices, n)())
    same = tuple(got.shape) == tuple(ref.shape) and torch.equal(got, ref)
    ok = ok and same
# the same gather through the C ABI (tkb_shard_init / tkb_allgather_matrices[_strided]), both layouts; the strided
# layout also on a list with equal shares (the per-xi NCCL group) gathered in two chunks of local xi
tk.shard.init_capi()
for layout in ("strided", "blocks"):
    counts, _ = tk.shard.xi_partition(xi.shape[0], world, layout)
    tk.shard.initialize_kernels_sharded(nx, xi, grid_type=gt, rank=rank, world_size=world, layout=layout)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    for n, ref in full.items():
        got = tk.shard.gather_matrix_capi(n, counts, layout=layout)
        torch.cuda.synchronize()
        same = tuple(got.shape) == tuple(ref.shape) and torch.equal(got, ref)
        if not same:
            print("rank %d: C-ABI gather (%s) of %s differs" % (rank, layout, n))
        ok = ok and same
xi_even = tk.shard.clamp_xi(xi[:2 * world])
tk.matrices.initialize_kernels(nx, xi_even, grid_type=gt)
tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
full_even = {n: getattr(tk.matrices, n)().clone() for n in full}
outs = {n: torch.empty((2 * world, Q2.shape[0], r.shape[0], r.shape[0]), dtype=torch.float64, device="cuda") for n, r in full_even.items()}
for chunk in range(2):                 # one local xi per chunk: k_first = chunk
    mine_c = xi_even[rank::world][chunk:chunk + 1]
    tk.shard.f90src.make_kernels_shard(nx, mine_c, gt, 0, 0)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    for n in full_even:
        tk.shard.gather_matrix_capi(n, [2] * world, out=outs[n], layout="strided", k_first=chunk)
torch.cuda.synchronize()
for n, ref in full_even.items():
    same = torch.equal(outs[n].permute(2, 3, 0, 1), ref)
    if not same:
        print("rank %d: chunked strided gather of %s differs" % (rank, n))
    ok = ok and same
# the exchange by the copy engines (tkb_p2p_*): two chunks of one local xi each pushed into every rank's buffers
views = tk.shard.init_p2p(nx, 2 * world, Q2.shape[0])
for chunk in range(2):
    mine_c = xi_even[rank::world][chunk:chunk + 1]
    tk.shard.f90src.make_kernels_shard(nx, mine_c, gt, 0, 0)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    tk.shard.push_matrices_p2p(k_first=chunk)
torch.cuda.synchronize()
dist.barrier()
for slot, n in enumerate(("matrix_VNS", "matrix_ANS", "matrix_VSG", "matrix_ASG")):     # ANS(+1) is VNS(-1)
    same = torch.equal(views[slot], full_even[n])
    if not same:
        print("rank %d: copy-engine exchange of %s differs" % (rank, n))
    ok = ok and same
dist.barrier()
tk.shard.close_p2p()
# bandwidth of the one exchange step at the bench's size (nx = 160, 256 xi per rank, Q2 4 -> 10): each rank
# receives (world - 1) slabs of every matrix family over NVLink; device-timed, max over ranks
if len(sys.argv) > 1 and sys.argv[1] == "--bw":
    nxb, per = 160, 256
    xib = np.geomspace(2e-5, 0.5, per * world)
    tk.shard.initialize_kernels_sharded(nxb, xib, grid_type=2, rank=rank, world_size=world)
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 10.0]), nlo=True)
    cb, _ = tk.shard.xi_partition(xib.shape[0], world)
    for n in ("matrix_VNS", "matrix_VSG"):
        m = nxb * (2 if n.endswith("SG") else 1)
        recv = 8.0 * m * m * 2 * per * (world - 1)
        best = None
        for rep in range(6):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            got = tk.shard.gather_matrix_capi(n, cb)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep and (best is None or t.item() < best):
                best = t.item()
            del got
        if rank == 0:
            print("all-gather of %s over %d GPUs: %.1f MB received per rank in %.3f ms = %.1f GB/s per rank (incl. allocation of the result)"
                  % (n, world, recv / 1e6, best, recv / best / 1e6))
tk._lib.lib().tkb_shard_finalize()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("shard+gather check over %d GPUs (%d xi, local %d): %s" % (world, xi.shape[0], mine.shape[0], "OK" if flag.item() == 1 else "MISMATCH"))
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
