"""BASELINE config 5 / the north star's target: NLO evolution-matrix sets for a sweep of xi values at nx = 320,
grid_type = 2, sharded by xi over the GPUs of one node, then the final all-gather of the four matrix families
over NVLink.  Run:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29519 \
      tools/c5_sweep.py [nxi_total=4096] [nx=320]
Device-timed (CUDA events), max over ranks.  Checks: the gathered array holds this rank's block bit for bit and
every foreign block has the checksum its owner computed."""
import json
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
nxi = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 320
xi = np.geomspace(2e-5, 0.9, nxi)
Q2 = np.array([4.0, 10.0])
counts, offsets = tk.shard.xi_partition(nxi, world)


def timed(fn):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, t.item()


def build():
    tk.shard.initialize_kernels_sharded(nx, xi, grid_type=2, rank=rank, world_size=world)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)


# warm-up: a small build, and one gather of it so that NCCL sets up its connections outside the timed region
wx = xi[:: max(1, nxi // (2 * world))][: 2 * world]
tk.shard.initialize_kernels_sharded(24, wx, grid_type=2, rank=rank, world_size=world)
tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
tk.shard.init_capi()
tk.shard.gather_matrix_capi("matrix_VSG", tk.shard.xi_partition(wx.shape[0], world)[0])
torch.cuda.synchronize()
_, t_build = timed(build)
tk.f90wrap.matrices.dummy.zero_copy = True
names = ("matrix_VNS", "matrix_ANS", "matrix_VSG", "matrix_ASG")
ok = True
t_gather, gathered_bytes = 0.0, 0
for n in names:
    loc = getattr(tk.matrices, n)()
    m_ = loc.shape[0]
    buf = torch.empty((nxi, Q2.shape[0], m_, m_), dtype=torch.float64, device="cuda")      # allocation is not timed
    got, t = timed(lambda: tk.shard.gather_matrix_capi(n, counts, out=buf))
    t_gather += t
    gathered_bytes += got.numel() * 8
    mine = got[:, :, offsets[rank]:offsets[rank] + counts[rank], :]
    ok = ok and torch.equal(mine, loc)
    sums = torch.stack([got[:, :, offsets[r]:offsets[r] + counts[r], :].nan_to_num().sum() for r in range(world)])
    owner = sums.clone()
    for r in range(world):
        dist.broadcast(owner[r:r + 1], src=r)
    ok = ok and torch.equal(sums, owner)
    del got, buf
tk._lib.lib().tkb_shard_finalize()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"workload": "C5: NLO matrix sets, nx=%d, grid_type=2, %d xi geomspace(2e-5, 0.9), Q2 4->10, %d GPUs" % (nx, nxi, world),
                      "xi_per_gpu": counts[0], "build_ms": t_build, "xi_pts_per_s_build": nxi / (t_build * 1e-3),
                      "allgather_ms": t_gather, "allgather_GB_per_rank": gathered_bytes / 1e9,
                      "allgather_GBps_per_rank": gathered_bytes * (world - 1) / world / (t_gather * 1e-3) / 1e9 if world > 1 else None,
                      "xi_pts_per_s_with_gather": nxi / ((t_build + t_gather) * 1e-3),
                      "gather_check": "OK" if flag.item() == 1 else "MISMATCH"}))
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
