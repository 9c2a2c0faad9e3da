"""World-size-2 test of the xi-sharded gather on CPU (gloo): the only collective of the hot path is
the final all-gather of per-xi matrices."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, uneven, layout, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tiktaalik_b200 import shard
    nxi_total, m, nq = (5 if uneven else 4), 6, 3
    full = torch.arange(m * m * nxi_total * nq, dtype=torch.float64).reshape(nxi_total, nq, m, m).permute(2, 3, 0, 1)
    counts, offs = shard.xi_partition(nxi_total, world, layout)
    if layout == "strided":
        local = full[:, :, rank::world, :]
    else:
        local = full[:, :, offs[rank]:offs[rank] + counts[rank], :]
    got = shard.gather_matrix(local, counts=None if uneven else counts, layout=layout)
    ok = tuple(got.shape) == (m, m, nxi_total, nq) and torch.equal(got, full)       # the order of the global list
    xi = np.geomspace(1e-7, 0.9999999, nxi_total)
    mine, c2, o2 = shard.local_xi(xi, rank, world, layout)
    ok = ok and c2 == counts and mine.shape[0] == counts[rank] and mine.min() >= 1e-6 and mine.max() <= 0.999999
    ref = shard.clamp_xi(xi)
    ok = ok and np.array_equal(mine, ref[rank::world] if layout == "strided" else ref[offs[rank]:offs[rank] + counts[rank]])
    # fewer xi than ranks: every rank raises before any work (none is left waiting in a collective)
    try:
        shard.initialize_kernels_sharded(8, [0.5], rank=rank, world_size=world, layout=layout)
        ok = False
    except ValueError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _run(uneven, layout="strided"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, uneven, layout, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_gather_matrix_even_world2():
    _run(False)


def test_gather_matrix_uneven_world2():
    _run(True)


def test_gather_matrix_blocks_layout_world2():
    _run(False, "blocks")
    _run(True, "blocks")
