"""
shard.py -- multi-GPU partitioning of the hot path by xi (one process per GPU).

Every xi is independent in assembly and stepping (matevo.f90:291,325,359,381,426,452,514,550,584)
and in application (evolution.py:507-509), so each rank builds the matrices of its share of the xi
values with no data-path collective.  The only exchange is the optional final all-gather of the
per-xi matrices (NCCL over NVLink): the library stores M as [xi][Q2][i][j], so one xi is one
contiguous slab.

Two layouts of the share:
  "strided" (default)  rank r holds xi[r], xi[r + world], ...  The cost of a xi varies along the list (the QAGS
                       constants need 1.5x more integrand evaluations at xi = 2e-5 than at 0.5), so interleaving
                       gives every rank the same mix; bench.py uses the same partition.
  "blocks"             rank r holds a contiguous block (earlier ranks take the remainder).
Either way the gathered array is in the order of the global list.

The reference clamps only the global first and last xi (matevo.f90:688-689); the clamp is
therefore applied before sharding and no rank clamps again.
"""
import numpy as np
import torch

from .f90wrap.matrices import dummy as f90src


def xi_partition(nxi, world_size, layout="strided"):
    """(counts, offsets) per rank.  blocks: contiguous, earlier ranks take the remainder; strided: rank r holds
    the indices r, r + world, ... (offsets are then the first indices, i.e. the ranks themselves)."""
    nxi, world_size = int(nxi), int(world_size)
    if layout == "strided":
        counts = [len(range(r, nxi, world_size)) for r in range(world_size)]
        return counts, list(range(world_size))
    if layout != "blocks":
        raise ValueError("layout must be 'strided' or 'blocks'")
    base, rem = divmod(nxi, world_size)
    counts = [base + (1 if r < rem else 0) for r in range(world_size)]
    offsets = [int(v) for v in np.cumsum([0] + counts[:-1])]
    return counts, offsets


def clamp_xi(xi):
    """initialize_2D's clamp on the global list (matevo.f90:688-689)"""
    xi = np.array(np.atleast_1d(xi), dtype=np.float64)
    if xi[0] < 1e-6:
        xi[0] = 1e-6
    if xi[-1] > 0.999999:
        xi[-1] = 0.999999
    return xi


def local_xi(xi, rank, world_size, layout="strided"):
    xi = clamp_xi(xi)
    counts, offsets = xi_partition(xi.shape[0], world_size, layout)
    if layout == "strided":
        return np.ascontiguousarray(xi[rank::world_size]), counts, offsets
    return xi[offsets[rank]:offsets[rank] + counts[rank]], counts, offsets


def initialize_kernels_sharded(nx, xi, grid_type=1, rank=0, world_size=1, layout="strided"):
    """matrices.initialize_kernels for this rank's share of the global xi list."""
    assert(nx==(nx//2)*2)
    assert(nx>=6)
    nxi_total = np.atleast_1d(np.asarray(xi)).shape[0]
    if nxi_total < world_size:
        # the same test on every rank, before any work: all ranks raise, none is left waiting in the gather
        raise ValueError("%d xi values cannot be sharded over %d ranks (nxi < world_size)" % (nxi_total, world_size))
    mine, counts, offsets = local_xi(xi, rank, world_size, layout)
    # the list is already clamped, so no further clamp is applied on any rank
    f90src.make_kernels_shard(nx, mine, grid_type, 0, 0)
    return mine


def _global_order(allp, counts, layout):
    """allp: (world, cmax, ...) padded per-rank slabs -> (sum counts, ...) in the order of the global list"""
    world, cmax = allp.shape[0], allp.shape[1]
    if layout == "strided":
        # global index k * world + r <- rank r, local k.  The ranks with the short share are the last ones
        # (counts = [c + 1] * rem + [c] * (world - rem)), so the padded rows end up behind the global list
        out = allp.transpose(0, 1).reshape((world * cmax,) + tuple(allp.shape[2:]))
        return out[:sum(counts)] if len(set(counts)) > 1 else out
    return torch.cat([allp[r, :c] for r, c in enumerate(counts)], dim=0)


def gather_matrix(M_local, counts=None, group=None, layout="strided"):
    """All-gather per-xi matrices along xi.  M_local: (m, m, nxi_local, nQ2) as returned by the
    matrix getters.  Returns (m, m, nxi_total, nQ2) on every rank, xi in the order of the global list.
    World size 1: returns M_local."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return M_local
    world = dist.get_world_size(group)
    m, _, nloc, nq = M_local.shape
    store = M_local.permute(2, 3, 0, 1).contiguous()          # [xi][Q2][i][j], one slab per xi
    if counts is None:
        cnt = torch.tensor([nloc], dtype=torch.int64, device=store.device)
        allc = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(allc, cnt, group=group)
        counts = [int(c.item()) for c in allc]
    cmax = max(counts)
    if nloc < cmax:              # collectives need equal contributions: pad to the largest share, trim afterwards
        padded = torch.zeros((cmax, nq, m, m), dtype=store.dtype, device=store.device)
        padded[:nloc] = store
        store = padded
    allp = torch.empty((world, cmax, nq, m, m), dtype=store.dtype, device=store.device)
    dist.all_gather_into_tensor(allp.view(world * cmax, nq, m, m), store, group=group)
    return _global_order(allp, counts, layout).permute(2, 3, 0, 1)


# --------------------------------------------------------------------------------------------
# the same gather through the library's own C ABI (tkb_shard_init / tkb_allgather_matrices): what a
# non-Python host would call.  torch.distributed is only used to hand the NCCL id to the other ranks.

_MATRIX_TYPES = {"matrix_VNS": 0, "matrix_VSG": 1, "matrix_ANS": 2, "matrix_ASG": 3}


def init_capi(group=None):
    """Create the library's NCCL communicator over the ranks of `group` (rank 0 makes the id)."""
    import ctypes as C
    import torch.distributed as dist
    from . import _lib
    _lib.require_cuda()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(_lib.lib().tkb_shard_unique_id(buf))
    box = [buf.raw if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    _lib.check(_lib.lib().tkb_shard_init(box[0], rank, world))


def gather_matrix_capi(name, counts, ns_type=1, out=None, layout="strided", k_first=0, stream=None):
    """All-gather one matrix family along xi through the library's own NCCL communicator.  name: 'matrix_VNS' |
    'matrix_VSG' | 'matrix_ANS' | 'matrix_ASG'; counts: xi counts of all ranks; out: optional contiguous
    (nxi_total, nQ2, m, m) CUDA storage to gather into.  Returns (m, m, nxi_total, nQ2), xi in the order of the
    global list.  layout "strided" with equal counts: tkb_allgather_matrices_strided (per-xi all-gathers straight
    into place; k_first > 0 gathers a later chunk of the rank's xi into the same `out`); layout "blocks":
    tkb_allgather_matrices.  A strided split with unequal counts goes through torch.distributed (gather_matrix)."""
    import ctypes as C
    from . import _lib
    from . import matrices
    L = _lib.lib()
    n, nq = L.tkb_get_nx(), L.tkb_get_nq2()
    m = n * (2 if name.endswith("SG") else 1)
    total = int(sum(counts))
    if layout == "strided" and len(set(counts)) > 1:
        M = getattr(matrices, name)(ns_type) if name.endswith("NS") else getattr(matrices, name)()
        return gather_matrix(M, counts=list(counts), layout="strided")
    if out is None:
        out = torch.empty((total, nq, m, m), dtype=torch.float64, device=_lib.require_cuda())
    elif tuple(out.shape) != (total, nq, m, m) or out.dtype != torch.float64 or not out.is_cuda or not out.is_contiguous():
        raise ValueError("gather_matrix_capi: out must be a contiguous float64 CUDA tensor of shape %s" % ((total, nq, m, m),))
    st = _lib.stream_ptr() if stream is None else C.c_void_p(stream.cuda_stream)
    if layout == "strided":
        _lib.check(L.tkb_allgather_matrices_strided(_MATRIX_TYPES[name], int(ns_type), int(k_first),
                                                    C.c_void_p(out.data_ptr()), st))
    else:
        cnt = (C.c_int * len(counts))(*[int(c) for c in counts])
        _lib.check(L.tkb_allgather_matrices(_MATRIX_TYPES[name], int(ns_type), cnt, C.c_void_p(out.data_ptr()), st))
    return out.permute(2, 3, 0, 1)


# --------------------------------------------------------------------------------------------
# the exchange by the copy engines (tkb_p2p_*): no NCCL kernel, no SMs, so it runs under the QAGS kernels of the next
# chunk / step.  torch.distributed only carries the 64-byte CUDA IPC handles of the receive buffers.

_P2P_SLOTS = {"matrix_VNS+": 0, "matrix_VNS-": 1, "matrix_VSG": 2, "matrix_ASG": 3}
_P2P_FAMILY = {0: (0, 1), 1: (0, -1), 2: (1, 1), 3: (3, 1)}        # slot -> (tkb type, nstype)


def init_p2p(nx, nxi_total, nq2, group=None, strict=True):
    """Allocate this rank's four receive buffers (VNS+, VNS-, VSG, ASG; [xi_total][Q2][m][m]) and map every peer's.
    Returns the four buffers as tensors of logical shape (m, m, nxi_total, nQ2).  Every step is agreed on by all ranks
    (a failure on one rank -- e.g. CUDA IPC not permitted between the processes -- is a failure on all, so nobody is left
    waiting in a collective); strict=False returns None instead of raising, for callers with another exchange to fall
    back to."""
    import ctypes as C
    import torch.distributed as dist
    from . import _lib
    L = _lib.lib()
    dev = _lib.require_cuda()
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def agreed(ok):
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return int(flag.item()) == 1

    views, handles, err = [], [], None
    try:
        for slot in range(4):
            m = nx * (2 if slot >= 2 else 1)
            addr = C.c_void_p()
            _lib.check(L.tkb_p2p_alloc(slot, int(nxi_total) * int(nq2) * m * m * 8, C.byref(addr)))
            views.append(_lib.view_tensor(addr.value, (nxi_total, nq2, m, m)).permute(2, 3, 0, 1))
            buf = C.create_string_buffer(64)
            _lib.check(L.tkb_p2p_handle(slot, buf))
            handles.append(buf.raw)
    except RuntimeError as exc:
        err = exc
    if agreed(err is None):
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=group)
        try:
            for slot in range(4):
                _lib.check(L.tkb_p2p_open(slot, rank, world, b"".join(gathered[r][slot] for r in range(world))))
        except RuntimeError as exc:
            err = exc
        if agreed(err is None):
            return views
    L.tkb_p2p_close()
    if strict:
        raise RuntimeError("tiktaalik_b200.shard.init_p2p failed on at least one rank" + (": %s" % err if err else ""))
    return None


def push_matrices_p2p(k_first=0, stream=None):
    """Push this rank's current four matrix families into every rank's receive buffers (strided split: local xi k ->
    global position (k_first + k) * world + rank).  Asynchronous on `stream` (default: the current stream); the
    gathered arrays are complete once every rank's push has finished (dist.barrier after a stream synchronize)."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    st = _lib.stream_ptr() if stream is None else C.c_void_p(stream.cuda_stream)
    for slot, (t, ns) in _P2P_FAMILY.items():
        _lib.check(L.tkb_p2p_push_matrices(slot, t, ns, int(k_first), st))


def close_p2p():
    from . import _lib
    _lib.check(_lib.lib().tkb_p2p_close())
