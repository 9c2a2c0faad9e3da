"""
shard.py -- multi-GPU partitioning of the hot path by xi (one process per GPU).

Every xi is independent in assembly and stepping (matevo.f90:291,325,359,381,426,452,514,550,584)
and in application (evolution.py:507-509), so each rank builds the matrices of a contiguous
block of xi values with no data-path collective.  The only exchange is the optional final
all-gather of the per-xi matrices (NCCL over NVLink through torch.distributed): the library
stores M as [xi][Q2][i][j], so a rank's contribution is one contiguous slab.

The reference clamps only the global first and last xi (matevo.f90:688-689); the clamp is
therefore applied before sharding and each rank is told whether it holds an end point.
"""
import numpy as np
import torch

from .f90wrap.matrices import dummy as f90src


def xi_partition(nxi, world_size):
    """contiguous block sizes, earlier ranks take the remainder"""
    base, rem = divmod(int(nxi), int(world_size))
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


def local_xi(xi, rank, world_size):
    xi = clamp_xi(xi)
    counts, offsets = xi_partition(xi.shape[0], world_size)
    return xi[offsets[rank]:offsets[rank] + counts[rank]], counts, offsets


def initialize_kernels_sharded(nx, xi, grid_type=1, rank=0, world_size=1):
    """matrices.initialize_kernels for this rank's block of the global xi list."""
    assert(nx==(nx//2)*2)
    assert(nx>=6)
    mine, counts, offsets = local_xi(xi, rank, world_size)
    if mine.shape[0] == 0:
        raise ValueError("rank %d has no xi values (nxi < world_size)" % rank)
    # the list is already clamped, so no further clamp is applied on any rank
    f90src.make_kernels_shard(nx, mine, grid_type, 0, 0)
    return mine


def gather_matrix(M_local, counts=None, group=None):
    """All-gather per-xi matrices along xi.  M_local: (m, m, nxi_local, nQ2) as returned by the
    matrix getters.  Returns (m, m, nxi_total, nQ2) on every rank.  World size 1: returns M_local."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return M_local
    world = dist.get_world_size(group)
    m, _, nloc, nq = M_local.shape
    store = M_local.permute(2, 3, 0, 1).contiguous()          # [xi][Q2][i][j], one slab
    if counts is None:
        cnt = torch.tensor([nloc], dtype=torch.int64, device=store.device)
        allc = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(allc, cnt, group=group)
        counts = [int(c.item()) for c in allc]
    if len(set(counts)) == 1:
        out = torch.empty((world * nloc, nq, m, m), dtype=store.dtype, device=store.device)
        dist.all_gather_into_tensor(out, store, group=group)
    else:
        # collectives need equal contributions: pad every slab to the largest block, trim afterwards
        cmax = max(counts)
        padded = torch.zeros((cmax, nq, m, m), dtype=store.dtype, device=store.device)
        padded[:nloc] = store
        allp = torch.empty((world * cmax, nq, m, m), dtype=store.dtype, device=store.device)
        dist.all_gather_into_tensor(allp, padded, group=group)
        out = torch.cat([allp[r * cmax:r * cmax + c] for r, c in enumerate(counts)], dim=0)
    return out.permute(2, 3, 0, 1)


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


def gather_matrix_capi(name, counts, ns_type=1, out=None):
    """All-gather one matrix family along xi with tkb_allgather_matrices.  name: 'matrix_VNS' | 'matrix_VSG' |
    'matrix_ANS' | 'matrix_ASG'; counts: xi block sizes of all ranks; out: optional contiguous
    (nxi_total, nQ2, m, m) CUDA storage to gather into.  Returns (m, m, nxi_total, nQ2)."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    n, nq = L.tkb_get_nx(), L.tkb_get_nq2()
    m = n * (2 if name.endswith("SG") else 1)
    total = int(sum(counts))
    if out is None:
        out = torch.empty((total, nq, m, m), dtype=torch.float64, device=_lib.require_cuda())
    elif tuple(out.shape) != (total, nq, m, m) or out.dtype != torch.float64 or not out.is_cuda or not out.is_contiguous():
        raise ValueError("gather_matrix_capi: out must be a contiguous float64 CUDA tensor of shape %s" % ((total, nq, m, m),))
    cnt = (C.c_int * len(counts))(*[int(c) for c in counts])
    _lib.check(L.tkb_allgather_matrices(_MATRIX_TYPES[name], int(ns_type), cnt, C.c_void_p(out.data_ptr()),
                                        _lib.stream_ptr()))
    return out.permute(2, 3, 0, 1)
