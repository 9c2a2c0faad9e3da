// shard.cpp -- the one exchange step of the xi-sharded path: an all-gather of the evolution
// matrices along xi over NCCL / NVLink.  Assembly, stepping and application need no collective: every loop of
// the reference is independent across xi (matevo.f90:291,325,359,381,426,452,514,550,584; evolution.py:507-509).
//
// The library stores M as [xi][Q2][i][j], so a rank's contribution is one contiguous slab and the gathered
// array is the concatenation of the slabs in rank order.  Equal blocks go through ncclAllGather; when the block
// sizes differ between ranks the gather is one NCCL group of `world` broadcasts (root r writes slab r).
//
// A xi list sharded by STRIDE (rank r holds xi[r], xi[r + world], ...: every rank sees the same mix of the xi-dependent
// QAGS cost, which varies by 1.5x along the benchmark's list) is gathered straight into the order of the global list
// by one NCCL group of per-xi all-gathers: the all-gather of local xi k lands in out[k * world + 0 .. world - 1].
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy torch has
// already loaded) so that the library has no link-time dependency on it and single-GPU use never touches it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/tiktaalik_b200.h"
#include "tkb_internal.h"

namespace {

struct UniqueId { char internal[128]; };                    // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* Comm;                                         // ncclComm_t
enum { kNcclFloat64 = 8 };                                  // ncclDataType_t: ncclFloat64 / ncclDouble

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    Comm comm = nullptr;
    int rank = 0, world = 1;
} N;

int nccl_fail(const char* what, int rc)
{
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, N.GetErrorString ? N.GetErrorString(rc) : "NCCL error");
    return tkb::set_error(buf);
}

int load_nccl()
{
    if (N.handle) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return tkb::set_error("tkb_shard: libnccl.so.2 not found (multi-GPU gather needs NCCL)");
#define TKB_SYM(field, name)                                                                  \
    *(void**)(&N.field) = dlsym(h, name);                                                     \
    if (!N.field) { dlclose(h); return tkb::set_error("tkb_shard: NCCL symbol " name " missing"); }
    TKB_SYM(GetUniqueId, "ncclGetUniqueId")
    TKB_SYM(CommInitRank, "ncclCommInitRank")
    TKB_SYM(CommDestroy, "ncclCommDestroy")
    TKB_SYM(GroupStart, "ncclGroupStart")
    TKB_SYM(GroupEnd, "ncclGroupEnd")
    TKB_SYM(Broadcast, "ncclBroadcast")
    TKB_SYM(AllGather, "ncclAllGather")
    TKB_SYM(GetErrorString, "ncclGetErrorString")
#undef TKB_SYM
    N.handle = h;
    return 0;
}

}  // namespace

extern "C" {

int tkb_shard_unique_id(char id_out[128])
{
    if (!id_out) return tkb::set_error("tkb_shard_unique_id: null pointer");
    if (load_nccl()) return 1;
    UniqueId id;
    const int rc = N.GetUniqueId(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id_out, id.internal, 128);
    return 0;
}

int tkb_shard_init(const char id_in[128], int rank, int world)
{
    if (!id_in || world < 1 || rank < 0 || rank >= world) return tkb::set_error("tkb_shard_init: bad arguments");
    if (load_nccl()) return 1;
    if (N.comm) { N.CommDestroy(N.comm); N.comm = nullptr; }
    UniqueId id;
    memcpy(id.internal, id_in, 128);
    const int rc = N.CommInitRank(&N.comm, world, id, rank);       // binds to the current CUDA device
    if (rc != 0) { N.comm = nullptr; return nccl_fail("ncclCommInitRank", rc); }
    N.rank = rank;
    N.world = world;
    return 0;
}

int tkb_shard_finalize(void)
{
    if (N.comm && N.CommDestroy) N.CommDestroy(N.comm);
    N.comm = nullptr;
    N.rank = 0;
    N.world = 1;
    return 0;
}

int tkb_allgather_matrices(int type, int nstype, const int* nxi_per_rank, double* dev_out, void* stream)
{
    if (!N.comm) return tkb::set_error("tkb_allgather_matrices: call tkb_shard_init first");
    if (!nxi_per_rank || !dev_out) return tkb::set_error("tkb_allgather_matrices: null pointer");
    const double* local = nullptr;
    int64_t shape[4], strides[4];
    if (tkb_evomatrix(type, nstype, &local, shape, strides)) return 1;
    const int64_t m = shape[0], nxi = shape[2], nq2 = shape[3];
    if (nxi_per_rank[N.rank] != nxi) return tkb::set_error("tkb_allgather_matrices: nxi_per_rank[rank] differs from the assembled xi count");
    const size_t per_xi = (size_t)nq2 * (size_t)m * (size_t)m;
    bool even = true;
    for (int r = 0; r < N.world; r++) even = even && (nxi_per_rank[r] == nxi_per_rank[0]);
    if (even) {         // equal blocks: one all-gather (all NVLink ports at once; 8 grouped broadcasts reach a fraction of it)
        const int rc1 = N.AllGather(local, dev_out, (size_t)nxi * per_xi, kNcclFloat64, N.comm, (cudaStream_t)stream);
        if (rc1 != 0) return nccl_fail("ncclAllGather", rc1);
        return 0;
    }
    int rc = N.GroupStart();
    if (rc != 0) return nccl_fail("ncclGroupStart", rc);
    size_t off = 0;
    for (int r = 0; r < N.world; r++) {
        if (nxi_per_rank[r] < 0) { N.GroupEnd(); return tkb::set_error("tkb_allgather_matrices: negative block size"); }
        const size_t count = (size_t)nxi_per_rank[r] * per_xi;
        if (count) {
            const void* send = (r == N.rank) ? (const void*)local : (const void*)(dev_out + off);
            rc = N.Broadcast(send, dev_out + off, count, kNcclFloat64, r, N.comm, (cudaStream_t)stream);
            if (rc != 0) { N.GroupEnd(); return nccl_fail("ncclBroadcast", rc); }
        }
        off += count;
    }
    rc = N.GroupEnd();
    if (rc != 0) return nccl_fail("ncclGroupEnd", rc);
    return 0;
}

int tkb_allgather_matrices_strided(int type, int nstype, int k_first, double* dev_out, void* stream)
{
    if (!N.comm) return tkb::set_error("tkb_allgather_matrices_strided: call tkb_shard_init first");
    if (!dev_out || k_first < 0) return tkb::set_error("tkb_allgather_matrices_strided: bad arguments");
    const double* local = nullptr;
    int64_t shape[4], strides[4];
    if (tkb_evomatrix(type, nstype, &local, shape, strides)) return 1;
    const int64_t m = shape[0], nxi = shape[2], nq2 = shape[3];
    const size_t per_xi = (size_t)nq2 * (size_t)m * (size_t)m;
    int rc = N.GroupStart();
    if (rc != 0) return nccl_fail("ncclGroupStart", rc);
    for (int64_t k = 0; k < nxi; k++) {
        // local xi k of every rank -> global positions (k_first + k) * world + rank: the global list's order
        rc = N.AllGather(local + (size_t)k * per_xi, dev_out + (size_t)(k_first + k) * (size_t)N.world * per_xi, per_xi,
                         kNcclFloat64, N.comm, (cudaStream_t)stream);
        if (rc != 0) { N.GroupEnd(); return nccl_fail("ncclAllGather", rc); }
    }
    rc = N.GroupEnd();
    if (rc != 0) return nccl_fail("ncclGroupEnd", rc);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The same exchange without NCCL and without SMs: every rank PUSHES its per-xi slabs into the receive buffers of all
// ranks with strided device-to-device copies (cudaMemcpy2DAsync: one call per peer and matrix family, row = one xi
// slab, destination pitch = world slabs), which the copy engines carry over NVLink / NVSwitch.  The QAGS kernels keep
// every SM's register file and shared memory for the whole assembly, so an NCCL kernel launched next to them only
// gets to run when they drain; a copy-engine transfer runs under them.  The receive buffers are cudaMalloc
// allocations of the library, exported to the other processes as CUDA IPC handles.

namespace {
constexpr int P2P_SLOTS = 4, P2P_MAXW = 16;
struct P2P {
    double* mine[P2P_SLOTS] = { nullptr, nullptr, nullptr, nullptr };
    size_t bytes[P2P_SLOTS] = { 0, 0, 0, 0 };
    double* peer[P2P_SLOTS][P2P_MAXW] = {};
    int rank = 0, world = 0;
} X;
}  // namespace

int tkb_p2p_alloc(int slot, int64_t bytes, double** dev_ptr)
{
    if (slot < 0 || slot >= P2P_SLOTS || bytes <= 0 || !dev_ptr) return tkb::set_error("tkb_p2p_alloc: bad arguments");
    if (X.mine[slot] && X.bytes[slot] != (size_t)bytes) {
        if (X.world) return tkb::set_error("tkb_p2p_alloc: buffer is mapped by peers, call tkb_p2p_close first");
        cudaFree(X.mine[slot]);
        X.mine[slot] = nullptr;
    }
    if (!X.mine[slot]) {
        const cudaError_t e = cudaMalloc(&X.mine[slot], (size_t)bytes);
        if (e != cudaSuccess) return tkb::set_error(cudaGetErrorString(e));
        X.bytes[slot] = (size_t)bytes;
    }
    *dev_ptr = X.mine[slot];
    return 0;
}

int tkb_p2p_handle(int slot, char handle_out[64])
{
    if (slot < 0 || slot >= P2P_SLOTS || !X.mine[slot] || !handle_out) return tkb::set_error("tkb_p2p_handle: no buffer in this slot");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, X.mine[slot]);
    if (e != cudaSuccess) return tkb::set_error(cudaGetErrorString(e));
    memcpy(handle_out, &h, 64);
    return 0;
}

int tkb_p2p_open(int slot, int rank, int world, const char* handles_all)
{
    if (slot < 0 || slot >= P2P_SLOTS || world < 1 || world > P2P_MAXW || rank < 0 || rank >= world || !handles_all || !X.mine[slot])
        return tkb::set_error("tkb_p2p_open: bad arguments");
    for (int p = 0; p < world; p++) {
        if (p == rank) { X.peer[slot][p] = X.mine[slot]; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles_all + (size_t)p * 64, 64);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            char buf[256];
            snprintf(buf, sizeof buf, "cudaIpcOpenMemHandle (rank %d -> buffer of rank %d): %s", rank, p, cudaGetErrorString(e));
            return tkb::set_error(buf);
        }
        X.peer[slot][p] = (double*)ptr;
    }
    X.rank = rank;
    X.world = world;
    return 0;
}

static int p2p_mode_kernel()
{
    // default: pitched peer copies by the driver / copy engines; TKB_P2P_COPY=kernel: exchange.cu's peer-store kernel
    // (measured on 8 B200, C3: 147.5 ms per step with the copies, 154.2 ms with the kernel, which takes issue slots and
    // memory bandwidth from the QAGS kernels it runs beside and reaches 510 GB/s alone against 675 GB/s)
    static const int v = [] { const char* e = getenv("TKB_P2P_COPY"); return (e && e[0] == 'k') ? 1 : 0; }();
    return v;
}

int tkb_p2p_push_matrices(int slot, int type, int nstype, int k_first, void* stream)
{
    if (slot < 0 || slot >= P2P_SLOTS || !X.world || !X.peer[slot][X.rank] || k_first < 0)
        return tkb::set_error("tkb_p2p_push_matrices: call tkb_p2p_alloc / tkb_p2p_open first");
    const double* local = nullptr;
    int64_t shape[4], strides[4];
    if (tkb_evomatrix(type, nstype, &local, shape, strides)) return 1;
    const int64_t m = shape[0], nxi = shape[2], nq2 = shape[3];
    const size_t slab = (size_t)nq2 * (size_t)m * (size_t)m * sizeof(double);
    if (((size_t)(k_first + nxi) * (size_t)X.world) * slab > X.bytes[slot])
        return tkb::set_error("tkb_p2p_push_matrices: the receive buffer is too small for this chunk");
    // local xi k -> position (k_first + k) * world + rank of every rank's buffer (own copy first)
    if (p2p_mode_kernel())
        return tkb::launch_push_slabs(local, X.peer[slot], slab / sizeof(double), (int)nxi, X.world, X.rank, k_first, (cudaStream_t)stream);
    for (int q = 0; q < X.world; q++) {
        const int p = (X.rank + q) % X.world;
        char* dst = (char*)X.peer[slot][p] + ((size_t)k_first * (size_t)X.world + (size_t)X.rank) * slab;
        const cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)X.world * slab, local, slab, slab, (size_t)nxi,
                                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) return tkb::set_error(cudaGetErrorString(e));
    }
    return 0;
}

int tkb_p2p_close(void)
{
    for (int s = 0; s < P2P_SLOTS; s++) {
        for (int p = 0; p < X.world; p++)
            if (p != X.rank && X.peer[s][p]) cudaIpcCloseMemHandle(X.peer[s][p]);
        for (int p = 0; p < P2P_MAXW; p++) X.peer[s][p] = nullptr;
        if (X.mine[s]) cudaFree(X.mine[s]);
        X.mine[s] = nullptr;
        X.bytes[s] = 0;
    }
    X.world = 0;
    X.rank = 0;
    return 0;
}

}  // extern "C"
