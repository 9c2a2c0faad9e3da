// exchange.cu -- the xi all-gather as ONE small-footprint kernel that stores straight into peer memory over NVLink
// (alternative form of tkb_p2p_push_matrices, TKB_P2P_COPY=kernel; the default are pitched peer copies, which
// measured faster: see shard.cpp).
//
// The QAGS kernels hold every SM for the whole assembly with one 768-thread block of 80 registers per thread (61 440
// of the SM's 65 536 registers, 170 KB of its shared memory).  An NCCL kernel launched next to them waits until they
// drain, and the driver runs a pitched peer cudaMemcpy2DAsync the same way (measured: 7 ms of a 148 ms step exposed
// on 8 GPUs).  A block of 128 threads at 32 registers and no shared memory is exactly what still fits beside a QAGS
// block, so this kernel runs UNDER the next chunk's assembly: each thread moves 16-byte words from the rank's own
// per-xi slabs to the same global position of every rank's receive buffer (peer pointers from CUDA IPC handles).
#include <cuda_runtime.h>
#include "tkb_internal.h"

namespace tkb {

struct PeerPtrs { double* p[16]; };

// local[k][slab] -> peer q: base_q[((k_first + k) * world + rank) * slab ...], k = 0 .. nxi - 1; slab in 16-byte words
__global__ void __launch_bounds__(128, 16) push_slabs_kernel(const double2* __restrict__ local, PeerPtrs peers, long long slab2, int nxi,
                                                        int world, int rank, int k_first)
{
    const long long total = (long long)nxi * slab2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int q = 0; q < world; q++) {
        const int peer = (rank + q) % world;                 // every rank starts with its own buffer, then its neighbours
        double* base = nullptr;
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (j == peer) base = peers.p[j];                // constant indices: the pointers stay in the parameter space
        double2* dst = reinterpret_cast<double2*>(base);
        const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        long long k = e0 / slab2, w = e0 - k * slab2;         // one division per thread and peer, then incremental
        for (long long e = e0; e < total; e += stride) {
            dst[((long long)(k_first + k) * world + rank) * slab2 + w] = local[e];
            w += stride;
            while (w >= slab2) { w -= slab2; k++; }
        }
    }
}

int launch_push_slabs(const double* local, double* const* peer_base, size_t slab_doubles, int nxi, int world, int rank,
                      int k_first, cudaStream_t st)
{
    if (world > 16) return set_error("push_slabs: at most 16 ranks");
    if (slab_doubles % 2 != 0) return set_error("push_slabs: odd slab size");
    PeerPtrs P;
    for (int q = 0; q < 16; q++) P.p[q] = (q < world) ? peer_base[q] : nullptr;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    push_slabs_kernel<<<sms, 128, 0, st>>>(reinterpret_cast<const double2*>(local), P, (long long)(slab_doubles / 2), nxi, world,
                                            rank, k_first);
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

}  // namespace tkb
