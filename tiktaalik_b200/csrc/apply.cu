// apply.cu -- applying stored evolution matrices to GPD samples on sm_100a.
//
// Reference path replaced (file:line): np.einsum('ijkl,jk...->ik...l', M, gpd) in
// tiktaalik/evolution.py:259,270,281,326,337,348,507-509 and examples/examples.py:77,96.
//     out[i, k, s, q] = sum_j M[i, j, k, q] * g[j, k, s]
// Device layouts (shape/index semantics are the API, strides are free -- SURVEY.md 8b):
//     M   [xi][Q2][i][j]   row-major m x m tiles (what make_matrices produces)
//     g   [j][xi][s]       the caller's (m, nxi, S) C-contiguous array, consumed in place
//     out [xi][Q2][i][s]   returned to Python as a permuted view of shape (m, nxi, S, nQ2)
// Two regimes: HBM-bound on reading M up to S ~ 24 (0.25..6 flop/B), FP64-bound beyond.  Both run
// the batched FP64 tensor-core GEMM of rk4.cu (cp.async, three stages) with a tile as narrow as the sample count
// (16 columns up to S = 16, so M is streamed exactly once) and the right-hand side shared by the nQ2 matrices of
// one xi.  The warp-per-row kernel below (S <= 8) was the first small-S path; it is kept behind
// TKB_APPLY_SMALL_MAX and serves matrices with fewer than 40 rows.
#include <cstdlib>
#include <cuda_runtime.h>
#include "tkb_internal.h"

namespace tkb {

constexpr int GEMV_MAXS = 8;
constexpr int GEMV_WARPS = 8;

template <int S>
__global__ void __launch_bounds__(GEMV_WARPS * 32)
apply_small_kernel(int m, int nq2, const double* __restrict__ M, const double* __restrict__ g, int64_t ldg,
                   int64_t xig, double* __restrict__ out)
{
    extern __shared__ double s_g[];                 // [S][m]: sample-major so that lanes read consecutive words
    const int b = blockIdx.x;                       // b = k * nq2 + q
    const int k = b / nq2;
    const double* gk = g + (int64_t)k * xig;
    for (int e = threadIdx.x; e < m * S; e += blockDim.x) {
        const int j = e / S, s = e - j * S;
        s_g[s * m + j] = gk[(int64_t)j * ldg + s];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* Mb = M + (int64_t)b * m * m;
    double* ob = out + (int64_t)b * m * S;
    // rows split over gridDim.y slices so that small batches still fill the GPU
    for (int i = blockIdx.y * GEMV_WARPS + warp; i < m; i += gridDim.y * GEMV_WARPS) {
        const double2* row = reinterpret_cast<const double2*>(Mb + (int64_t)i * m);
        double acc[S];
#pragma unroll
        for (int s = 0; s < S; s++) acc[s] = 0.0;
        for (int j2 = lane; j2 < m / 2; j2 += 32) {
            const double2 v = row[j2];
#pragma unroll
            for (int s = 0; s < S; s++) {
                const double2 gv = reinterpret_cast<const double2*>(s_g + s * m)[j2];
                acc[s] = fma(v.x, gv.x, acc[s]);
                acc[s] = fma(v.y, gv.y, acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            double a = acc[s];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
            acc[s] = a;
        }
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < S; s++) ob[(int64_t)i * S + s] = acc[s];
        }
    }
}

template <int S>
static int launch_small(int m, int nb, int nq2, const double* M, const double* g, int64_t ldg, int64_t xig, double* out,
                        cudaStream_t st)
{
    int ysl = 1;
    while ((long long)nb * ysl < 148LL * 4 && ysl * GEMV_WARPS < m) ysl *= 2;
    dim3 grid(nb, ysl);
    const size_t smem = (size_t)m * S * sizeof(double);
    apply_small_kernel<S><<<grid, GEMV_WARPS * 32, smem, st>>>(m, nq2, M, g, ldg, xig, out);
    cudaError_t e = cudaGetLastError();
    ++g_launches;
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

int launch_apply(int m, int nxi, int nq2, int64_t S, const double* M, const double* g, int64_t ldg, int64_t xig,
                 double* out, cudaStream_t st)
{
    if (S <= 0) return 0;
    const int nb = nxi * nq2;
    static const int small_max = [] { const char* e = getenv("TKB_APPLY_SMALL_MAX"); const int v = e ? atoi(e) : 0;   /* measured: the cp.async DMMA GEMM with a 16-column tile is faster for every S */ return v < 0 ? 0 : (v > GEMV_MAXS ? GEMV_MAXS : v); }();
    if ((S <= small_max || (m < 40 && S <= GEMV_MAXS)) && (m % 2) == 0 && (size_t)m * S * sizeof(double) <= 48 * 1024) {
        switch ((int)S) {
        case 1: return launch_small<1>(m, nb, nq2, M, g, ldg, xig, out, st);
        case 2: return launch_small<2>(m, nb, nq2, M, g, ldg, xig, out, st);
        case 3: return launch_small<3>(m, nb, nq2, M, g, ldg, xig, out, st);
        case 4: return launch_small<4>(m, nb, nq2, M, g, ldg, xig, out, st);
        case 5: return launch_small<5>(m, nb, nq2, M, g, ldg, xig, out, st);
        case 6: return launch_small<6>(m, nb, nq2, M, g, ldg, xig, out, st);
        case 7: return launch_small<7>(m, nb, nq2, M, g, ldg, xig, out, st);
        default: return launch_small<8>(m, nb, nq2, M, g, ldg, xig, out, st);
        }
    }
    if (S > 0x7fffffffLL || ldg > 0x7fffffffLL) return set_error("tkb_apply: S too large for one call; tile over samples");
    // C[b] (m x S) = M[b] (m x m) * g_k (m x S), b = k*nq2 + q, right-hand side shared by nq2 consecutive b
    return launch_gemm(m, (int)S, m, 1.0, M, m, (int64_t)m * m, g, (int)ldg, xig, nq2, nullptr, 0, 0,
                       out, (int)S, (int64_t)m * S, nb, st);
}

}  // namespace tkb
