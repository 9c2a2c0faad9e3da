// assemble.cu -- kernel-matrix assembly on sm_100a: K1 fixed Gauss-Kronrod quadrature,
// K2 batched QAGS for the NLO plus-prescription constants, K3 composition into the
// nfl-dependent kernel matrices.  Compiled with -fmad=false (see tkb_fp.cuh).
//
// Reference path replaced (file:line):
//   matevo.f90:283-480   make_kernels_NS_0 / NS_1 / SG_0 / SG_1  (24 pixel_conv per (xi,i,j))
//   convolution.f90:26-97 pixel_conv = reg + pls + cst
//   integration.f90:18-49 integrate = 6 x qk21 (126 fixed nodes), :51-64 adaptive_integrate = 3 x qags
//   kernels_nlo.f90:60-78 ... 925-944  the eight *_cst functions
//
// Design (differs from the reference on purpose):
//   * the integrand factorises over the column j: for fixed (xi, i) the 126 nodes y_k and all
//     22 kernel values K_f(x_i, y_k, xi) do not depend on j, only the basis value f_j(y_k) does,
//     and it is non-zero for 6 consecutive j.  K1 evaluates every kernel ONCE per node
//     (one thread per (node, half), shared transcendental values, tkb_kernels.cuh) and then
//     replays qk21's accumulation per column from shared memory.  Per-term values and the
//     summation order are those of the reference, so entries agree bit for bit with a
//     straight evaluation, at 1/nx of the kernel evaluations.
//   * K2 runs batches of QAGS instances per block, split by role (assemble_cst_pipe.cuh): the 2 x 21 nodes
//     of a bisection step are 168 half-kernel tasks evaluated by the block's evaluator warps, the scalar
//     QUADPACK bookkeeping (qsort, qextr) runs on bookkeeping warps, one lane pair per instance.  Control
//     flow is QUADPACK's, comparison for comparison, so the bisection path is the reference's.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tkb_fp.cuh"
#include "tkb_kernels.cuh"
#include "tkb_grid.cuh"
#include "tkb_internal.h"

namespace tkb {

// Gauss-Kronrod 21 abscissae / weights (integration.f90:139-152), 1-based like the reference
__constant__ double c_xgk[12] = { 0.0,
    9.956571630258081e-01, 9.739065285171717e-01, 9.301574913557082e-01, 8.650633666889845e-01,
    7.808177265864169e-01, 6.794095682990244e-01, 5.627571346686047e-01, 4.333953941292472e-01,
    2.943928627014602e-01, 1.488743389816312e-01, 0.000000000000000 };
__constant__ double c_wgk[12] = { 0.0,
    1.169463886737187e-02, 3.255816230796473e-02, 5.475589657435200e-02, 7.503967481091995e-02,
    9.312545458369761e-02, 1.093871588022976e-01, 1.234919762620659e-01, 1.347092173114733e-01,
    1.427759385770601e-01, 1.477391049013385e-01, 1.494455540029169e-01 };
__constant__ double c_wg[6] = { 0.0,
    6.667134430868814e-02, 1.494513491505806e-01, 2.190863625159820e-01, 2.692667193099964e-01,
    2.955242247147529e-01 };

// plus-type kernels: integrand K (f_j(y) - f_j(x)); the rest are regular: K f_j(y)
__host__ __device__ constexpr bool is_plus(int k)
{
    return k == K_K0QQ || k == K_A0GG || k == K_NSP || k == K_NSM || k == K_NSN ||
           k == K_V1GG || k == K_A1GG || k == K_V1GGN || k == K_A1GGN;
}

// position q (0..20) in the per-interval node table -> eta.  q = 0 centre; q = 2t-1 / 2t are
// centr -/+ hlgth*xgk(t), t = 1..10 (qk21, integration.f90:174-203)
TKB_DEV double node_eta(double centr, double hlgth, int q)
{
    if (q == 0) return centr;
    const int t = (q + 1) >> 1;
    const double absc = hlgth * c_xgk[t];
    return (q & 1) ? (centr - absc) : (centr + absc);
}

// ---------------------------------------------------------------------------------------
// K1: fixed quadrature.  grid = (ceil(n / K1_ROWS), nxi), block = 128 K1_ROWS threads, K1_ROWS rows (xi, i) per block.
// P layout: P[f][ixi][i][j] (row-major n x n per xi).  diag layout: diag[ixi][i][DIAG_N].

// One-argument logs / dilogs of the fixed point x_i (eval_half, FIX): the same for the 126 nodes of row (xi, i) in K1 and
// for the 21 QAGS instances of that row in K2, so they are computed once per row by a small kernel and read from
// this table: PRE_ROW doubles per row: [half][PRE_N] for U = 0.5 (1 +- x_i / xi), then x_i and the two U (which the
// QAGS bookkeeping would otherwise recompute, pow included, at each of the row's 21 instance starts).
constexpr int PRE_ROW = 16, PRE_X = 10, PRE_U1 = 11, PRE_U2 = 12;
__device__ const double* d_pre_tab = nullptr;

__global__ void pre_table_kernel(int n, int nxi, const double* __restrict__ xi_arr, int grid_type, double* __restrict__ tab)
{
    log_tab_init();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * nxi * PRE_ROW) return;
    const int row = e / PRE_ROW, v = e - row * PRE_ROW;
    const int ixi = row / n, i = row - ixi * n + 1;
    const double xi = xi_arr[ixi];
    const double x = grid_node(i, n, xi, grid_type);
    const double U1 = 0.5 * (1.0 + x / xi), U2 = 0.5 * (1.0 - x / xi);
    double val = 0.0;
    if (v < 2 * PRE_N) val = pre_value(v % PRE_N, (v / PRE_N) ? U2 : U1);
    else if (v == PRE_X) val = x;
    else if (v == PRE_U1) val = U1;
    else if (v == PRE_U2) val = U2;
    tab[e] = val;
}

#ifdef TKB_K2_STATS
__device__ unsigned long long g_k2stats[24];      // in-kernel cycle counters of the stats build (K1 and K2)
#define K1_CLK(slot) do { if (tid == 0) { const long long n_ = clock64(); atomicAdd(&g_k2stats[slot], (unsigned long long)(n_ - tc_)); tc_ = n_; } } while (0)
#else
#define K1_CLK(slot)
#endif
#ifndef TKB_K1_ROWS
#define TKB_K1_ROWS 4
#endif
constexpr int K1_ROWS = TKB_K1_ROWS;           // rows (xi, i) per block
constexpr int K1_THREADS = 128 * K1_ROWS;
constexpr int NNODE = 126;

// Shared memory of one block: K1_ROWS rows.  A row has 2 x 126 half-kernel tasks of which about half are live, i.e.
// four warps' worth; with one row per block (round 1) the other four warps of the block waited at the barrier while
// holding their 128 registers, so an SM had 8 live warps in the evaluation phase -- a latency-bound chain of several
// thousand dependent instructions per task (ncu: 13 warps/SM, issue slots 34 % busy, barrier 2.9 and wait 3.0
// warp-cycles per issue, 16.5 of 32 lanes active).  With K1_ROWS = 4 rows per 512-thread block every warp has live
// tasks and the SM runs 16 of them.
struct K1Shared {
    double H[K1_ROWS][2][NKERN][NNODE];   // half-kernel values; after the combination pass [r][0][kernel][node] holds K(x_i, y_k)
    double Y[K1_ROWS][NNODE][2];          // 0.5 (1 +- y/xi) at the nodes
    double J[K1_ROWS][NNODE];             // Jacobians
    double b[K1_ROWS][NNODE][6];          // basis rows at the nodes
    double aa[K1_ROWS][8], hl[K1_ROWS][6], bx[K1_ROWS][6];
    double pre[K1_ROWS][2][PRE_N];        // one-argument logs/dilogs of the row's fixed point (eval_half, FIX = 1)
    double w[6];                          // Lagrange weights (the same for every row)
    int iS[K1_ROWS][NNODE];               // stencil origins (1-based)
    int jlo[K1_ROWS][6], jhi[K1_ROWS][6]; // per integration interval: smallest / largest stencil origin of its 21 nodes
    int iSx[K1_ROWS];
    int allfinite[K1_ROWS];
    int nlive;
    int next2;                            // phase 2: next 32 items of the replay work list
    unsigned short live[K1_ROWS * 2 * NNODE];
};

__global__ void __launch_bounds__(K1_THREADS, (K1_ROWS <= 2) ? 2 : 1)
assemble_fixed_kernel(int n, int nxi, const double* __restrict__ xi_arr, int grid_type,
                      double* __restrict__ P, double* __restrict__ diag)
{
    extern __shared__ __align__(16) unsigned char k1_smem[];
    K1Shared& sh = *reinterpret_cast<K1Shared*>(k1_smem);
    log_tab_init();
    const int tid = threadIdx.x;
    const int ixi = blockIdx.y;
    const int i0 = blockIdx.x * K1_ROWS;           // 0-based first row of the block
    const int nrows = (n - i0 < K1_ROWS) ? (n - i0) : K1_ROWS;
    const double xi = xi_arr[ixi];
    const double* tab0 = d_pre_tab + ((size_t)ixi * n + (size_t)i0) * PRE_ROW;   // filled by pre_table_kernel
#ifdef TKB_K2_STATS
    long long tc_ = clock64();
#endif

    // set-up, spread over a few threads of different warps (every piece is a chain of dependent log / pow / divisions):
    // integration.f90:27-35 breakpoints in eta space; Lagrange weights; eta of x_i; the row's one-argument logs/dilogs
    {
        const int r = tid >> 7, t = tid & 127;      // 128 threads per row
        if (r < nrows) {
            const double* rowtab = tab0 + (size_t)r * PRE_ROW;
            const double x = rowtab[PRE_X];
            if (t < 7) {
                const double ax = fabs(x), axi = fabs(xi);
                const double arg = (t == 0) ? -1.0 : (t == 1) ? -ax : (t == 2) ? -axi : (t == 3) ? 0.0
                                 : (t == 4) ? axi : (t == 5) ? ax : 1.0;
                sh.aa[r][t + 1] = pull_back(arg, xi, grid_type);
            } else if (t == 32 && r == 0) {
                double w[6];
                lagrange_weights(n, w);
                for (int m = 0; m < 6; m++) sh.w[m] = w[m];
            } else if (t == 64) {
                sh.aa[r][0] = pull_back(x, xi, grid_type);             // eta of x_i, for the basis row below
                sh.allfinite[r] = 1;
            } else if (t >= 96 && t < 96 + 2 * PRE_N) {
                const int v = t - 96;
                sh.pre[r][v / PRE_N][v % PRE_N] = rowtab[v];
            }
        }
        if (tid == 0) { sh.nlive = 0; sh.next2 = 0; }
    }
    __syncthreads();
    {
        const int r = tid >> 7, t = tid & 127;
        if (r < nrows) {
            if (t == 0) {
                double aa[8];
                for (int q = 1; q <= 7; q++) aa[q] = sh.aa[r][q];
                if (aa[3] < aa[2]) { const double z = aa[2]; aa[2] = aa[3]; aa[3] = z; }      // inner pairs sorted
                if (aa[6] < aa[5]) { const double z = aa[5]; aa[5] = aa[6]; aa[6] = z; }
                for (int q = 1; q <= 7; q++) sh.aa[r][q] = aa[q];
                for (int q = 0; q < 6; q++) sh.hl[r][q] = 5.0e-01 * (aa[q + 2] - aa[q + 1]);
            } else if (t == 32) {
                // basis row at x_i itself (interpixel(j, x): pull_back then Lagrange, convolution.f90:77,93)
                double w[6], bx[6];
                for (int m = 0; m < 6; m++) w[m] = sh.w[m];
                sh.iSx[r] = basis_row(sh.aa[r][0], n, w, bx);
                for (int m = 0; m < 6; m++) sh.bx[r][m] = bx[m];
            }
        }
    }
    __syncthreads();

    K1_CLK(0);
    // ---- phase 1a: one thread per (row, node, half): node, Jacobian / basis row, support class of the half ----
    // A half kernel whose support steps rho and theta_ERBL both vanish is an exact zero when every skipped piece
    // is finite (node_is_safe, as in K2); such halves -- about half of them -- are not evaluated, and the live
    // ones are compacted so that whole warps drop out of the evaluation pass.
    for (int e = tid; e < nrows * 2 * NNODE; e += K1_THREADS) {
        const int r = e / (2 * NNODE), t = e - r * 2 * NNODE;
        const int k = t >> 1, half = t & 1;
        const double* rowtab = tab0 + (size_t)r * PRE_ROW;
        const int iv = k / 21, q = k - 21 * iv;
        const double a = sh.aa[r][iv + 1], b = sh.aa[r][iv + 2];
        const double centr = 5.0e-01 * (a + b);
        const double hlgth = 5.0e-01 * (b - a);
        const double eta = node_eta(centr, hlgth, q);
        const double y = push_forward(eta, xi, grid_type);
        if (half == 0) {
            sh.J[r][k] = push_jacob(eta, xi, grid_type);
        } else {
            double w[6], bb[6];
#pragma unroll
            for (int m = 0; m < 6; m++) w[m] = sh.w[m];
            sh.iS[r][k] = basis_row(pull_back(y, xi, grid_type), n, w, bb);
#pragma unroll
            for (int m = 0; m < 6; m++) sh.b[r][k][m] = bb[m];
        }
        const double X1 = rowtab[PRE_U1], X2 = rowtab[PRE_U2];
        const double Y1 = 0.5 * (1.0 + y / xi), Y2 = 0.5 * (1.0 - y / xi);
        if (half == 0) { sh.Y[r][k][0] = Y1; sh.Y[r][k][1] = Y2; }
        const bool safe = node_is_safe(X1, X2, Y1, Y2);
        const int cls = half ? task_class(X2, Y2, safe) : task_class(X1, Y1, safe);
        if (cls != 0) {
            sh.live[atomicAdd(&sh.nlive, 1)] = (unsigned short)e;
        } else {
#pragma unroll
            for (int f = 0; f < NKERN; f++) sh.H[r][half][f][k] = 0.0;
        }
    }
    __syncthreads();
    K1_CLK(1);
    // range of columns an interval's nodes touch (phase 2 skips the other intervals of an ordinary column)
    if (tid < nrows * 6) {
        const int r = tid / 6, iv = tid - 6 * r;
        int lo = sh.iS[r][21 * iv], hi = lo;
        for (int q = 1; q < 21; q++) {
            const int v = sh.iS[r][21 * iv + q];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        sh.jlo[r][iv] = lo;
        sh.jhi[r][iv] = hi;
    }
    // ---- phase 1b: the live halves, all 22 kernels each, in eight passes of two or three kernels ----
    // One evaluation of all 22 kernels is ~100 KB of straight-line code; as a single pass every warp streamed it from
    // L2 on its own (the warps drift apart by more than the instruction cache holds).  Here the transcendental values
    // of the node, which all kernels share, are computed first (node_values: small hot code), then the kernels follow
    // group by group with a block barrier between the groups, so that the 16 warps run the same ~10 KB of formulas at
    // the same time.  Values and formulas are those of the single pass: bit-identical results.
    {
        const int nlive = sh.nlive;
        for (int base = 0; base < nlive; base += K1_THREADS) {
            const int idx = base + tid;
            const bool active = idx < nlive;
            const int e = active ? sh.live[idx] : 0;
            const int r = e / (2 * NNODE), t = e - r * 2 * NNODE;
            const int k = t >> 1, half = t & 1;
            const double* rowtab = tab0 + (size_t)r * PRE_ROW;
            const double X1 = rowtab[PRE_U1], X2 = rowtab[PRE_U2];
            const double Y1 = sh.Y[r][k][0], Y2 = sh.Y[r][k][1];
            const double X = half ? X2 : X1, Y = half ? Y2 : Y1, Xo = half ? X1 : X2, Yo = half ? Y1 : Y2;
            const double* preS = sh.pre[r][half];
            const double* preO = sh.pre[r][half ^ 1];
            double tv[TV_N];
            if (active) node_values(X, Y, Xo, tv);
            __syncthreads();
            K1_CLK(2);
#define TKB_K1_GROUP(MASK)                                                                   \
            __syncthreads();                                                                  \
            if (active) {                                                                     \
                double hv[NKERN];                                                             \
                eval_half<(MASK), true, 4>(X, Y, Xo, Yo, hv, preS, preO, 0, tv);              \
                _Pragma("unroll")                                                             \
                for (int f = 0; f < NKERN; f++)                                               \
                    if ((MASK) & bit(f)) sh.H[r][half][f][k] = hv[f];                         \
            }
            TKB_K1_GROUP(0x7fu | bit(K_NSN))                   // the seven LO kernels and the nfl part of the non-singlet one
            TKB_K1_GROUP(bit(K_NSP) | bit(K_NSM))
            TKB_K1_GROUP(bit(K_V1QQ) | bit(K_A1QQ))
            TKB_K1_GROUP(bit(K_V1QG) | bit(K_A1QG))
            TKB_K1_GROUP(bit(K_V1GQ) | bit(K_V1GQN))
            TKB_K1_GROUP(bit(K_A1GQ) | bit(K_A1GQN))
            TKB_K1_GROUP(bit(K_V1GG) | bit(K_V1GGN))
            TKB_K1_GROUP(bit(K_A1GG) | bit(K_A1GGN))
#undef TKB_K1_GROUP
        }
    }
    __syncthreads();
    K1_CLK(3);
    // ---- phase 1c: K(x_i, y_k) from the two halves, in place ----
    for (int e = tid; e < nrows * NNODE; e += K1_THREADS) {
        const int r = e / NNODE, k = e - r * NNODE;
        bool fin = true;
#pragma unroll
        for (int f = 0; f < NKERN; f++) {
            const double K = combine_halves_rt(f, sh.H[r][0][f][k], sh.H[r][1][f][k], xi);
            sh.H[r][0][f][k] = K;
            fin = fin && (fabs(K) <= huge_());
        }
        if (!(fin && fabs(sh.J[r][k]) <= huge_())) sh.allfinite[r] = 0;
    }
    __syncthreads();

    K1_CLK(4);
    // ---- phase 2: one thread per (row, column j); replay qk21 per interval in the reference's order ----
    const size_t fstride = (size_t)nxi * n * n;
    // The six columns whose basis function is non-zero at x_i carry the subtraction term f_j(x_i) of the plus
    // kernels at every node (66 accumulation steps instead of the ~4 of an ordinary column).  They are split by
    // kernel pair (a plus-type and a regular kernel where possible: two independent accumulation chains of similar
    // length) AND by integration interval: qk21's sum over the 21 nodes of an interval does not depend on the other
    // intervals, so item (row, column, pair, interval) forms resk of its interval, parks it in the dead half of the value
    // array, and a short second pass adds the six resk * hlgth in the reference's order (same operations, same order:
    // bit-identical).  These 11-step items and the ordinary columns form ONE work list, ordinary columns (the heavier
    // items) first, handed out 32 items at a time to whichever warp is free: as one 66-step item per thread the
    // replay left a third of the threads waiting for the slowest (17 % of the kernel's warp time at the barrier).
    constexpr int NPAIR = NKERN / 2;
    // the nine plus-type kernels 0, 5, 7, 8, 9, 18, 19, 20, 21 each with a regular one, then two regular pairs
    constexpr int PA[NPAIR] = { K_K0QQ, K_A0GG, K_NSP, K_NSM, K_NSN, K_V1GG, K_A1GG, K_V1GGN, K_A1GGN, K_V1GQ, K_V1GQN };
    constexpr int PB[NPAIR] = { K_A0QG, K_VMA0QG, K_A0GQ, K_VMA0GQ, K_VMA0GG, K_V1QQ, K_A1QQ, K_V1QG, K_A1QG, K_A1GQ, K_A1GQN };
    const int nord = nrows * n, nspec = nrows * 6 * NPAIR;
    for (;;) {
    int base2 = 0;
    if ((tid & 31) == 0) base2 = atomicAdd(&sh.next2, 32);
    base2 = __shfl_sync(0xffffffffu, base2, 0);
    if (base2 >= nord + 6 * nspec) break;
    const int e = base2 + (tid & 31);
    if (e >= nord + 6 * nspec) continue;
    if (e >= nord) {
        const int es = e - nord;
        const int iv = es % 6, ep = es / 6;
        const int r = ep / (6 * NPAIR), t = ep - r * 6 * NPAIR;
        if (sh.allfinite[r] == 0) continue;            // a NaN/Inf kernel value poisons every column (0*NaN): dense pass below
        const int jh = t / NPAIR, pr = t - NPAIR * jh;
        int fa = PA[0], fb = PB[0];
#pragma unroll
        for (int q = 1; q < NPAIR; q++)
            if (q == pr) { fa = PA[q]; fb = PB[q]; }
        const int j = sh.iSx[r] + jh;
        const double fx = sh.bx[r][jh];
        if (fx == 0.0) continue;
        const double (*sK)[NNODE] = sh.H[r][0];
        const bool plusa = is_plus(fa), plusb = is_plus(fb);
        double reska = 0.0, reskb = 0.0;
        for (int s = 0; s <= 10; s++) {
            const int tt = (s == 0) ? 11 : ((s <= 5) ? 2 * s : 2 * (s - 5) - 1);
            const int k1 = 21 * iv + ((s == 0) ? 0 : 2 * tt - 1);
            const int k2 = k1 + 1;
            const int m1 = j - sh.iS[r][k1];
            const bool in1 = (m1 >= 0 && m1 < 6);
            bool in2 = false;
            int m2 = 0;
            if (s != 0) { m2 = j - sh.iS[r][k2]; in2 = (m2 >= 0 && m2 < 6); }
            const bool doa = plusa || in1 || in2, dob = plusb || in1 || in2;
            if (!(doa || dob)) continue;
            const double fy1 = in1 ? sh.b[r][k1][m1] : 0.0;
            const double J1 = sh.J[r][k1];
            const double w = c_wgk[tt];
            if (s == 0) {
                if (doa) { const double fc = plusa ? (sK[fa][k1] * (fy1 - fx)) * J1 : (sK[fa][k1] * fy1) * J1; reska = w * fc; }
                if (dob) { const double fc = plusb ? (sK[fb][k1] * (fy1 - fx)) * J1 : (sK[fb][k1] * fy1) * J1; reskb = w * fc; }
            } else {
                const double fy2 = in2 ? sh.b[r][k2][m2] : 0.0;
                const double J2 = sh.J[r][k2];
                if (doa) {
                    const double v1 = plusa ? (sK[fa][k1] * (fy1 - fx)) * J1 : (sK[fa][k1] * fy1) * J1;
                    const double v2 = plusa ? (sK[fa][k2] * (fy2 - fx)) * J2 : (sK[fa][k2] * fy2) * J2;
                    const double fsum = v1 + v2;
                    reska = reska + w * fsum;
                }
                if (dob) {
                    const double v1 = plusb ? (sK[fb][k1] * (fy1 - fx)) * J1 : (sK[fb][k1] * fy1) * J1;
                    const double v2 = plusb ? (sK[fb][k2] * (fy2 - fx)) * J2 : (sK[fb][k2] * fy2) * J2;
                    const double fsum = v1 + v2;
                    reskb = reskb + w * fsum;
                }
            }
        }
        double* park = &sh.H[r][1][0][0] + (size_t)(t * 2) * 6;      // [column, pair][a, b][interval], NKERN * NNODE doubles per row are free
        park[iv] = reska;
        park[6 + iv] = reskb;
    } else {
        const int e2 = e;
        const int r = e2 / n, j = e2 - r * n + 1;
        const bool dense = (sh.allfinite[r] == 0);
        const int iSx = sh.iSx[r];
        const int mx = j - iSx;
        const double fx = (mx >= 0 && mx < 6) ? sh.bx[r][mx] : 0.0;
        if (!dense && fx != 0.0) continue;                     // done above, one thread per kernel
        const double (*sK)[NNODE] = sh.H[r][0];
        double integral[NKERN];
#pragma unroll
        for (int f = 0; f < NKERN; f++) integral[f] = 0.0;
        for (int iv = 0; iv < 6; iv++) {
            double resk[NKERN];
#pragma unroll
            for (int f = 0; f < NKERN; f++) resk[f] = 0.0;
            // an ordinary column (f_j(x_i) = 0) outside the interval's stencils gets exact zeros from all 21 nodes
            const bool touched = dense || fx != 0.0 || (j >= sh.jlo[r][iv] && j <= sh.jhi[r][iv] + 5);
            // qk21 order: centre (weight wgk(11)), then t = 2,4,..,10, then t = 1,3,..,9
            for (int s = 0; touched && s <= 10; s++) {
                const int tt = (s == 0) ? 11 : ((s <= 5) ? 2 * s : 2 * (s - 5) - 1);
                const int k1 = 21 * iv + ((s == 0) ? 0 : 2 * tt - 1);
                const int k2 = k1 + 1;                       // unused for the centre
                const int m1 = j - sh.iS[r][k1];
                const bool in1 = (m1 >= 0 && m1 < 6);
                bool in2 = false;
                int m2 = 0;
                if (s != 0) { m2 = j - sh.iS[r][k2]; in2 = (m2 >= 0 && m2 < 6); }
                if (!(dense || in1 || in2 || fx != 0.0)) continue;
                const double fy1 = in1 ? sh.b[r][k1][m1] : 0.0;
                const double J1 = sh.J[r][k1];
                const double w = c_wgk[tt];
                if (s == 0) {
#pragma unroll
                    for (int f = 0; f < NKERN; f++) {
                        const double fc = is_plus(f) ? (sK[f][k1] * (fy1 - fx)) * J1 : (sK[f][k1] * fy1) * J1;
                        resk[f] = w * fc;
                    }
                } else {
                    const double fy2 = in2 ? sh.b[r][k2][m2] : 0.0;
                    const double J2 = sh.J[r][k2];
#pragma unroll
                    for (int f = 0; f < NKERN; f++) {
                        const double v1 = is_plus(f) ? (sK[f][k1] * (fy1 - fx)) * J1 : (sK[f][k1] * fy1) * J1;
                        const double v2 = is_plus(f) ? (sK[f][k2] * (fy2 - fx)) * J2 : (sK[f][k2] * fy2) * J2;
                        const double fsum = v1 + v2;
                        resk[f] = resk[f] + w * fsum;
                    }
                }
            }
            const double hl = sh.hl[r][iv];
#pragma unroll
            for (int f = 0; f < NKERN; f++) integral[f] = integral[f] + resk[f] * hl;
        }
        const size_t rowbase = ((size_t)ixi * n + (size_t)(i0 + r)) * n;
#pragma unroll
        for (int f = 0; f < NKERN; f++) P[(size_t)f * fstride + rowbase + (size_t)(j - 1)] = integral[f];
    }
    }
    __syncthreads();
    // second pass of the six heavy columns: integral = sum over the intervals of resk * hlgth, in the reference's order
    for (int ep = tid; ep < nspec; ep += K1_THREADS) {
        const int r = ep / (6 * NPAIR), t = ep - r * 6 * NPAIR;
        if (sh.allfinite[r] == 0) continue;
        const int jh = t / NPAIR, pr = t - NPAIR * jh;
        if (sh.bx[r][jh] == 0.0) continue;
        int fa = PA[0], fb = PB[0];
#pragma unroll
        for (int q = 1; q < NPAIR; q++)
            if (q == pr) { fa = PA[q]; fb = PB[q]; }
        const double* park = &sh.H[r][1][0][0] + (size_t)(t * 2) * 6;
        double inta = 0.0, intb = 0.0;
#pragma unroll
        for (int iv = 0; iv < 6; iv++) {
            inta = inta + park[iv] * sh.hl[r][iv];
            intb = intb + park[6 + iv] * sh.hl[r][iv];
        }
        const size_t base = ((size_t)ixi * n + (size_t)(i0 + r)) * n + (size_t)(sh.iSx[r] + jh - 1);
        P[(size_t)fa * fstride + base] = inta;
        P[(size_t)fb * fstride + base] = intb;
    }

    __syncthreads();
    K1_CLK(5);
    // ---- diagonal bookkeeping: f_i(x_i) and the analytic LO constants ----
    if ((tid & 127) == 0 && (tid >> 7) < nrows) {
        const int r = tid >> 7, i = i0 + r + 1;
        const double x = tab0[(size_t)r * PRE_ROW + PRE_X];
        double* d = diag + ((size_t)ixi * n + (size_t)(i - 1)) * DIAG_N;
        const int mi = i - sh.iSx[r];
        d[DIAG_FX] = (mi >= 0 && mi < 6) ? sh.bx[r][mi] : 0.0;
        d[DIAG_X] = x;
        d[DIAG_K0QQ] = K0_qq_cst(x, xi);
        d[DIAG_A0GG] = KA0_GG_cst(x, xi);
        d[DIAG_A0GGN] = KA0_GG_nfl_cst();
    }
}

// ---------------------------------------------------------------------------------------
// K2: QAGS (integration.f90:246-904).  This file holds QUADPACK's scalar routines (qsort, qextr, the qk21 sums) as
// device code; the per-instance state and the bisection step are in qags_step.cuh, the kernels in
// assemble_cst_pipe.cuh (default) and assemble_cst_queue.cuh.
// Interval lists: the first entries of alist/blist/rlist/elist/iord live in shared memory (typical instances need
// 25-60 intervals), entries beyond spill to a per-slot global scratch (the reference's limit is 500, reached by the
// round-off dominated V-type gluon constants at xi ~ 2e-5).
// Evaluation: a bisection step needs the integrand at 2 x 21 nodes = 168 half-kernel tasks (node, direction
// K(x,y) / K(y,x), half).  A half kernel whose support steps rho_step and erbl_step both vanish is exactly zero
// (every piece is multiplied by one of them), so the tasks are first classified, the live ones compacted, and only
// those are evaluated.  A task is only skipped when every denominator and log argument of any half kernel is
// non-zero (`safe`), so a 0*Inf = NaN of the reference is still reproduced.

constexpr int QLIMIT = 500;

#ifdef TKB_K2_STATS
#endif

template <typename T, int CAP>
struct SpillList {                       // 1-based list, head (indices < CAP) in shared memory, tail in global memory
    T* s;
    T* g;
    __device__ __forceinline__ T& operator[](int i) const { return (i < CAP) ? s[i] : g[i]; }
    static constexpr int cap = CAP;
    // entries lo .. hi move one place up (to lo+1 .. hi+1) / one place down (to lo-1 .. hi-1); the shared and the
    // global part are walked by separate loops without the per-access select
    __device__ __forceinline__ void shift_up(int lo, int hi) const
    {
        int t = hi;
        for (; t >= lo && t >= CAP; t--) g[t + 1] = g[t];                 // source and destination in global memory
        if (t >= lo && t == CAP - 1) { g[CAP] = s[CAP - 1]; t--; }         // the entry that crosses the boundary
        for (; t >= lo; t--) s[t + 1] = s[t];
    }
    __device__ __forceinline__ void shift_down(int lo, int hi) const
    {
        int t = lo;
        for (; t <= hi && t < CAP; t++) s[t - 1] = s[t];
        if (t <= hi && t == CAP) { s[CAP - 1] = g[CAP]; t++; }
        for (; t <= hi; t++) g[t - 1] = g[t];
    }
};

// integration.f90:616-729
template <typename DL, typename IL>
static __device__ __noinline__ void qsort_(int limit, int last, int& maxerr, double& ermax, DL elist, IL iord, int& nrmax)
{
    int i, ibeg, isucc, j, jbnd, jupbn, k;
    double errmax, errmin;
    if (last <= 2) {
        iord[1] = 1;
        iord[2] = 2;
        goto L90;
    }
    errmax = elist[maxerr];
    {
        const int ntrip = nrmax - 1;
        for (i = 1; i <= ntrip; i++) {
            isucc = iord[nrmax - 1];
            if (errmax <= elist[isucc]) break;
            iord[nrmax] = isucc;
            nrmax = nrmax - 1;
        }
    }
    jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    errmin = elist[last];
    jbnd = jupbn - 1;
    ibeg = nrmax + 1;
    for (i = ibeg; i <= jbnd; i++) {
        isucc = iord[i];
        if (errmax >= elist[isucc]) goto L60;
        iord[i - 1] = isucc;
    }
    iord[jbnd] = maxerr;
    iord[jupbn] = last;
    goto L90;
L60:
    iord[i - 1] = maxerr;
    k = jbnd;
    for (j = i; j <= jbnd; j++) {
        isucc = iord[k];
        if (errmin < elist[isucc]) goto L80;
        iord[k + 1] = isucc;
        k = k - 1;
    }
    iord[i] = last;
    goto L90;
L80:
    iord[k + 1] = last;
L90:
    maxerr = iord[nrmax];
    ermax = elist[maxerr];
}

// The same ordering update with the two linear scans of qsort_ replaced by galloping + binary searches and
// block shifts.  The bottom scan of qsort_ walks, on average, a third of the list with two dependent loads
// and an FP64 compare per entry; on the long bisection paths of the plus-prescription constants (hundreds of
// intervals) that made the ordering update the largest part of the bookkeeping time.  Both scans test a
// predicate that is monotone along a list sorted by decreasing error estimate, which qsort_ maintains as
// long as no estimate is a NaN (the caller falls back to qsort_ for good once it has seen one); the first
// index where the predicate flips is then the same whichever way it is searched, ties included, so the
// resulting list, maxerr and nrmax are those of qsort_.
template <typename DL, typename IL>
static __device__ __noinline__ void qsort_bs_(int limit, int last, int& maxerr, double& ermax, DL elist, IL iord, int& nrmax)
{
    if (last <= 2) {
        iord[1] = 1;
        iord[2] = 2;
        maxerr = iord[nrmax];
        ermax = elist[maxerr];
        return;
    }
    const double errmax = elist[maxerr];
    {
        const int ntrip = nrmax - 1;
        for (int i = 1; i <= ntrip; i++) {
            const int isucc = iord[nrmax - 1];
            if (errmax <= elist[isucc]) break;
            iord[nrmax] = isucc;
            nrmax = nrmax - 1;
        }
    }
    int jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    const double errmin = elist[last];
    const int jbnd = jupbn - 1, ibeg = nrmax + 1;
    // i = first index in [ibeg, jbnd] with errmax >= elist[iord[i]], jbnd + 1 if there is none
    int lo = ibeg, hi = jbnd + 1;
    for (int p = ibeg, step = 1; p <= jbnd; p += step, step <<= 1) {
        if (errmax >= elist[iord[p]]) { hi = p; break; }
        lo = p + 1;
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (errmax >= elist[iord[mid]]) hi = mid; else lo = mid + 1;
    }
    const int i = lo;
    iord.shift_down(ibeg, i - 1);
    if (i > jbnd) {
        iord[jbnd] = maxerr;
        iord[jupbn] = last;
    } else {
        iord[i - 1] = maxerr;
        // k = largest index in [i, jbnd] with errmin < elist[iord[k]], i - 1 if there is none
        int klo = i - 1, khi = jbnd + 1;
        for (int p = jbnd, step = 1; p >= i; p -= step, step <<= 1) {
            if (errmin < elist[iord[p]]) { klo = p; break; }
            khi = p;
        }
        while (khi - klo > 1) {
            const int mid = (klo + khi) >> 1;
            if (errmin < elist[iord[mid]]) klo = mid; else khi = mid;
        }
        iord.shift_up(klo + 1, jbnd);
        iord[klo + 1] = last;
    }
    maxerr = iord[nrmax];
    ermax = elist[maxerr];
}

// integration.f90:732-904
static __device__ __noinline__ void qextr_(int& n, double* epstab, double& result, double& abserr, double* res3la, int& nres)
{
    int i, ib, ib2, ie, indx, k1, k2, k3, limexp, newelm, num;
    double delta1, delta2, delta3, epsinf, error, err1, err2, err3, e0, e1, e1abs, e2, e3, res, ss, tol1, tol2, tol3;
    nres = nres + 1;
    abserr = huge_();
    result = epstab[n];
    if (n < 3) goto L100;
    limexp = 50;
    epstab[n + 2] = epstab[n];
    newelm = (n - 1) / 2;
    epstab[n] = huge_();
    num = n;
    k1 = n;
    for (i = 1; i <= newelm; i++) {
        k2 = k1 - 1;
        k3 = k1 - 2;
        res = epstab[k1 + 2];
        e0 = epstab[k3];
        e1 = epstab[k2];
        e2 = res;
        e1abs = fabs(e1);
        delta2 = e2 - e1;
        err2 = fabs(delta2);
        tol2 = fmax2(fabs(e2), e1abs) * eps_();
        delta3 = e1 - e0;
        err3 = fabs(delta3);
        tol3 = fmax2(e1abs, fabs(e0)) * eps_();
        if (err2 <= tol2 && err3 <= tol3) {
            result = res;
            abserr = err2 + err3;
            goto L100;
        }
        e3 = epstab[k1];
        epstab[k1] = e1;
        delta1 = e1 - e3;
        err1 = fabs(delta1);
        tol1 = fmax2(e1abs, fabs(e3)) * eps_();
        if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) goto L20;
        ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
        epsinf = fabs(ss * e1);
        if (epsinf > 1.0e-04) goto L30;
L20:
        n = i + i - 1;
        break;
L30:
        res = e1 + 1.0 / ss;
        epstab[k1] = res;
        k1 = k1 - 2;
        error = err2 + fabs(res - e2) + err3;
        if (error <= abserr) {
            abserr = error;
            result = res;
        }
    }
    if (n == limexp) n = 2 * (limexp / 2) - 1;
    if ((num / 2) * 2 == num) ib = 2; else ib = 1;
    ie = newelm + 1;
    for (i = 1; i <= ie; i++) {
        ib2 = ib + 2;
        epstab[ib] = epstab[ib2];
        ib = ib2;
    }
    if (num != n) {
        indx = num - n + 1;
        for (i = 1; i <= n; i++) {
            epstab[i] = epstab[indx];
            indx = indx + 1;
        }
    }
    if (nres < 4) {
        res3la[nres] = result;
        abserr = huge_();
    } else {
        abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
        res3la[1] = res3la[2];
        res3la[2] = res3la[3];
        res3la[3] = result;
    }
L100:
    abserr = fmax2(abserr, 0.5 * eps_() * fabs(result));
}

// qk21's reductions from the 21 stored integrand values (integration.f90:159-222).
// fv[q]: q = 0 centre, 2t-1 / 2t = value at centr -/+ hlgth*xgk(t).
static __device__ __noinline__ void qk21_reduce(const double* fv, double a, double b,
                                                double& result, double& abserr, double& resabs, double& resasc)
{
    const double hlgth = 5.0e-01 * (b - a);
    const double dhlgth = fabs(hlgth);
    double resg = 0.0;
    const double fc = fv[0];
    double resk = c_wgk[11] * fc;
    resabs = fabs(resk);
    for (int j = 1; j <= 5; j++) {
        const int jtw = 2 * j;
        const double fval1 = fv[2 * jtw - 1], fval2 = fv[2 * jtw];
        const double fsum = fval1 + fval2;
        resg = resg + c_wg[j] * fsum;
        resk = resk + c_wgk[jtw] * fsum;
        resabs = resabs + c_wgk[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (int j = 1; j <= 5; j++) {
        const int jtwm1 = 2 * j - 1;
        const double fval1 = fv[2 * jtwm1 - 1], fval2 = fv[2 * jtwm1];
        const double fsum = fval1 + fval2;
        resk = resk + c_wgk[jtwm1] * fsum;
        resabs = resabs + c_wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    const double reskh = resk * 5.0e-01;
    resasc = c_wgk[11] * fabs(fc - reskh);
    for (int j = 1; j <= 10; j++)
        resasc = resasc + c_wgk[j] * (fabs(fv[2 * j - 1] - reskh) + fabs(fv[2 * j] - reskh));
    result = resk * hlgth;
    resabs = resabs * dhlgth;
    resasc = resasc * dhlgth;
    abserr = fabs((resk - resg) * hlgth);
    if (resasc != 0.0 && abserr != 0.0)
        abserr = resasc * fmin2(1.0, pow15_capped(2.0e+02 * abserr / resasc));
    if (resabs > tiny_() / (5.0e+01 * eps_()))
        abserr = fmax2((eps_() * 5.0e+01) * resabs, abserr);
}

}  // namespace tkb
#include "qags_step.cuh"
#include "assemble_cst_pipe.cuh"
#include "assemble_cst_pair.cuh"
#include "assemble_cst_queue.cuh"
namespace tkb {

// ---------------------------------------------------------------------------------------
// K3: composition of the nfl slice (matevo.f90:293, 331-335, 370-376, 389-396, 437-446, 463-472).
// Output matrices are row-major per xi: NS (n x n): NS0, NS1p, NS1m; SG (2n x 2n): V0, A0, V1, A1.

__global__ void compose_kernel(int n, int nxi, int nfl, const double* __restrict__ P, const double* __restrict__ diag,
                               const double* __restrict__ cst_res,
                               double* __restrict__ NS0, double* __restrict__ NS1p, double* __restrict__ NS1m,
                               double* __restrict__ SGV0, double* __restrict__ SGA0,
                               double* __restrict__ SGV1, double* __restrict__ SGA1)
{
    const size_t nn = (size_t)n * n;
    const size_t total = nn * nxi;
    const size_t fstride = total;
    const double fn = (double)nfl;
    const size_t m = (size_t)2 * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t ixi = e / nn;
        const size_t rem = e - ixi * nn;
        const size_t i = rem / n, j = rem - i * n;
        // pixel_conv = ((0 + reg) + pls) + cst; cst = Kcst(x,xi) * f_i(x_i) on the diagonal only
        double c_k0qq = 0.0, c_nsp = 0.0, c_nsm = 0.0, c_nsn = 0.0, c_a0gg = 0.0, c_a0ggn = 0.0,
               c_vgg = 0.0, c_agg = 0.0, c_vggn = 0.0, c_aggn = 0.0;
        if (i == j) {
            const double* d = diag + (ixi * n + i) * DIAG_N;
            const double fx = d[DIAG_FX];
            const double* q = cst_res + (ixi * n + i) * 21;
            double integ[7];
#pragma unroll
            for (int c = 0; c < 7; c++) {
                double s = 0.0;
                s = s + q[c * 3 + 0];
                s = s + q[c * 3 + 1];
                s = s + q[c * 3 + 2];
                integ[c] = s;
            }
            c_k0qq = d[DIAG_K0QQ] * fx;
            c_a0gg = d[DIAG_A0GG] * fx;
            c_a0ggn = d[DIAG_A0GGN] * fx;
            c_nsp = (nlo_explicit_cst(0) + integ[0]) * fx;
            c_nsm = integ[1] * fx;
            c_nsn = integ[2] * fx;
            c_vgg = integ[3] * fx;
            c_agg = integ[4] * fx;                                  // explicit term overwritten in the reference
            c_vggn = (nlo_explicit_cst(5) + integ[5]) * fx;
            c_aggn = (nlo_explicit_cst(6) + integ[6]) * fx;
        }
#define PV(k) P[(size_t)(k) * fstride + e]
        const double ns0 = PV(K_K0QQ) + c_k0qq;
        const double nsn = PV(K_NSN) + c_nsn;
        const double ns1p = (PV(K_NSP) + c_nsp) + nsn * fn;
        const double ns1m = (PV(K_NSM) + c_nsm) + nsn * fn;
        NS0[e] = ns0;
        NS1p[e] = ns1p;
        NS1m[e] = ns1m;
        const size_t base = ixi * m * m;
        const size_t qq = base + i * m + j, qg = base + i * m + (j + n);
        const size_t gq = base + (i + n) * m + j, gg = base + (i + n) * m + (j + n);
        const double a_qg = fn * PV(K_A0QG);
        const double a_gq = PV(K_A0GQ);
        const double a_gg = (PV(K_A0GG) + c_a0gg) + fn * (0.0 + c_a0ggn);
        SGA0[qq] = ns0; SGA0[qg] = a_qg; SGA0[gq] = a_gq; SGA0[gg] = a_gg;
        SGV0[qq] = 0.0 + ns0;
        SGV0[qg] = fn * PV(K_VMA0QG) + a_qg;
        SGV0[gq] = PV(K_VMA0GQ) + a_gq;
        SGV0[gg] = PV(K_VMA0GG) + a_gg;
        SGV1[qq] = fn * PV(K_V1QQ) + ns1p;
        SGV1[qg] = fn * PV(K_V1QG);
        SGV1[gq] = PV(K_V1GQ) + fn * PV(K_V1GQN);
        SGV1[gg] = (PV(K_V1GG) + c_vgg) + fn * (PV(K_V1GGN) + c_vggn);
        SGA1[qq] = fn * PV(K_A1QQ) + ns1m;
        SGA1[qg] = fn * PV(K_A1QG);
        SGA1[gq] = PV(K_A1GQ) + fn * PV(K_A1GQN);
        SGA1[gg] = (PV(K_A1GG) + c_agg) + fn * (PV(K_A1GGN) + c_aggn);
#undef PV
    }
}

// ---------------------------------------------------------------------------------------
// test hooks: single kernel / dilog / log / pow evaluations (used by the parity tests through the C-ABI)

__global__ void eval_kernels_kernel(int npts, const double* __restrict__ xs, const double* __restrict__ ys,
                                    const double* __restrict__ xis, double* __restrict__ out /*[npts][NKERN]*/)
{
    log_tab_init();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const double x = xs[p], y = ys[p], xi = xis[p];
    const double X1 = 0.5 * (1.0 + x / xi), X2 = 0.5 * (1.0 - x / xi);
    const double Y1 = 0.5 * (1.0 + y / xi), Y2 = 0.5 * (1.0 - y / xi);
    double h1[NKERN], h2[NKERN];
    eval_half<MASK_ALL>(X1, Y1, X2, Y2, h1);
    eval_half<MASK_ALL>(X2, Y2, X1, Y1, h2);
    for (int f = 0; f < NKERN; f++) out[(size_t)p * NKERN + f] = combine_halves_rt(f, h1[f], h2[f], xi);
}

__global__ void eval_scalar_kernel(int which, int npts, const double* __restrict__ a, const double* __restrict__ b,
                                   double* __restrict__ out)
{
    log_tab_init();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    switch (which) {
    case 0: out[p] = log_(a[p]); break;
    case 1: out[p] = pow_(a[p], b[p]); break;
    case 2: out[p] = dilog(a[p]); break;
    case 3: out[p] = log_cr(a[p]); break;
    }
}

// calibration / test hook: the QAGS integrand K(x,y,xi) - K(y,x,xi) of constant KID (kernels_nlo.f90:71-77 and its
// siblings) evaluated straight from the reference formulas -- all four half kernels in full, no support-class
// skipping, no per-row tables.  Its FP64 instruction count per point is the ALGORITHMIC work of one integrand
// evaluation that the roofline of K2 is quoted in (profiles/k2_flops_per_eval.json); the tests compare its values
// with the oracle's kernel functions.
template <int KID>
__global__ void eval_integrand_kernel(int npts, const double* __restrict__ xs, const double* __restrict__ ys,
                                      const double* __restrict__ xis, double* __restrict__ out)
{
    log_tab_init();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const double x = xs[p], y = ys[p], xi = xis[p];
    const double xx1 = 0.5 * (1.0 + x / xi), xx2 = 0.5 * (1.0 - x / xi);
    const double yy1 = 0.5 * (1.0 + y / xi), yy2 = 0.5 * (1.0 - y / xi);
    double h[4];
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int dir = k >> 1, half = k & 1;
        const double U1 = dir ? yy1 : xx1, U2 = dir ? yy2 : xx2, V1 = dir ? xx1 : yy1, V2 = dir ? xx2 : yy2;
        double o[NKERN];
        eval_half<bit(KID), true>(half ? U2 : U1, half ? V2 : V1, half ? U1 : U2, half ? V1 : V2, o);
        h[k] = o[KID];
    }
    out[p] = combine_halves<KID>(h[0], h[1], xi) - combine_halves<KID>(h[2], h[3], xi);
}

int launch_eval_integrand(int c, int npts, const double* xs, const double* ys, const double* xis, double* out, cudaStream_t st)
{
    const int g = (npts + 127) / 128;
    switch (c) {
    case 0: eval_integrand_kernel<K_NSP><<<g, 128, 0, st>>>(npts, xs, ys, xis, out); break;
    case 1: eval_integrand_kernel<K_NSM><<<g, 128, 0, st>>>(npts, xs, ys, xis, out); break;
    case 2: eval_integrand_kernel<K_NSN><<<g, 128, 0, st>>>(npts, xs, ys, xis, out); break;
    case 3: eval_integrand_kernel<K_V1GG><<<g, 128, 0, st>>>(npts, xs, ys, xis, out); break;
    case 4: eval_integrand_kernel<K_A1GG><<<g, 128, 0, st>>>(npts, xs, ys, xis, out); break;
    case 5: eval_integrand_kernel<K_V1GGN><<<g, 128, 0, st>>>(npts, xs, ys, xis, out); break;
    case 6: eval_integrand_kernel<K_A1GGN><<<g, 128, 0, st>>>(npts, xs, ys, xis, out); break;
    default: return set_error("tkb_eval_integrand: constant index out of range");
    }
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

// test hook for the task skipping of K1 / K2: per point and for the four half-kernel tasks of a node
// (K(x,y) halves 0/1, K(y,x) halves 0/1): the support class the assembly kernels would assign, the value of the
// full evaluation and the value of the erbl-only evaluation of one QAGS kernel.  The tests require: class 0 =>
// the full value is an exact zero; class 1 => the erbl-only value equals the full value bit for bit.
template <int KID>
__global__ void task_classes_kernel(int npts, const double* __restrict__ xs, const double* __restrict__ ys,
                                    const double* __restrict__ xis, int* __restrict__ cls /*[npts][4]*/,
                                    double* __restrict__ full /*[npts][4]*/, double* __restrict__ erbl /*[npts][4]*/)
{
    log_tab_init();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const double x = xs[p], y = ys[p], xi = xis[p];
    const double xx1 = 0.5 * (1.0 + x / xi), xx2 = 0.5 * (1.0 - x / xi);
    const double yy1 = 0.5 * (1.0 + y / xi), yy2 = 0.5 * (1.0 - y / xi);
    const bool safe = node_is_safe(xx1, xx2, yy1, yy2);
    for (int k = 0; k < 4; k++) {
        const int dir = k >> 1, half = k & 1;
        const double U1 = dir ? yy1 : xx1, U2 = dir ? yy2 : xx2, V1 = dir ? xx1 : yy1, V2 = dir ? xx2 : yy2;
        const double X = half ? U2 : U1, Y = half ? V2 : V1, Xo = half ? U1 : U2, Yo = half ? V1 : V2;
        double out[NKERN];
        cls[4 * p + k] = task_class(X, Y, safe);
        eval_half<bit(KID), true>(X, Y, Xo, Yo, out);
        full[4 * p + k] = out[KID];
        eval_half<bit(KID), false>(X, Y, Xo, Yo, out);
        erbl[4 * p + k] = out[KID];
    }
}

// the same for K1, which skips class-0 halves for all 22 kernels at once: full[.] = number of the 22 half-kernel
// values that are not an exact zero (NaN counts), erbl[.] = 0
__global__ void task_classes_all_kernel(int npts, const double* __restrict__ xs, const double* __restrict__ ys,
                                        const double* __restrict__ xis, int* __restrict__ cls,
                                        double* __restrict__ full, double* __restrict__ erbl)
{
    log_tab_init();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const double x = xs[p], y = ys[p], xi = xis[p];
    const double xx1 = 0.5 * (1.0 + x / xi), xx2 = 0.5 * (1.0 - x / xi);
    const double yy1 = 0.5 * (1.0 + y / xi), yy2 = 0.5 * (1.0 - y / xi);
    const bool safe = node_is_safe(xx1, xx2, yy1, yy2);
    for (int k = 0; k < 4; k++) {
        const int dir = k >> 1, half = k & 1;
        const double U1 = dir ? yy1 : xx1, U2 = dir ? yy2 : xx2, V1 = dir ? xx1 : yy1, V2 = dir ? xx2 : yy2;
        const double X = half ? U2 : U1, Y = half ? V2 : V1, Xo = half ? U1 : U2, Yo = half ? V1 : V2;
        double out[NKERN];
        cls[4 * p + k] = task_class(X, Y, safe);
        eval_half<MASK_ALL, true>(X, Y, Xo, Yo, out);
        int nz = 0;
        for (int f = 0; f < NKERN; f++) nz += !(out[f] == 0.0);
        full[4 * p + k] = (double)nz;
        erbl[4 * p + k] = 0.0;
    }
}

int launch_task_classes(int c, int npts, const double* xs, const double* ys, const double* xis, int* cls, double* full,
                        double* erbl, cudaStream_t st)
{
    const int g = (npts + 63) / 64;
    switch (c) {
    case 0: task_classes_kernel<K_NSP><<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    case 1: task_classes_kernel<K_NSM><<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    case 2: task_classes_kernel<K_NSN><<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    case 3: task_classes_kernel<K_V1GG><<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    case 4: task_classes_kernel<K_A1GG><<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    case 5: task_classes_kernel<K_V1GGN><<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    case 6: task_classes_kernel<K_A1GGN><<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    case 7: task_classes_all_kernel<<<g, 64, 0, st>>>(npts, xs, ys, xis, cls, full, erbl); break;
    default: return set_error("tkb_task_classes: constant index out of range");
    }
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------
// host launchers

#define TKB_LAUNCH_CHECK()                                                   \
    do {                                                                     \
        cudaError_t e_ = cudaGetLastError();                                 \
        ++g_launches;                                                        \
        if (e_ != cudaSuccess) return set_error(cudaGetErrorString(e_));     \
    } while (0)

size_t pre_table_doubles(size_t rows) { return rows * PRE_ROW; }
int launch_pre_table(int n, int nxi, const double* xi_dev, int grid_type, double* tab, cudaStream_t st)
{
    cudaError_t e = cudaMemcpyToSymbolAsync(d_pre_tab, &tab, sizeof(tab), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    const int total = n * nxi * PRE_ROW;
    pre_table_kernel<<<(total + 127) / 128, 128, 0, st>>>(n, nxi, xi_dev, grid_type, tab);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

int set_grid_variant(int eta_recip_f32, int x0_recip_f32)
{
    const int v[2] = { eta_recip_f32 != 0, x0_recip_f32 != 0 };
    const cudaError_t e = cudaMemcpyToSymbol(c_variant, v, sizeof v);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

int launch_assemble_fixed(int n, int nxi, const double* xi_dev, int grid_type, double* P, double* diag, cudaStream_t st)
{
    dim3 grid((n + K1_ROWS - 1) / K1_ROWS, nxi);
    const size_t smem = sizeof(K1Shared);
    cudaError_t e = cudaFuncSetAttribute(assemble_fixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    assemble_fixed_kernel<<<grid, K1_THREADS, smem, st>>>(n, nxi, xi_dev, grid_type, P, diag);
    TKB_LAUNCH_CHECK();
#ifdef TKB_K2_STATS
    {
        cudaStreamSynchronize(st);
        unsigned long long h[24];
        cudaMemcpyFromSymbol(h, g_k2stats, sizeof h);
        const double nb = (double)grid.x * grid.y;
        fprintf(stderr, "K1 phases, cycles per block (%d rows): set-up %.0f, nodes + classes %.0f, node transcendentals %.0f, 22 kernels in 8 passes %.0f, combine halves %.0f, column replay %.0f\n",
                K1_ROWS, h[0] / nb, h[1] / nb, h[2] / nb, h[3] / nb, h[4] / nb, h[5] / nb);
        unsigned long long z[24] = { 0 };
        cudaMemcpyToSymbol(g_k2stats, z, sizeof z);
    }
#endif
    return 0;
}

// ---- K2 launchers.  One launch per constant on its own stream, so that a draining launch does not hold up the next.
#ifndef TKB_QQG
#define TKB_QQG 16
#endif
#ifndef TKB_QQN
#define TKB_QQN 2
#endif
#ifndef TKB_QQT
#define TKB_QQT 768
#endif
#ifndef TKB_QQG_L
#define TKB_QQG_L TKB_QQG
#endif
#ifndef TKB_QQN_L
#define TKB_QQN_L TKB_QQN
#endif
#ifndef TKB_QPG
#define TKB_QPG 16
#endif
#ifndef TKB_QPT
#define TKB_QPT 768
#endif
#ifndef TKB_QPBPS
#define TKB_QPBPS 1
#endif
#ifndef TKB_QPP
#define TKB_QPP 8                 // pair kernel: pair slots per group
#endif
#ifndef TKB_QPBWG
#define TKB_QPBWG 1
#endif
#ifndef TKB_QPBWG_NFL
#define TKB_QPBWG_NFL TKB_QPBWG
#endif
template <int KID>
static int launch_pipe_one(int c, int n, int nxi, const double* xi_dev, int grid_type, QLists* lists, int sms,
                           unsigned long long* counter, double* cst_res, int* cst_last, int* cst_ifail, cudaStream_t st)
{
    constexpr int QG = TKB_QPG, NT = TKB_QPT, BPS = TKB_QPBPS;
    // bookkeeping warps per group: the cheap nfl kernels are bookkeeping-bound
    constexpr int BWG = (KID == K_NSN || KID == K_V1GGN || KID == K_A1GGN) ? TKB_QPBWG_NFL : TKB_QPBWG;
    const size_t smem = sizeof(QPipeShared<QG>);
    auto kern = assemble_cst_pipe_kernel<KID, QG, NT, BPS, BWG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    long long nb = (long long)sms * BPS;
    const long long want = (long long)((total + 2 * QG - 1) / (2 * QG));
    if (nb > want) nb = want;
    kern<<<(int)nb, NT, smem, st>>>(c, n, nxi, xi_dev, grid_type, lists, counter, cst_res, cst_last, cst_ifail);
    TKB_LAUNCH_CHECK();
#ifdef TKB_K2_STATS
    {
        cudaStreamSynchronize(st);
        unsigned long long h[24];
        cudaMemcpyFromSymbol(h, g_k2stats, sizeof h);
        fprintf(stderr, "K2 pipe phases KID %d: node pass %.0f, evaluation %.0f, K(x,y)-K(y,x) %.0f cyc per group-round; tasks per group-round %.1f\n",
                KID, h[6] ? (double)h[16] / h[6] : 0., h[6] ? (double)h[17] / h[6] : 0., h[6] ? (double)h[18] / h[6] : 0., h[6] ? (double)h[19] / h[6] : 0.);
        fprintf(stderr, "K2 pipe stats KID %d: bookkeeper per round: wait-values %.0f, bookkeeping to the hand-over %.0f, deferred extrapolation %.0f cyc (%llu rounds) | evaluators per round: wait-requests %.0f, work %.0f cyc (%llu rounds) | block lifetime %.3g cyc, of which draining after the queue ran dry %.3g (%.1f %%)\n",
                KID, h[3] ? (double)h[0] / h[3] : 0., h[3] ? (double)h[1] / h[3] : 0., h[3] ? (double)h[2] / h[3] : 0., h[3],
                h[6] ? (double)h[4] / h[6] : 0., h[6] ? (double)h[5] / h[6] : 0., h[6], (double)h[7], (double)h[15],
                h[7] ? 100.0 * (double)h[15] / (double)h[7] : 0.0);
        unsigned long long z[24] = { 0 };
        cudaMemcpyToSymbol(g_k2stats, z, sizeof z);
    }
#endif
    return 0;
}

// KV1_NSp and KV1_NSm in one launch (assemble_cst_pair.cuh)
template <int KA, int KB>
static int launch_pair_one(int ca, int cb, int n, int nxi, const double* xi_dev, int grid_type, QLists* lists, int sms,
                           unsigned long long* counter, double* cst_res, int* cst_last, int* cst_ifail, cudaStream_t st)
{
    constexpr int QP = TKB_QPP, NT = TKB_QPT, BPS = TKB_QPBPS;
    const size_t smem = sizeof(QPairShared<QP>);
    auto kern = assemble_cst_pair_kernel<KA, KB, QP, NT, BPS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    long long nb = (long long)sms * BPS;
    const long long want = (long long)((total + 2 * QP - 1) / (2 * QP));
    if (nb > want) nb = want;
    kern<<<(int)nb, NT, smem, st>>>(ca, cb, n, nxi, xi_dev, grid_type, lists, counter, cst_res, cst_last, cst_ifail);
    TKB_LAUNCH_CHECK();
#ifdef TKB_K2_STATS
    {
        cudaStreamSynchronize(st);
        unsigned long long h[24];
        cudaMemcpyFromSymbol(h, g_k2stats, sizeof h);
        const double nw = h[6] ? (double)h[6] : 1., nr = h[5] ? (double)h[5] : 1., nb = h[3] ? (double)h[3] : 1.;
        fprintf(stderr, "K2 pair KID %d+%d: evaluator warp per group-round (mean over warps): wait-requests %.0f, node pass %.0f, barrier %.0f, evaluation %.0f, barrier %.0f, K pass + hand-over %.0f cyc; tasks per group-round %.1f (%.0f group-rounds) | bookkeeper per round: wait-values %.0f, to hand-over %.0f, deferred %.0f cyc (%.0f rounds), active states %.1f of %d | pair slots per round: shared request %.2f, different requests %.2f, one state waiting %.2f\n",
                KA, KB, h[4] / nw, h[16] / nw, h[7] / nw, h[17] / nw, h[15] / nw, h[18] / nw, h[19] / nr, nr,
                h[0] / nb, h[1] / nb, h[2] / nb, nb, h[13] / nb, 2 * QP, h[14] / nr, h[21] / nr, h[22] / nr);
        unsigned long long z[24] = { 0 };
        cudaMemcpyToSymbol(g_k2stats, z, sizeof z);
    }
#endif
    return 0;
}

template <int KID>
static int launch_queue_one(int c, int n, int nxi, const double* xi_dev, int grid_type, QLists* lists, int sms,
                            unsigned long long* counter, double* cst_res, int* cst_last, int* cst_ifail, cudaStream_t st)
{
    constexpr bool LIGHT = (KID == K_NSN || KID == K_V1GGN || KID == K_A1GGN);
    constexpr int QG = LIGHT ? TKB_QQG_L : TKB_QQG, NG = LIGHT ? TKB_QQN_L : TKB_QQN, NT = TKB_QQT;
    const size_t smem = sizeof(QQueueShared<QG, NG>);
    auto kern = assemble_cst_queue_kernel<KID, QG, NG, NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    long long nb = (long long)sms;
    const long long want = (long long)((total + QG * NG - 1) / (QG * NG));
    if (nb > want) nb = want;
    kern<<<(int)nb, NT, smem, st>>>(c, n, nxi, xi_dev, grid_type, lists, counter, cst_res, cst_last, cst_ifail);
    TKB_LAUNCH_CHECK();
#ifdef TKB_K2_STATS
    {
        cudaStreamSynchronize(st);
        unsigned long long h[24];
        cudaMemcpyFromSymbol(h, g_k2stats, sizeof h);
        const double wl = (double)h[7];            // summed worker-warp lifetimes
        fprintf(stderr, "K2 queue KID %d: worker time: NODE %.1f %% EVAL %.1f %% K %.1f %% idle(sleep) %.1f %% other %.1f %% | %llu chunks, %.0f cyc/chunk | bookkeeper per round: wait %.0f, work %.0f cyc (%llu rounds)\n",
                KID, 100. * h[16] / wl, 100. * h[17] / wl, 100. * h[18] / wl, 100. * h[4] / wl,
                100. * (wl - h[16] - h[17] - h[18] - h[4]) / wl, h[19], h[19] ? (double)(h[16] + h[17] + h[18]) / h[19] : 0.,
                h[3] ? (double)h[0] / h[3] : 0., h[3] ? (double)h[1] / h[3] : 0., h[3]);
        fprintf(stderr, "   bookkeeping detail: qk21 sums %.0f cyc per round | qsort %.0f cyc x %llu calls (%.1f per round) | qextr %.0f cyc x %llu calls (%.2f per round)\n",
                h[3] ? (double)h[12] / h[3] : 0., h[9] ? (double)h[8] / h[9] : 0., h[9], h[3] ? (double)h[9] / h[3] : 0.,
                h[11] ? (double)h[10] / h[11] : 0., h[11], h[3] ? (double)h[11] / h[3] : 0.);
        unsigned long long z[24] = { 0 };
        cudaMemcpyToSymbol(g_k2stats, z, sizeof z);
    }
#endif
    return 0;
}

// list slots reserved per SM and constant in the global spill area
constexpr int K2_SLOTS = (2 * TKB_QPG * TKB_QPBPS > TKB_QQG * TKB_QQN)
                             ? ((2 * TKB_QPG * TKB_QPBPS > TKB_QQG_L * TKB_QQN_L) ? 2 * TKB_QPG * TKB_QPBPS : TKB_QQG_L * TKB_QQN_L)
                             : ((TKB_QQG * TKB_QQN > TKB_QQG_L * TKB_QQN_L) ? TKB_QQG * TKB_QQN : TKB_QQG_L * TKB_QQN_L);
size_t qags_lists_bytes(int sms) { return sizeof(QLists) * (size_t)K2_SLOTS * (size_t)sms * 7; }

// All seven constants: instance -> (ixi, i, sub); results go to cst_res[((ixi*n + i)*7 + c)*3 + sub] (same index for
// last / ifail).  Blocks pull instances from one global counter per constant: costs vary by three orders of magnitude
// between instances.  queue = the work-queue form (TKB_K2_MODE=queue), else the role-split pipeline.
int launch_assemble_cst(int n, int nxi, const double* xi_dev, int grid_type, void* lists, int sms, int queue,
                        unsigned long long* counters /*[7]*/, double* cst_res, int* cst_last, int* cst_ifail,
                        cudaStream_t st, cudaStream_t* side /*[7]*/, cudaEvent_t fork, cudaEvent_t* join /*[7]*/)
{
    cudaError_t e = cudaMemsetAsync(counters, 0, 7 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    e = cudaEventRecord(fork, st);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    for (int k = 0; k < 7; k++) {
        e = cudaStreamWaitEvent(side[k], fork, 0);
        if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    }
    QLists* Lp = (QLists*)lists;
    const size_t region = (size_t)K2_SLOTS * sms;
#define K2_ONE(c, KID, strm)                                                                                           \
    if (queue ? launch_queue_one<KID>(c, n, nxi, xi_dev, grid_type, Lp + region * (c), sms, counters + (c), cst_res,   \
                                      cst_last, cst_ifail, strm)                                                       \
              : launch_pipe_one<KID>(c, n, nxi, xi_dev, grid_type, Lp + region * (c), sms, counters + (c), cst_res,    \
                                     cst_last, cst_ifail, strm)) return 1;
    // the expensive, long-tailed constants first.  KV1_NSp and KV1_NSm share one launch (one evaluation gives both
    // integrands; TKB_K2_PAIR=0 runs them separately for comparison)
    static const bool pair_ns = [] { const char* e = getenv("TKB_K2_PAIR"); return !(e && e[0] == '0'); }();
    K2_ONE(3, K_V1GG, side[0])
    if (pair_ns && !queue) {
        if (launch_pair_one<K_NSP, K_NSM>(0, 1, n, nxi, xi_dev, grid_type, Lp + region * 0, sms, counters + 0, cst_res, cst_last,
                                          cst_ifail, side[1])) return 1;
    } else {
        K2_ONE(0, K_NSP, side[1])
    }
    K2_ONE(4, K_A1GG, side[2])
    if (!(pair_ns && !queue)) { K2_ONE(1, K_NSM, side[5]) }
    K2_ONE(5, K_V1GGN, side[3])
    K2_ONE(6, K_A1GGN, side[4])
    K2_ONE(2, K_NSN, side[6])
#undef K2_ONE
    for (int k = 0; k < 7; k++) {
        e = cudaEventRecord(join[k], side[k]);
        if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
        e = cudaStreamWaitEvent(st, join[k], 0);
        if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    }
    return 0;
}

int launch_compose(int n, int nxi, int nfl, const double* P, const double* diag, const double* cst_res,
                   double* NS0, double* NS1p, double* NS1m, double* SGV0, double* SGA0, double* SGV1, double* SGA1,
                   cudaStream_t st)
{
    const size_t total = (size_t)n * n * nxi;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    compose_kernel<<<blocks, 256, 0, st>>>(n, nxi, nfl, P, diag, cst_res, NS0, NS1p, NS1m, SGV0, SGA0, SGV1, SGA1);
    TKB_LAUNCH_CHECK();
    return 0;
}

int launch_eval_kernels(int npts, const double* xs, const double* ys, const double* xis, double* out, cudaStream_t st)
{
    eval_kernels_kernel<<<(npts + 63) / 64, 64, 0, st>>>(npts, xs, ys, xis, out);
    TKB_LAUNCH_CHECK();
    return 0;
}
int launch_eval_scalar(int which, int npts, const double* a, const double* b, double* out, cudaStream_t st)
{
    eval_scalar_kernel<<<(npts + 127) / 128, 128, 0, st>>>(which, npts, a, b, out);
    TKB_LAUNCH_CHECK();
    return 0;
}

}  // namespace tkb
