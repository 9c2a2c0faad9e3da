// capi.cu -- the implicit per-process context and the C ABI (include/tiktaalik_b200.h).
//
// Mirrors the reference's module state (matevo.f90:27-46): one live set of kernel matrices
// (invalidated by make_kernels) and one set of evolution matrices (replaced by make_matrices).
// The library owns all device buffers; inputs (xi, Q2) are small host arrays.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include "../../include/tiktaalik_b200.h"
#include "tkb_internal.h"

namespace tkb {

static std::string g_err;
long long g_launches = 0;

int set_error(const char* msg)
{
    g_err = msg ? msg : "unknown error";
    return 1;
}

// NVTX v3 is header-only: the calls are no-ops unless a profiler injects itself (nsys, ncu --nvtx)
void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

#define CUDA_OK(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            char buf_[512];                                                            \
            snprintf(buf_, sizeof buf_, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return set_error(buf_);                                                    \
        }                                                                              \
    } while (0)

constexpr int NKERN_ = 22;

struct Composed {              // one nfl slice (matevo.f90:35-39), row-major per xi
    double* a[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    bool valid = false;
};

struct Context {
    int device = -1;            // the CUDA device the context (streams, events, buffers) is bound to
    int nx = 0, nxi = 0, grid_type = 1, nq2 = 0, lnlo = 0;
    std::vector<double> xi, q2;
    double* d_xi = nullptr;
    double* P = nullptr;        // [22][nxi][n][n]
    double* diag = nullptr;     // [nxi][n][DIAG_N]
    double* pre_tab = nullptr;  // per (xi, i) row: one-argument logs / dilogs of the fixed point, x_i, 0.5 (1 +- x_i / xi)
    double* cst_res = nullptr;  // [nxi][n][7][3]
    int* cst_last = nullptr;
    int* cst_ifail = nullptr;
    Composed comp[3];           // nfl = 3, 4, 5
    double* M[4] = { nullptr, nullptr, nullptr, nullptr };   // NS+, NS-, SG V, SG A : [xi][Q2][i][j]
    void* scratch = nullptr;    // QAGS spill lists + counters, kept across calls
    size_t scratch_bytes = 0;
    double* ws = nullptr;       // RK4 workspace, kept across calls
    size_t ws_bytes = 0;
    size_t M_elems = 0;         // elements of M[0] as allocated
    cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    bool ev_k = false, ev_m = false;
    cudaStream_t side[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    cudaStream_t k1_stream = nullptr;
    cudaEvent_t k1_fork = nullptr, k1_done = nullptr;
    bool k1_overlapped = false;
    cudaEvent_t fork = nullptr, join[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
};
static Context g;

static void free_composed()
{
    for (auto& c : g.comp) {
        for (auto& p : c.a) { if (p) cudaFree(p); p = nullptr; }
        c.valid = false;
    }
}
static void invalidate_matrices()
{
    g.nq2 = 0;
    g.q2.clear();
}
static void free_matrices()
{
    for (auto& p : g.M) { if (p) cudaFree(p); p = nullptr; }
    g.M_elems = 0;
    invalidate_matrices();
}
static void free_kernels()
{
    if (g.d_xi) cudaFree(g.d_xi);
    if (g.P) cudaFree(g.P);
    if (g.diag) cudaFree(g.diag);
    if (g.pre_tab) cudaFree(g.pre_tab);
    g.pre_tab = nullptr;
    if (g.cst_res) cudaFree(g.cst_res);
    if (g.cst_last) cudaFree(g.cst_last);
    if (g.cst_ifail) cudaFree(g.cst_ifail);
    g.d_xi = g.P = g.diag = g.cst_res = nullptr;
    g.cst_last = g.cst_ifail = nullptr;
    free_composed();
    g.nx = g.nxi = 0;
}

static int ensure_events()
{
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    if (g.device >= 0 && g.device != dev)
        return set_error("the library's context is bound to another CUDA device: call tkb_free() on that device first "
                         "(one implicit context per process, like the reference's module state)");
    g.device = dev;
    for (auto& e : g.ev)
        if (!e) CUDA_OK(cudaEventCreate(&e));
    if (!g.fork) {
        CUDA_OK(cudaStreamCreateWithFlags(&g.k1_stream, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&g.k1_fork, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreate(&g.k1_done));
        CUDA_OK(cudaEventCreateWithFlags(&g.fork, cudaEventDisableTiming));
        for (int k = 0; k < 7; k++) {
            CUDA_OK(cudaStreamCreateWithFlags(&g.side[k], cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreateWithFlags(&g.join[k], cudaEventDisableTiming));
        }
    }
    return 0;
}

static int make_kernels_impl(int nx, int nxi, const double* xi_host, int grid_type, int clamp_first, int clamp_last,
                             cudaStream_t st)
{
    if (nx < 6 || (nx % 2) != 0) return set_error("tkb_make_kernels: nx must be even and >= 6");
    if (nxi < 1 || !xi_host) return set_error("tkb_make_kernels: need at least one xi");
    // matevo.f90:70: everything is deallocated first.  Device buffers are kept when the sizes repeat
    // (an xi sweep re-enters with the same nx, nxi): the contents are invalidated, not the allocations.
    const bool same = (g.P != nullptr) && g.nx == nx && g.nxi == nxi;
    invalidate_matrices();
    if (!same) { free_matrices(); free_kernels(); }
    for (auto& c : g.comp) c.valid = false;
    if (ensure_events()) return 1;
    g.nx = nx;
    g.nxi = nxi;
    g.grid_type = grid_type;
    g.xi.assign(xi_host, xi_host + nxi);
    // matevo.f90:688-689
    if (clamp_first && g.xi[0] < 1e-6) g.xi[0] = 1e-6;
    if (clamp_last && g.xi[nxi - 1] > 0.999999) g.xi[nxi - 1] = 0.999999;

    const size_t nn = (size_t)nx * nx * nxi;
    const size_t rows = (size_t)nx * nxi;
    if (!same) {
        CUDA_OK(cudaMalloc(&g.d_xi, sizeof(double) * nxi));
        CUDA_OK(cudaMalloc(&g.P, sizeof(double) * nn * NKERN_));
        CUDA_OK(cudaMalloc(&g.diag, sizeof(double) * rows * DIAG_N));
        CUDA_OK(cudaMalloc(&g.pre_tab, sizeof(double) * pre_table_doubles(rows)));
        CUDA_OK(cudaMalloc(&g.cst_res, sizeof(double) * rows * 21));
        CUDA_OK(cudaMalloc(&g.cst_last, sizeof(int) * rows * 21));
        CUDA_OK(cudaMalloc(&g.cst_ifail, sizeof(int) * rows * 21));
    }
    CUDA_OK(cudaMemcpyAsync(g.d_xi, g.xi.data(), sizeof(double) * nxi, cudaMemcpyHostToDevice, st));

    // K1 (fixed quadrature) and K2 (QAGS constants) are independent.  K2's seven persistent launches go
    // first, K1 runs on its own stream and fills the SMs that K2's draining launches free.
    const bool overlap_k1 = !getenv("TKB_K1_SERIAL");
    if (launch_pre_table(nx, nxi, g.d_xi, grid_type, g.pre_tab, st)) return 1;
    CUDA_OK(cudaEventRecord(g.ev[0], st));
    if (!overlap_k1) {
        nvtx_push("tkb K1 fixed quadrature");
        const int rc1 = launch_assemble_fixed(nx, nxi, g.d_xi, grid_type, g.P, g.diag, st);
        nvtx_pop();
        if (rc1) return 1;
    }
    CUDA_OK(cudaEventRecord(g.ev[1], st));

    // QAGS: one persistent block per SM and constant pulling instances from a counter (assemble.cu).  Form: the
    // role-split pipeline (default) or the work queue (TKB_K2_MODE=queue, kept for comparison)
    int dev = 0, sms = 148;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const char* mode = getenv("TKB_K2_MODE");
    const int queue = (mode && mode[0] == 'q') ? 1 : 0;
    const size_t need = qags_lists_bytes(sms) + 64;                // 7 counters, then the interval-list spill area
    if (g.scratch_bytes < need) {
        if (g.scratch) cudaFree(g.scratch);
        g.scratch = nullptr;
        g.scratch_bytes = 0;
        CUDA_OK(cudaMalloc(&g.scratch, need));
        g.scratch_bytes = need;
    }
    nvtx_push("tkb K2 QAGS constants");
    const int rc = launch_assemble_cst(nx, nxi, g.d_xi, grid_type, (char*)g.scratch + 64, sms, queue,
                                       (unsigned long long*)g.scratch, g.cst_res, g.cst_last, g.cst_ifail, st, g.side,
                                       g.fork, g.join);
    nvtx_pop();
    if (overlap_k1) {
        CUDA_OK(cudaEventRecord(g.k1_fork, st));                 // after the K2 launches were issued: ordering hint only
        CUDA_OK(cudaStreamWaitEvent(g.k1_stream, g.ev[0], 0));   // data dependency: the xi upload
        CUDA_OK(cudaEventRecord(g.ev[5], g.k1_stream));
        nvtx_push("tkb K1 fixed quadrature");
        const int rc1 = launch_assemble_fixed(nx, nxi, g.d_xi, grid_type, g.P, g.diag, g.k1_stream);
        nvtx_pop();
        if (rc1) return 1;
        CUDA_OK(cudaEventRecord(g.k1_done, g.k1_stream));
        CUDA_OK(cudaStreamWaitEvent(st, g.k1_done, 0));
    }
    g.k1_overlapped = overlap_k1;
    CUDA_OK(cudaEventRecord(g.ev[2], st));
    g.ev_k = true;
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc) return 1;
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "kernel assembly failed: %s", cudaGetErrorString(e));
        return set_error(buf);
    }
    return 0;
}

static int ensure_composed(int nfl, cudaStream_t st)
{
    if (!g.P) return set_error("kernels are not initialised: call tkb_make_kernels first");
    if (nfl < 3 || nfl > 5) return set_error("nfl must be 3, 4 or 5");
    Composed& c = g.comp[nfl - 3];
    if (c.valid) return 0;
    const size_t nn = (size_t)g.nx * g.nx * g.nxi;
    for (int k = 0; k < 7; k++)
        if (!c.a[k]) CUDA_OK(cudaMalloc(&c.a[k], sizeof(double) * nn * (k < 3 ? 1 : 4)));
    nvtx_push("tkb K3 compose nfl slice");
    const int rc3 = launch_compose(g.nx, g.nxi, nfl, g.P, g.diag, g.cst_res, c.a[0], c.a[1], c.a[2], c.a[3], c.a[4], c.a[5],
                                   c.a[6], st);
    nvtx_pop();
    if (rc3) return 1;
    c.valid = true;
    return 0;
}

// one matrix type over the whole Q^2 schedule (matevo.f90:485-595)
static int step_type(int m, const std::vector<double>& q2, int nlo, int k0_idx, int k1_idx, double* M, double* ws,
                     cudaStream_t st)
{
    const int nxi = g.nxi, nq2 = (int)q2.size();
    const size_t mm = (size_t)m * m;
    const size_t cnt = mm * nxi;
    double* Pi = ws; double* Pm = Pi + cnt; double* Pe = Pm + cnt;
    double* W2 = Pe + cnt; double* W3 = W2 + cnt; double* W4 = W3 + cnt; double* S = W4 + cnt;
    const int64_t strideM = (int64_t)nq2 * mm;
    if (launch_set_identity(m, nxi, M, strideM, st)) return 1;
    const double two_pi = 2.0 * 3.141592653589793;
    for (int iq = 1; iq < nq2; iq++) {
        const double Q2i = q2[iq - 1], Q2f = q2[iq];
        int nfl = neff(Q2i);
        if (nfl > 5) return set_error("Q2 above the top threshold is outside the reference's nfl range (3..5)");
        if (ensure_composed(nfl, st)) return 1;
        const Composed& c = g.comp[nfl - 3];
        const double h = std::log(Q2f / Q2i);
        const double Q2m = std::sqrt(Q2i * Q2f);
        const double a_i = alpha_qcd(Q2i) / two_pi, a_m = alpha_qcd(Q2m) / two_pi, a_e = alpha_qcd(Q2f) / two_pi;
        if (launch_build_P(cnt, a_i, a_m, a_e, nlo, c.a[k0_idx], c.a[k1_idx], Pi, Pm, Pe, st)) return 1;
        // W2 = Pm + h/2 Pm W1 (W1 = Pi); W3 = Pm + h/2 Pm W2; W4 = Pe + h Pe W3
        if (launch_gemm(m, m, m, 0.5 * h, Pm, m, mm, Pi, m, mm, 1, Pm, m, mm, W2, m, mm, nxi, st)) return 1;
        if (launch_gemm(m, m, m, 0.5 * h, Pm, m, mm, W2, m, mm, 1, Pm, m, mm, W3, m, mm, nxi, st)) return 1;
        if (launch_gemm(m, m, m, h, Pe, m, mm, W3, m, mm, 1, Pe, m, mm, W4, m, mm, nxi, st)) return 1;
        if (launch_rk4_combine(m, nxi, h, Pi, W2, W3, W4, S, st)) return 1;
        // M(Q_k) = S M(Q_{k-1})
        if (launch_gemm(m, m, m, 1.0, S, m, mm, M + (size_t)(iq - 1) * mm, m, strideM, 1, nullptr, 0, 0,
                        M + (size_t)iq * mm, m, strideM, nxi, st))
            return 1;
    }
    return 0;
}

static int make_matrices_impl(int nq2, const double* q2_host, int nlo, cudaStream_t st)
{
    if (!g.P) return set_error("tkb_make_matrices: call tkb_make_kernels first");
    if (nq2 < 2 || !q2_host) return set_error("tkb_make_matrices: need at least two Q2 values");
    const int n = g.nx, nxi = g.nxi;
    const size_t nn = (size_t)n * n * nxi * nq2;
    if (g.M_elems != nn) {
        free_matrices();
        for (int t = 0; t < 4; t++) CUDA_OK(cudaMalloc(&g.M[t], sizeof(double) * nn * (t < 2 ? 1 : 4)));
        g.M_elems = nn;
    }
    // the schedule becomes visible to the getters only after the stepping has succeeded
    invalidate_matrices();
    const std::vector<double> q2v(q2_host, q2_host + nq2);
    const int lnlo = nlo ? 1 : 0;
    const size_t ws_need = sizeof(double) * (size_t)4 * n * n * nxi * 7;
    if (g.ws_bytes < ws_need) {
        if (g.ws) cudaFree(g.ws);
        g.ws = nullptr;
        g.ws_bytes = 0;
        CUDA_OK(cudaMalloc(&g.ws, ws_need));
        g.ws_bytes = ws_need;
    }
    double* ws = g.ws;
    CUDA_OK(cudaEventRecord(g.ev[3], st));
    int rc = 0;
    nvtx_push("tkb K4 RK4 stepping");
    // composed-array indices: 0 NS0, 1 NS1p, 2 NS1m, 3 SGV0, 4 SGA0, 5 SGV1, 6 SGA1
    rc = rc || step_type(n, q2v, lnlo, 0, 1, g.M[0], ws, st);
    if (!rc) {
        if (lnlo) rc = step_type(n, q2v, lnlo, 0, 2, g.M[1], ws, st);
        else {   // at LO the minus-type matrix is the plus-type one (matevo.f90:517-520 with l_nlo false)
            cudaError_t e = cudaMemcpyAsync(g.M[1], g.M[0], sizeof(double) * nn, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) rc = set_error(cudaGetErrorString(e));
        }
    }
    rc = rc || step_type(2 * n, q2v, lnlo, 3, 5, g.M[2], ws, st);
    rc = rc || step_type(2 * n, q2v, lnlo, 4, 6, g.M[3], ws, st);
    nvtx_pop();
    cudaEventRecord(g.ev[4], st);
    g.ev_m = true;
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc) return 1;
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "evolution-matrix stepping failed: %s", cudaGetErrorString(e));
        return set_error(buf);
    }
    g.q2 = q2v;
    g.nq2 = nq2;
    g.lnlo = lnlo;
    return 0;
}

static const double* matrix_ptr(int type, int nstype, int& m)
{
    m = g.nx;
    switch (type) {
    case TKB_VNS: return nstype == 1 ? g.M[0] : (nstype == -1 ? g.M[1] : nullptr);
    case TKB_ANS: return nstype == 1 ? g.M[1] : (nstype == -1 ? g.M[0] : nullptr);    // matevo.f90:263-272
    case TKB_VSG: m = 2 * g.nx; return g.M[2];
    case TKB_ASG: m = 2 * g.nx; return g.M[3];
    }
    return nullptr;
}

}  // namespace tkb

using namespace tkb;

extern "C" {

const char* tkb_last_error(void) { return g_err.c_str(); }
int tkb_version(void) { return 100; }

int tkb_make_kernels(int nx, int nxi, const double* xi_host, int grid_type, void* stream)
{
    return make_kernels_impl(nx, nxi, xi_host, grid_type, 1, 1, (cudaStream_t)stream);
}
int tkb_make_kernels_shard(int nx, int nxi, const double* xi_host, int grid_type, int clamp_first, int clamp_last,
                           void* stream)
{
    return make_kernels_impl(nx, nxi, xi_host, grid_type, clamp_first, clamp_last, (cudaStream_t)stream);
}
int tkb_make_matrices(int nq2, const double* q2_host, int nlo, void* stream)
{
    return make_matrices_impl(nq2, q2_host, nlo, (cudaStream_t)stream);
}
int tkb_get_nx(void) { return g.nx; }
int tkb_get_nxi(void) { return g.nxi; }
int tkb_get_nq2(void) { return g.nq2; }
int tkb_get_lnlo(void) { return g.lnlo; }

int tkb_evomatrix(int type, int nstype, const double** dev_ptr, int64_t shape[4], int64_t strides[4])
{
    if (!g.M[0] || g.nq2 == 0) return set_error("evolution matrices are not initialised: call tkb_make_matrices first");
    int m;
    const double* p = matrix_ptr(type, nstype, m);
    if (!p) return set_error("tkb_evomatrix: bad type / nstype");
    *dev_ptr = p;
    const int64_t mm = (int64_t)m * m;
    shape[0] = m; shape[1] = m; shape[2] = g.nxi; shape[3] = g.nq2;
    strides[0] = m; strides[1] = 1; strides[2] = (int64_t)g.nq2 * mm; strides[3] = mm;
    return 0;
}
int tkb_evomatrix_copy(int type, int nstype, double* dev_out, void* stream)
{
    if (!g.M[0] || g.nq2 == 0) return set_error("evolution matrices are not initialised: call tkb_make_matrices first");
    int m;
    const double* p = matrix_ptr(type, nstype, m);
    if (!p) return set_error("tkb_evomatrix_copy: bad type / nstype");
    const size_t bytes = sizeof(double) * (size_t)m * m * g.nxi * g.nq2;
    CUDA_OK(cudaMemcpyAsync(dev_out, p, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int tkb_evokernel(int type, double q2, int nfl, int nlo, int nstype, double* dev_out, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (type < 0 || type > 7) return set_error("tkb_evokernel: bad type");
    if (ensure_composed(nfl, st)) return 1;
    const Composed& c = g.comp[nfl - 3];
    const int n = g.nx;
    const double a = alpha_qcd(q2) / (2.0 * 3.141592653589793);
    const bool isA = type >= 4;
    const int blk = type & 3;
    const size_t count = (size_t)n * n * g.nxi;
    if (blk == 0) {
        // matevo.f90:124-143, 183-203: K = a K_NS_0 (+ a^2 {NS_1p | NS_1m | QQ block of SG_1})
        const double* K1 = nullptr;
        int m1 = n;
        int use_nlo = nlo;
        const int t = isA ? -nstype : nstype;
        if (nstype == 0) { K1 = c.a[isA ? 6 : 5]; m1 = 2 * n; }
        else if (t == 1) K1 = c.a[1];
        else if (t == -1) K1 = c.a[2];
        else use_nlo = 0;                       // select case without a matching case: nothing added
        if (!use_nlo || m1 == n)
            return launch_scale_add(count, a, a * a, use_nlo, c.a[0], 0, K1 ? K1 : c.a[0], n, n, 0, 0, dev_out, st);
        // LO part from the n x n array, NLO part from the 2n x 2n array: two passes
        // (a K0) + (a^2 K1) with K0 = QQ block of SG_0 of the same type, which equals K_NS_0 (matevo.f90:370, 389)
        return launch_scale_add(count, a, a * a, 1, c.a[isA ? 4 : 3], 0, K1, n, 2 * n, 0, 0, dev_out, st);
    }
    const int roff = (blk == 2 || blk == 3) ? n : 0, coff = (blk == 1 || blk == 3) ? n : 0;
    return launch_scale_add(count, a, a * a, nlo, c.a[isA ? 4 : 3], 0, c.a[isA ? 6 : 5], n, 2 * n, roff, coff, dev_out, st);
}

int tkb_pixelspace(int nx, double xi, int grid_type, double* host_out)
{
    if (nx < 1 || !host_out) return set_error("tkb_pixelspace: bad arguments");
    for (int ix = 1; ix <= nx; ix++) {
        // f90wrap/matrices.f90:28: float32 quotient, float64 subtraction
        const double eta = (double)((float)(2 * ix - 1) / (float)nx) - 1.0;
        double x = eta;
        if (grid_type == 2) {
            if (eta < -0.5) x = -std::pow(xi * xi, 1.0 + eta);
            else if (eta <= 0.5) x = 2.0 * xi * eta;
            else x = std::pow(xi * xi, 1.0 - eta);
        }
        host_out[ix - 1] = x;
    }
    return 0;
}

int tkb_alpha(int n, const double* q2_host, double* out_host)
{
    for (int i = 0; i < n; i++) out_host[i] = alpha_qcd(q2_host[i]);
    return 0;
}

int tkb_apply(int m, int nxi, int nq2, int64_t S, const double* M_dev, const double* g_dev, int64_t ldg,
              double* out_dev, void* stream)
{
    if (m < 1 || nxi < 1 || nq2 < 1 || S < 0 || !M_dev || !g_dev || !out_dev) return set_error("tkb_apply: bad arguments");
    nvtx_push("tkb K5 apply");
    const int rc = launch_apply(m, nxi, nq2, S, M_dev, g_dev, ldg, S, out_dev, (cudaStream_t)stream);
    nvtx_pop();
    return rc;
}

int tkb_apply_strided(int m, int nxi, int nq2, int64_t S, const double* M_dev, const double* g_dev, int64_t ldg,
                      int64_t xig, double* out_dev, void* stream)
{
    if (m < 1 || nxi < 1 || nq2 < 1 || S < 0 || !M_dev || !g_dev || !out_dev || xig < 0 || ldg < 0)
        return set_error("tkb_apply_strided: bad arguments");
    nvtx_push("tkb K5 apply");
    const int rc = launch_apply(m, nxi, nq2, S, M_dev, g_dev, ldg, xig, out_dev, (cudaStream_t)stream);
    nvtx_pop();
    return rc;
}

int64_t tkb_evolve_workspace(int nx, int nxi, int nq2, int64_t S, int nfl)
{
    if (nx < 1 || nxi < 1 || nq2 < 1 || S < 0 || nfl < 1 || nfl > 5) return -1;
    return (int64_t)evolve_workspace_doubles(nx, nxi, nq2, S, nfl);
}

int tkb_evolve_fixed_flavor(int nx, int nxi, int nq2, int64_t S, int nfl, int F, int Fout, const double* tpls_host,
                            const double* tpls_inv_host, const double* M_sg_dev, const double* M_nsp_dev,
                            const double* M_nsm_dev, const double* gpd_dev, double* out_dev, int64_t ldq, int q0, int skip,
                            double* work_dev, void* stream)
{
    if (nx < 1 || nxi < 1 || nq2 < 1 || S < 0 || nfl < 1 || nfl > 5 || F < 1 || Fout < nfl + 1 || skip < 0 ||
        skip > nq2 || q0 < 0 || ldq < q0 + (nq2 - skip) || !tpls_host || !tpls_inv_host || !M_sg_dev || !M_nsp_dev ||
        !M_nsm_dev || !gpd_dev || !out_dev || !work_dev)
        return set_error("tkb_evolve_fixed_flavor: bad arguments");
    nvtx_push("tkb K7 evolve_fixed_flavor");
    const int rc = launch_evolve_fixed_flavor(nx, nxi, nq2, S, nfl, F, Fout, tpls_host, tpls_inv_host, M_sg_dev, M_nsp_dev,
                                              M_nsm_dev, gpd_dev, out_dev, ldq, q0, skip, work_dev, (cudaStream_t)stream);
    nvtx_pop();
    return rc;
}

int tkb_gk(int which, int nx, int nxi, int nt, const double* x_dev, const double* xi_dev, const double* t_dev,
           double* out_dev, void* stream)
{
    if (which < 0 || which > 7 || nx < 0 || nxi < 0 || nt < 0) return set_error("tkb_gk: bad arguments");
    return launch_gk(which, nx, nxi, nt, x_dev, xi_dev, t_dev, out_dev, (cudaStream_t)stream);
}

int tkb_free(void)
{
    free_matrices();
    free_kernels();
    if (g.scratch) cudaFree(g.scratch);
    if (g.ws) cudaFree(g.ws);
    g.scratch = nullptr; g.ws = nullptr;
    g.scratch_bytes = g.ws_bytes = 0;
    // streams and events belong to the device the context was bound to; after tkb_free another device may bind
    for (auto& e : g.ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (int k = 0; k < 7; k++) {
        if (g.side[k]) cudaStreamDestroy(g.side[k]);
        if (g.join[k]) cudaEventDestroy(g.join[k]);
        g.side[k] = nullptr; g.join[k] = nullptr;
    }
    if (g.k1_stream) cudaStreamDestroy(g.k1_stream);
    if (g.k1_fork) cudaEventDestroy(g.k1_fork);
    if (g.k1_done) cudaEventDestroy(g.k1_done);
    if (g.fork) cudaEventDestroy(g.fork);
    g.k1_stream = nullptr; g.k1_fork = g.k1_done = g.fork = nullptr;
    g.ev_k = g.ev_m = false;
    g.device = -1;
    return 0;
}

int tkb_qags_info(const double** res_dev, const int** last_dev, const int** ifail_dev)
{
    if (!g.cst_res) return set_error("kernels are not initialised");
    *res_dev = g.cst_res;
    *last_dev = g.cst_last;
    *ifail_dev = g.cst_ifail;
    return 0;
}

int tkb_raw_kernel(int which, int nfl, const double** dev_ptr, void* stream)
{
    if (which < 0 || which > 6) return set_error("tkb_raw_kernel: bad index");
    if (ensure_composed(nfl, (cudaStream_t)stream)) return 1;
    *dev_ptr = g.comp[nfl - 3].a[which];
    return 0;
}

int tkb_eval_kernels(int npts, const double* x_dev, const double* y_dev, const double* xi_dev, double* out_dev,
                     void* stream)
{
    return launch_eval_kernels(npts, x_dev, y_dev, xi_dev, out_dev, (cudaStream_t)stream);
}
int tkb_eval_integrand(int c, int npts, const double* x_dev, const double* y_dev, const double* xi_dev, double* out_dev,
                       void* stream)
{
    if (npts < 0 || !x_dev || !y_dev || !xi_dev || !out_dev) return set_error("tkb_eval_integrand: bad arguments");
    if (npts == 0) return 0;
    return launch_eval_integrand(c, npts, x_dev, y_dev, xi_dev, out_dev, (cudaStream_t)stream);
}
int tkb_eval_scalar(int which, int npts, const double* a_dev, const double* b_dev, double* out_dev, void* stream)
{
    return launch_eval_scalar(which, npts, a_dev, b_dev, out_dev, (cudaStream_t)stream);
}

int tkb_task_classes(int c, int npts, const double* x_dev, const double* y_dev, const double* xi_dev, int* cls_dev,
                     double* full_dev, double* erbl_dev, void* stream)
{
    if (npts < 0 || !x_dev || !y_dev || !xi_dev || !cls_dev || !full_dev || !erbl_dev) return set_error("tkb_task_classes: bad arguments");
    if (npts == 0) return 0;
    return launch_task_classes(c, npts, x_dev, y_dev, xi_dev, cls_dev, full_dev, erbl_dev, (cudaStream_t)stream);
}

int tkb_stage_ms(double ms[3])
{
    ms[0] = ms[1] = ms[2] = 0.0;
    float t = 0.f;
    if (g.ev_k) {
        CUDA_OK(cudaEventSynchronize(g.ev[2]));
        if (g.k1_overlapped) {     // K1 ran concurrently: its own span, and the whole assembly span for K2
            CUDA_OK(cudaEventElapsedTime(&t, g.ev[5], g.k1_done)); ms[0] = t;
        } else {
            CUDA_OK(cudaEventElapsedTime(&t, g.ev[0], g.ev[1])); ms[0] = t;
        }
        CUDA_OK(cudaEventElapsedTime(&t, g.ev[1], g.ev[2])); ms[1] = t;
    }
    if (g.ev_m) {
        CUDA_OK(cudaEventSynchronize(g.ev[4]));
        CUDA_OK(cudaEventElapsedTime(&t, g.ev[3], g.ev[4])); ms[2] = t;
    }
    return 0;
}

long long tkb_launch_count(void) { return g_launches; }

int tkb_set_variant(int eta_recip_f32, int x0_recip_f32)
{
    // takes effect at the next tkb_make_kernels; kernels assembled before are not touched
    return set_grid_variant(eta_recip_f32, x0_recip_f32);
}

int tkb_fp64_peak(double* tflops)
{
    int dev = 0, sms = 148;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* scratch = nullptr;
    CUDA_OK(cudaMalloc(&scratch, 64));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    const int blocks = sms * 8, iters = 1 << 16;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_OK(cudaEventRecord(e0, 0));
        if (launch_dfma_peak(blocks, iters, scratch, 0)) return 1;
        CUDA_OK(cudaEventRecord(e1, 0));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(scratch);
    *tflops = best;
    return 0;
}

}  // extern "C"
