// tkb_internal.h -- declarations shared by the translation units of libtiktaalik_b200.so
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace tkb {

// per-(xi, i) diagonal record written by K1
enum DiagSlot : int { DIAG_FX = 0, DIAG_X, DIAG_K0QQ, DIAG_A0GG, DIAG_A0GGN, DIAG_N = 8 };

int set_error(const char* msg);          // stores the message, returns 1
void nvtx_push(const char* name);        // NVTX ranges around the stages K1 .. K5 (libnvToolsExt bound at run time; no-ops without it)
void nvtx_pop();
extern long long g_launches;             // kernel launches issued by the library

// assemble.cu
int set_grid_variant(int eta_recip_f32, int x0_recip_f32);
size_t pre_table_doubles(size_t rows);
int launch_pre_table(int n, int nxi, const double* xi_dev, int grid_type, double* tab, cudaStream_t st);
int launch_assemble_fixed(int n, int nxi, const double* xi_dev, int grid_type, double* P, double* diag, cudaStream_t st);
size_t qags_lists_bytes(int sms);
int launch_assemble_cst(int n, int nxi, const double* xi_dev, int grid_type, void* lists, int sms, int queue,
                        unsigned long long* counters, double* cst_res, int* cst_last, int* cst_ifail,
                        cudaStream_t st, cudaStream_t* side, cudaEvent_t fork, cudaEvent_t* join);
int launch_compose(int n, int nxi, int nfl, const double* P, const double* diag, const double* cst_res,
                   double* NS0, double* NS1p, double* NS1m, double* SGV0, double* SGA0, double* SGV1, double* SGA1,
                   cudaStream_t st);
int launch_task_classes(int c, int npts, const double* xs, const double* ys, const double* xis, int* cls, double* full,
                        double* erbl, cudaStream_t st);
int launch_eval_integrand(int c, int npts, const double* xs, const double* ys, const double* xis, double* out, cudaStream_t st);
int launch_eval_kernels(int npts, const double* xs, const double* ys, const double* xis, double* out, cudaStream_t st);
int launch_eval_scalar(int which, int npts, const double* a, const double* b, double* out, cudaStream_t st);

// rk4.cu
// C[b] = D[b] + alpha * A[b] B[b]   (D may be null), row-major, batch b = 0..batch-1;
// B's batch index is b / bdivB (one right-hand side shared by bdivB consecutive A's)
int launch_gemm(int M, int N, int K, double alpha,
                const double* A, int lda, int64_t strideA,
                const double* B, int ldb, int64_t strideB, int bdivB,
                const double* D, int ldd, int64_t strideD,
                double* C, int ldc, int64_t strideC, int batch, cudaStream_t st);
int launch_build_P(size_t count, double a_i, double a_m, double a_e, int nlo, const double* K0, const double* K1,
                   double* Pi, double* Pm, double* Pe, cudaStream_t st);
int launch_rk4_combine(int m, int batch, double h, const double* W1, const double* W2, const double* W3,
                       const double* W4, double* S, cudaStream_t st);
int launch_set_identity(int m, int batch, double* M, int64_t stride, cudaStream_t st);
int launch_scale_add(size_t count, double a, double a2, int nlo, const double* K0, int64_t ld0, const double* K1,
                     int n, int m, int roff, int coff, double* out, cudaStream_t st);

int launch_dfma_peak(int blocks, int iters, double* scratch, cudaStream_t st);

// apply.cu
size_t evolve_workspace_doubles(int n, int nxi, int nq, int64_t S, int nfl);
int launch_evolve_fixed_flavor(int n, int nxi, int nq, int64_t S, int nfl, int F, int Fout, const double* T,
                               const double* N, const double* M_sg, const double* M_nsp, const double* M_nsm,
                               const double* gpd, double* out, int64_t ldq, int q0, int skip, double* work,
                               cudaStream_t st);
int launch_apply(int m, int nxi, int nq2, int64_t S, const double* M, const double* g, int64_t ldg, int64_t xig,
                 double* out, cudaStream_t st);

// exchange.cu
int launch_push_slabs(const double* local, double* const* peer_base, size_t slab_doubles, int nxi, int world, int rank,
                      int k_first, cudaStream_t st);

// gk.cu
int launch_gk(int which, int nx, int nxi, int nt, const double* x, const double* xi, const double* t, double* out,
              cudaStream_t st);

// alpha.cpp (host): alpha_qcd.f90:22-127
double alpha_qcd(double Q2);
int neff(double Q2);

}  // namespace tkb
