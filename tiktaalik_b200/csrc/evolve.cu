// evolve.cu -- the Evolver's fixed-flavour step fused around the application kernel.
//
// Reference: tiktaalik/evolution.py:495-514 (evolve_fixed_flavor) = make_evo_basis (:435-462) -> three
// np.einsum('ijkl,jk...->ik...l') applications (:507-509) -> make_flv_basis (:465-492) -> zero padding
// (:532-540), about a dozen full-size numpy temporaries per threshold segment.  Here: one decoupling
// kernel, the three applications (apply.cu), one recoupling kernel that writes straight into the caller's
// [flavour][x][xi][Q2][sample] array at the segment's Q2 rows.
//
//   decouple:  q_f(x) -+ q_f(-x) for the nfl active quark rows (the x grid is symmetric: -x_j = x_{n-1-j}),
//              (Sigma, T3, T8, ...) = Tpls (q - qflip), the singlet stacked on the gluon row;
//   recouple:  (q - qflip)_f = Tpls^-1 (Sigma', T'), q_f = ((q - qflip)_f + (q + qflip)_f') / 2, gluon from the lower
//              half of the singlet vector, rows beyond nfl zero.
// (The reference calls q + qflip "qmin" and q - qflip "qpls"; they evolve with M_NSm and M_NSp.)
#include <cuda_runtime.h>
#include "tkb_internal.h"

namespace tkb {

struct FlavourRot { double T[25], N[25]; };    // Tpls (nfl x nfl, row-major) and its inverse, as the host computed them

// gpd [F][n][nxi][S] -> sg_in [2n][nxi][S], t_in [n][nxi][(nfl-1) S], qm_in [n][nxi][nfl S]
__global__ void evolve_decouple_kernel(int n, int nxi, int64_t S, int nfl, int F, FlavourRot R,
                                       const double* __restrict__ gpd, double* __restrict__ sg_in,
                                       double* __restrict__ t_in, double* __restrict__ qm_in)
{
    const int64_t per_row = (int64_t)nxi * S;
    const int64_t total = (int64_t)n * per_row;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / per_row, rem = e - j * per_row;          // rem = k * S + s
        const int64_t k = rem / S, s = rem - k * S;
        const int64_t jf = (int64_t)n - 1 - j;
        double qp[5];
#pragma unroll
        for (int f = 0; f < 5; f++) {
            if (f < nfl) {
                const bool have = (f + 1) < F;                            // missing flavours are zero rows
                const double q = have ? gpd[((int64_t)(f + 1) * n + j) * per_row + rem] : 0.0;
                const double qf = have ? gpd[((int64_t)(f + 1) * n + jf) * per_row + rem] : 0.0;
                qm_in[(j * nxi + k) * ((int64_t)nfl * S) + (int64_t)f * S + s] = q + qf;
                qp[f] = q - qf;
            } else qp[f] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < 5; i++) {
            if (i < nfl) {
                double st = 0.0;
#pragma unroll
                for (int f = 0; f < 5; f++)
                    if (f < nfl) st = st + R.T[i * nfl + f] * qp[f];
                if (i == 0) sg_in[j * per_row + rem] = st;
                else t_in[(j * nxi + k) * ((int64_t)(nfl - 1) * S) + (int64_t)(i - 1) * S + s] = st;
            }
        }
        sg_in[((int64_t)n + j) * per_row + rem] = gpd[j * per_row + rem];      // gluon row
    }
}

// sg_out [xi][Q][2n][S], t_out [xi][Q][n][(nfl-1) S], qm_out [xi][Q][n][nfl S]
//   -> out[f][i][xi][q0 + q - skip][s], q = skip .. nq-1, f = 0 .. Fout-1; out has ldq Q2 rows per (f, i, xi).
// The sample index stays the fastest one on both sides, so reads and writes are fully coalesced; the host layer
// hands the result out as a permuted view of logical shape (flavour, x, xi, sample, Q2).
__global__ void __launch_bounds__(256)
evolve_recouple_kernel(int n, int nxi, int nq, int64_t S, int nfl, int Fout, FlavourRot R,
                       const double* __restrict__ sg_out, const double* __restrict__ t_out,
                       const double* __restrict__ qm_out, double* __restrict__ out, int64_t ldq, int q0, int skip)
{
    const int nqk = nq - skip;
    const int64_t total = (int64_t)nxi * n * nqk * S;
    const int64_t fstride = (int64_t)n * nxi * ldq * S;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = e % S;
        int64_t r = e / S;
        const int qq = (int)(r % nqk); r /= nqk;
        const int64_t i = r % n;
        const int64_t k = r / n;
        const int64_t bm = k * nq + (qq + skip);                         // (xi, Q2) matrix index
        double st[5];
        st[0] = sg_out[(bm * 2 * n + i) * S + s];
#pragma unroll
        for (int f = 1; f < 5; f++)
            st[f] = (f < nfl) ? t_out[(bm * n + i) * ((int64_t)(nfl - 1) * S) + (int64_t)(f - 1) * S + s] : 0.0;
        const int64_t obase = (((i * nxi + k) * ldq) + (q0 + qq)) * S + s;
        out[obase] = sg_out[(bm * 2 * n + n + i) * S + s];               // gluon
#pragma unroll
        for (int f = 0; f < 5; f++) {
            if (f < nfl) {
                double qp = 0.0;
#pragma unroll
                for (int m = 0; m < 5; m++)
                    if (m < nfl) qp = qp + R.N[f * nfl + m] * st[m];
                const double qm = qm_out[(bm * n + i) * ((int64_t)nfl * S) + (int64_t)f * S + s];
                out[(int64_t)(f + 1) * fstride + obase] = (qp + qm) / 2.0;
            }
        }
        for (int f = nfl + 1; f < Fout; f++) out[(int64_t)f * fstride + obase] = 0.0;
    }
}

static int ew_grid(int64_t total)
{
    int64_t b = (total + 255) / 256;
    if (b > 148LL * 16) b = 148LL * 16;
    if (b < 1) b = 1;
    return (int)b;
}

size_t evolve_workspace_doubles(int n, int nxi, int nq, int64_t S, int nfl)
{
    return (size_t)n * nxi * (size_t)S * (size_t)(2 * nfl + 1) * (size_t)(1 + nq);
}

int launch_evolve_fixed_flavor(int n, int nxi, int nq, int64_t S, int nfl, int F, int Fout, const double* T,
                               const double* N, const double* M_sg, const double* M_nsp, const double* M_nsm,
                               const double* gpd, double* out, int64_t ldq, int q0, int skip, double* work,
                               cudaStream_t st)
{
    if (S <= 0 || nq - skip <= 0) return 0;
    FlavourRot R;
    for (int e = 0; e < 25; e++) { R.T[e] = (e < nfl * nfl) ? T[e] : 0.0; R.N[e] = (e < nfl * nfl) ? N[e] : 0.0; }
    const size_t u = (size_t)n * nxi * (size_t)S;
    double* sg_in = work;
    double* t_in = sg_in + 2 * u;
    double* qm_in = t_in + (size_t)(nfl - 1) * u;
    double* sg_out = qm_in + (size_t)nfl * u;
    double* t_out = sg_out + 2 * u * nq;
    double* qm_out = t_out + (size_t)(nfl - 1) * u * nq;
    evolve_decouple_kernel<<<ew_grid((int64_t)u), 256, 0, st>>>(n, nxi, S, nfl, F, R, gpd, sg_in, t_in, qm_in);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    if (launch_apply(2 * n, nxi, nq, S, M_sg, sg_in, (int64_t)nxi * S, S, sg_out, st)) return 1;
    if (nfl > 1 && launch_apply(n, nxi, nq, (int64_t)(nfl - 1) * S, M_nsp, t_in, (int64_t)nxi * (nfl - 1) * S,
                            (int64_t)(nfl - 1) * S, t_out, st)) return 1;
    if (launch_apply(n, nxi, nq, (int64_t)nfl * S, M_nsm, qm_in, (int64_t)nxi * nfl * S, (int64_t)nfl * S, qm_out, st)) return 1;
    evolve_recouple_kernel<<<ew_grid((int64_t)u * (nq - skip)), 256, 0, st>>>(n, nxi, nq, S, nfl, Fout, R, sg_out, t_out,
                                                                           qm_out, out, ldq, q0, skip);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

}  // namespace tkb
