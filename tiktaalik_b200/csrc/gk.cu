// gk.cu -- Goloskokov-Kroll model GPD H(x, xi, t) on the device (input generator of the examples).
//
// Reference: f90src/model/gk.f90:27-241, wrappers tiktaalik/f90wrap/model.f90:16-158.  One thread per
// (x, xi, t) point; z**p uses the library's own pow (tkb_fp.cuh), so values are bit-identical to the
// oracle's restatement.  Compiled with -fmad=false like the assembly unit.
#include <cuda_runtime.h>
#include "tkb_fp.cuh"
#include "tkb_internal.h"

namespace tkb {

__constant__ double c_gkC[4][4] = {            // gk.f90:59-64, C(i, j): i = 0 gluon, 1 u_val, 2 d_val, 3 strange
    { 2.23, 5.43, -34.0, 40.6 },
    { 1.52, 2.88, -0.095, 0.0 },
    { 0.76, 3.11, -3.99, 0.0 },
    { 0.123, -0.327, 0.692, -0.486 } };

TKB_DEV double gk_pw(int i, int j, int w, double t)            // gk.f90:233-239
{
    const double a0_val = 0.48, ap_val = 0.9, a0_sea = 1.1, ap_sea = 0.15;
    const double delta = (i == 0) ? (a0_sea - 1.0) : ((i == 3) ? a0_sea : a0_val);
    const double ap = (i == 0 || i == 3) ? ap_sea : ap_val;
    return (double)__fadd_rn(__fdiv_rn((float)j, 2.0f), (float)w) - delta - ap * t;
}
TKB_DEV double gk_kw(double x, double xi, int w)               // gk.f90:193-209
{
    if (w == 1) return 1.0 - (x * x) / (xi * xi);
    if (w == 2) return 2.0 * (x / (xi * xi) - 1.0);
    return 1.0 - 1.0 / (xi * xi);
}
TKB_DEV double gk_ew(double x, double xi, int w)               // gk.f90:211-231
{
    switch (w) {
    case 1: return gk_kw(x, xi, 1) * gk_kw(x, xi, 1);
    case 2: return 2.0 * gk_kw(x, xi, 1) * gk_kw(x, xi, 2);
    case 3: return gk_kw(x, xi, 2) * gk_kw(x, xi, 2) + 2.0 * gk_kw(x, xi, 1) * gk_kw(x, xi, 3);
    case 4: return 2.0 * gk_kw(x, xi, 2) * gk_kw(x, xi, 3);
    }
    return gk_kw(x, xi, 3) * gk_kw(x, xi, 3);
}
__device__ double gk_n1_term(double x, double xi, double t, double z, int i)    // gk.f90:161-175
{
    double H = 0.0;
    for (int j = 0; j <= 3; j++)
        for (int w = 1; w <= 3; w++) {
            const double p = gk_pw(i, j, w, t);
            H = H + c_gkC[i][j] * gk_kw(x, xi, w) * pow_(z, p) / p;
        }
    return H * 0.75 / xi;
}
__device__ double gk_n2_term(double x, double xi, double t, double z, int i)    // gk.f90:177-191
{
    double H = 0.0;
    for (int j = 0; j <= 3; j++)
        for (int w = 1; w <= 5; w++) {
            const double p = gk_pw(i, j, w, t);
            H = H + c_gkC[i][j] * gk_ew(x, xi, w) * pow_(z, p) / p;
        }
    return H * 1.875 / xi / 2.0;
}
__device__ double gk_n1(double x, double xi, double t, int i)                   // gk.f90:125-139
{
    const double zi = (x - xi) / (1.0 - xi), zf = (x + xi) / (1.0 + xi);
    if (x > xi) return gk_n1_term(x, xi, t, zf, i) - gk_n1_term(x, xi, t, zi, i);
    if (x > -xi) return gk_n1_term(x, xi, t, zf, i);
    return 0.0;
}
__device__ double gk_n2(double x, double xi, double t, int i, double sgn)       // gk.f90:141-159
{
    const double zi = (x - xi) / (1.0 - xi), zf = (x + xi) / (1.0 + xi);
    const double zj = (-x - xi) / (1.0 - xi), zm = (-x + xi) / (1.0 + xi);
    if (x > xi) return gk_n2_term(x, xi, t, zf, i) - gk_n2_term(x, xi, t, zi, i);
    if (x > -xi) return gk_n2_term(x, xi, t, zf, i) + sgn * gk_n2_term(-x, xi, t, zm, i);
    return sgn * (gk_n2_term(-x, xi, t, zm, i) - gk_n2_term(-x, xi, t, zj, i));
}

// which: 0 Hu 1 Hd 2 Hs 3 Hg 4 Hu_val 5 Hd_val 6 Hu_sea 7 Hd_sea (gk.f90:75-123); out[ix][ixi][it]
__global__ void gk_kernel(int which, int nx, int nxi, int nt, const double* __restrict__ x,
                          const double* __restrict__ xi, const double* __restrict__ t, double* __restrict__ out)
{
    const size_t total = (size_t)nx * nxi * nt;
    const double kappa_s = 1.68;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int it = (int)(e % nt), ixi = (int)((e / nt) % nxi), ix = (int)(e / ((size_t)nt * nxi));
        const double X = x[ix], XI = xi[ixi], T = t[it];
        double H;
        switch (which) {
        case 0: H = gk_n1(X, XI, T, 1) + kappa_s * gk_n2(X, XI, T, 3, -1.0); break;
        case 1: H = gk_n1(X, XI, T, 2) + kappa_s * gk_n2(X, XI, T, 3, -1.0); break;
        case 2: H = gk_n2(X, XI, T, 3, -1.0); break;
        case 3: H = gk_n2(X, XI, T, 0, 1.0); break;
        case 4: H = gk_n1(X, XI, T, 1); break;
        case 5: H = gk_n1(X, XI, T, 2); break;
        default: H = kappa_s * gk_n2(X, XI, T, 3, -1.0); break;
        }
        out[e] = H;
    }
}

int launch_gk(int which, int nx, int nxi, int nt, const double* x, const double* xi, const double* t, double* out,
              cudaStream_t st)
{
    const size_t total = (size_t)nx * nxi * nt;
    if (total == 0) return 0;
    int blocks = (int)((total + 127) / 128);
    if (blocks > 148 * 8) blocks = 148 * 8;
    gk_kernel<<<blocks, 128, 0, st>>>(which, nx, nxi, nt, x, xi, t, out);
    cudaError_t e = cudaGetLastError();
    ++g_launches;
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    return 0;
}

}  // namespace tkb
