// tkb_grid.cuh -- x-grid maps and the 6-point Lagrange finite-element basis (device code).
//
// Reference: gridspace.f90:23-102 (push_forward, pull_back, push_jacob),
// pixelation.f90:29-112 (interpixel, lagrange_basis, locate, weights),
// convolution.f90:45-46 (float32 node formula).  The float32 arithmetic of the reference
// (default-kind literals, real(int)) is reproduced with explicit round-to-nearest float
// intrinsics; it moves kernel entries at the 1e-4 level.
#pragma once
#include "tkb_fp.cuh"

namespace tkb {

// -Ofast ambiguity of the reference build: the float32 divisions by the pixel count may have been
// compiled as multiplications by a rounded reciprocal.  [0]: node formula (convolution.f90:45,67,92), [1]: stencil
// origin (pixelation.f90:85); 0 = IEEE division (canonical), 1 = reciprocal multiply.  Set by tkb_set_variant.
static __constant__ int c_variant[2] = { 0, 0 };
TKB_DEV float quot_f32(float a, float b, int recip)
{
    return recip ? __fmul_rn(a, __frcp_rn(b)) : __fdiv_rn(a, b);
}

TKB_DEV double push_forward(double eta, double xi, int grid_type)
{
    if (grid_type == 2) {
        if (eta < -0.5) return -pow_(xi * xi, 1.0 + eta);
        if (eta <= 0.5) return 2.0 * xi * eta;
        return pow_(xi * xi, 1.0 - eta);
    }
    return eta;
}

TKB_DEV double pull_back(double x, double xi, int grid_type)
{
    if (grid_type == 2) {
        if (x < -xi) return 0.5 * log_cr(-x / (xi * xi)) / log_cr(xi);
        if (x <= xi) return 0.5 * x / xi;
        return 0.5 * log_cr((xi * xi) / x) / log_cr(xi);
    }
    return x;
}

TKB_DEV double push_jacob(double eta, double xi, int grid_type)
{
    if (grid_type == 2) {
        if (eta < -0.5) return -pow_(xi * xi, 1.0 + eta) * 2.0 * log_cr(xi);
        if (eta <= 0.5) return 2.0 * xi;
        return -pow_(xi * xi, 1.0 - eta) * 2.0 * log_cr(xi);
    }
    return 1.0;
}

// x_i = push_forward( real(2*i-1)/real(n) - 1. ), the subtraction in float32 (convolution.f90:45)
TKB_DEV double grid_node(int i /*1-based*/, int n, double xi, int grid_type)
{
    const float etaf = __fsub_rn(quot_f32((float)(2 * i - 1), (float)n, c_variant[0]), 1.0f);
    return push_forward((double)etaf, xi, grid_type);
}

// Lagrange weights w(m) = prod_{l != m} 1/(dx*(m-l)), divided in index order (pixelation.f90:45-64)
TKB_DEV void lagrange_weights(int n, double* w /*[6]*/)
{
    const double dx = (double)__fdiv_rn(2.0f, (float)n);
    for (int i = 1; i <= 6; i++) {
        double wi = 1.0;
        for (int j = 1; j <= 6; j++)
            if (i != j) wi = wi / (dx * (double)(i - j));
        w[i - 1] = wi;
    }
}

// The six non-zero basis values at eta: f_j(eta) = b[j - iStart], j = iStart .. iStart+5 (1-based)
// (pixelation.f90:70-112).  lfact(m) is replaced by 1 before the in-order product.
TKB_DEV int basis_row(double eta, int n, const double* __restrict__ w, double* __restrict__ b)
{
    int iloc = (int)floor(0.5 * ((double)n * (eta + 1.0) + 1.0));
    if (iloc < 1) iloc = 1;
    if (iloc > n) iloc = n;
    int iStart = iloc - 2;
    if (iStart < 1) iStart = 1;
    if (iStart > n - 5) iStart = n - 5;
    const float x0f = __fadd_rn(-1.0f, quot_f32((float)(2 * iStart - 1), (float)n, c_variant[1]));
    const double x0 = (double)x0f;
    const double dx = (double)__fdiv_rn(2.0f, (float)n);
    double lf[6];
#pragma unroll
    for (int j = 0; j < 6; j++) lf[j] = eta - x0 - dx * (double)j;
#pragma unroll
    for (int m = 0; m < 6; m++) {
        double prod = 1.0;
#pragma unroll
        for (int j = 0; j < 6; j++) prod = prod * ((j == m) ? 1.0 : lf[j]);
        b[m] = w[m] * prod;
    }
    return iStart;
}

}  // namespace tkb
