/* tko_grid.c -- oracle restatement of the x-grid maps and the finite-element basis.
 *
 * TEST INFRASTRUCTURE ONLY (see tko_fp.h).
 * Follows f90src/helpers/gridspace.f90:23-102 (push_forward, pull_back, push_jacob),
 * f90src/helpers/pixelation.f90:29-112 (interpixel, Lagrange weights, locate),
 * tiktaalik/f90wrap/matrices.f90:21-30 (pixelspace_wrap) and
 * f90src/evolution/convolution.f90:45-46 (the float32 node formula).
 * Float32 arithmetic of the reference (default-kind literals, real(int)) is kept:
 * see SURVEY.md section 0.5 for why it matters at the 1e-4 level.
 */
#include "tko.h"
#include "tko_fp.h"

/* gridspace.f90:23-48 */
double tko_push_forward(double eta, double xi, int nx, int grid_type)
{
    (void)nx;
    double x;
    switch (grid_type) {
    case 2:
        if (eta < -0.5) x = -tko_pow(xi * xi, 1.0 + eta);
        else if (eta <= 0.5) x = 2.0 * xi * eta;
        else x = tko_pow(xi * xi, 1.0 - eta);
        break;
    case 1:
    default:
        x = eta;
    }
    return x;
}

/* gridspace.f90:50-75 */
double tko_pull_back(double x, double xi, int nx, int grid_type)
{
    (void)nx;
    double eta;
    switch (grid_type) {
    case 2:
        if (x < -xi) eta = 0.5 * tko_log_cr(-x / (xi * xi)) / tko_log_cr(xi);
        else if (x <= xi) eta = 0.5 * x / xi;
        else eta = 0.5 * tko_log_cr((xi * xi) / x) / tko_log_cr(xi);
        break;
    case 1:
    default:
        eta = x;
    }
    return eta;
}

/* gridspace.f90:77-102 */
double tko_push_jacob(double eta, double xi, int nx, int grid_type)
{
    (void)nx;
    double J;
    switch (grid_type) {
    case 2:
        if (eta < -0.5) J = -tko_pow(xi * xi, 1.0 + eta) * 2.0 * tko_log_cr(xi);
        else if (eta <= 0.5) J = 2.0 * xi;
        else J = -tko_pow(xi * xi, 1.0 - eta) * 2.0 * tko_log_cr(xi);
        break;
    case 1:
    default:
        J = 1.0;
    }
    return J;
}

/* pixelation.f90:19-21 module state */
static int N_order_cache = 0, Nx_cache = 0;
static double w_table[16];

/* pixelation.f90:45-64; dx = 2./real(nx) is a float32 quotient */
void tko_initialize_lagrange_weights(int nx, int norder)
{
    N_order_cache = norder;
    Nx_cache = nx;
    double dx = (double)(2.0f / (float)nx);
    for (int i = 1; i <= norder; i++) {
        w_table[i] = 1.0;
        for (int j = 1; j <= norder; j++)
            if (i != j) w_table[i] = w_table[i] / (dx * (double)(float)(i - j));
    }
}

/* pixelation.f90:104-112 */
static int locate(double x, int nx)
{
    int iloc = (int)floor(0.5 * ((double)(float)nx * (x + 1.0) + 1.0));
    if (iloc < 1) iloc = 1;
    if (iloc > nx) iloc = nx;
    return iloc;
}

/* pixelation.f90:89-102 */
static double lagrange_basis(double x, double x0, int i)
{
    double lfact[16];
    double dx = (double)(2.0f / (float)Nx_cache);
    for (int j = 1; j <= N_order_cache; j++)
        lfact[j] = x - x0 - dx * (double)(float)(j - 1);
    lfact[i] = 1.0;
    double prod = 1.0;                       /* product(lfact), in index order */
    for (int j = 1; j <= N_order_cache; j++) prod = prod * lfact[j];
    return w_table[i] * prod;
}

/* -Ofast ambiguity: the reference is built with -Ofast, which may turn the two float32
 * divisions by the loop-invariant pixel count into multiplications by a rounded reciprocal.  A 1-ulp float32
 * change of a node moves kernel entries by up to 2e-4, so the choice is exposed: strict IEEE division (0, the
 * canonical variant) or reciprocal multiply (1), separately for the node formula (convolution.f90:45,67,92)
 * and the stencil origin (pixelation.f90:85).  tkb_set_variant is the GPU side's switch. */
static int eta_recip_f32 = 0, x0_recip_f32 = 0;
void tko_set_variant(int eta_recip, int x0_recip) { eta_recip_f32 = eta_recip != 0; x0_recip_f32 = x0_recip != 0; }
static float quot_f32(float a, float b, int recip)
{
    if (!recip) return a / b;
    volatile float r = 1.0f / b;              /* rounded reciprocal, kept out of any re-association */
    return a * r;
}

/* pixelation.f90:70-87; x0 is evaluated entirely in float32 */
static double interpolate_lagrange_pixel(int ix, double x)
{
    int iLoc = locate(x, Nx_cache);
    int iStart = iLoc - (N_order_cache - 1) / 2;
    if (iStart < 1) iStart = 1;
    if (iStart > Nx_cache + 1 - N_order_cache) iStart = Nx_cache + 1 - N_order_cache;
    int iEnd = iStart + N_order_cache - 1;
    if (ix < iStart || ix > iEnd) return 0.0;
    float x0f = -1.0f + quot_f32((float)(2 * iStart - 1), (float)Nx_cache, x0_recip_f32);
    return lagrange_basis(x, (double)x0f, ix - iStart + 1);
}

/* pixelation.f90:29-40 */
double tko_interpixel(int n_pixels, int i, double x, double xi, int grid_type)
{
    double eta = tko_pull_back(x, xi, n_pixels, grid_type);
    return interpolate_lagrange_pixel(i, eta);
}

/* convolution.f90:45-46: eta = real(2*i-1)/real(n_pixels) - 1. (all float32), then push_forward */
double tko_grid_node(int i, int nx, double xi, int grid_type)
{
    float etaf = quot_f32((float)(2 * i - 1), (float)nx, eta_recip_f32) - 1.0f;
    return tko_push_forward((double)etaf, xi, nx, grid_type);
}

/* f90wrap/matrices.f90:21-30: float32 quotient, float64 subtraction */
void tko_pixelspace(int nx, double xi, int grid_type, double *x)
{
    for (int ix = 1; ix <= nx; ix++) {
        double eta = (double)((float)(2 * ix - 1) / (float)nx) - 1.0;
        x[ix - 1] = tko_push_forward(eta, xi, nx, grid_type);
    }
}

/* test hooks: the primitives of tko_fp.h as exported symbols */
double tko_log_x(double x) { return tko_log(x); }
double tko_log_cr_x(double x) { return tko_log_cr(x); }
double tko_pow_x(double x, double y) { return tko_pow(x, y); }
