/* tko_gk.c -- oracle restatement of the Goloskokov-Kroll model GPD (input generator).
 *
 * TEST INFRASTRUCTURE ONLY (see tko_fp.h).
 * Follows f90src/model/gk.f90:27-241 and the loop order of tiktaalik/f90wrap/model.f90:16-158.
 * Not on the hot path; used to build parity inputs (examples/examples.py:37-99).
 */
#include "tko.h"
#include "tko_fp.h"

/* gk.f90:30-56 */
static const double a0_val = 0.48, ap_val = 0.9, a0_sea = 1.1, ap_sea = 0.15, kappa_s = 1.68;
/* gk.f90:59-64: C(i,j), i = 0 gluon, 1 u_val, 2 d_val, 3 strange; j = power index */
static const double Cmat[4][4] = {
    { 2.23, 5.43, -34.0, 40.6 },
    { 1.52, 2.88, -0.095, 0.0 },
    { 0.76, 3.11, -3.99, 0.0 },
    { 0.123, -0.327, 0.692, -0.486 } };

/* gk.f90:233-239 */
static double pw(int i, int j, int w, double t)
{
    const double delta[4] = { a0_sea - 1.0, a0_val, a0_val, a0_sea };
    const double ap[4] = { ap_sea, ap_val, ap_val, ap_sea };
    return (double)((float)j / 2.0f + (float)w) - delta[i] - ap[i] * t;
}
/* gk.f90:193-209 */
static double kw(double x, double xi, int w)
{
    switch (w) {
    case 1: return 1.0 - (x * x) / (xi * xi);
    case 2: return 2.0 * (x / (xi * xi) - 1.0);
    case 3: return 1.0 - 1.0 / (xi * xi);
    }
    return 0.0;
}
/* gk.f90:211-231 */
static double ew(double x, double xi, int w)
{
    switch (w) {
    case 1: return kw(x, xi, 1) * kw(x, xi, 1);
    case 2: return 2.0 * kw(x, xi, 1) * kw(x, xi, 2);
    case 3: return kw(x, xi, 2) * kw(x, xi, 2) + 2.0 * kw(x, xi, 1) * kw(x, xi, 3);
    case 4: return 2.0 * kw(x, xi, 2) * kw(x, xi, 3);
    case 5: return kw(x, xi, 3) * kw(x, xi, 3);
    }
    return 0.0;
}
/* gk.f90:161-175 */
static double H_n1_term(double x, double xi, double t, double z, int i)
{
    double H = 0.0;
    for (int j = 0; j <= 3; j++)
        for (int w = 1; w <= 3; w++)
            H = H + Cmat[i][j] * kw(x, xi, w) * tko_pow(z, pw(i, j, w, t)) / pw(i, j, w, t);
    return H * 0.75 / xi;
}
/* gk.f90:177-191 */
static double H_n2_term(double x, double xi, double t, double z, int i)
{
    double H = 0.0;
    for (int j = 0; j <= 3; j++)
        for (int w = 1; w <= 5; w++)
            H = H + Cmat[i][j] * ew(x, xi, w) * tko_pow(z, pw(i, j, w, t)) / pw(i, j, w, t);
    return H * 1.875 / xi / 2.0;
}
/* gk.f90:125-139 */
static double H_n1(double x, double xi, double t, int i)
{
    double H = 0.0;
    double zi = (x - xi) / (1.0 - xi);
    double zf = (x + xi) / (1.0 + xi);
    if (x > xi) H = H_n1_term(x, xi, t, zf, i) - H_n1_term(x, xi, t, zi, i);
    else if (x > -xi) H = H_n1_term(x, xi, t, zf, i);
    return H;
}
/* gk.f90:141-159 */
static double H_n2(double x, double xi, double t, int i, double sgn)
{
    double H;
    double zi = (x - xi) / (1.0 - xi);
    double zf = (x + xi) / (1.0 + xi);
    double zj = (-x - xi) / (1.0 - xi);
    double zm = (-x + xi) / (1.0 + xi);
    if (x > xi) H = H_n2_term(x, xi, t, zf, i) - H_n2_term(x, xi, t, zi, i);
    else if (x > -xi) H = H_n2_term(x, xi, t, zf, i) + sgn * H_n2_term(-x, xi, t, zm, i);
    else H = sgn * (H_n2_term(-x, xi, t, zm, i) - H_n2_term(-x, xi, t, zj, i));
    return H;
}
/* gk.f90:75-123 */
double tko_gk(int which, double x, double xi, double t)
{
    switch (which) {
    case 0: return H_n1(x, xi, t, 1) + kappa_s * H_n2(x, xi, t, 3, -1.0);   /* Hu = Hu_val + Hu_sea */
    case 1: return H_n1(x, xi, t, 2) + kappa_s * H_n2(x, xi, t, 3, -1.0);   /* Hd */
    case 2: return H_n2(x, xi, t, 3, -1.0);                                 /* Hs */
    case 3: return H_n2(x, xi, t, 0, 1.0);                                  /* Hg */
    case 4: return H_n1(x, xi, t, 1);                                       /* Hu_val */
    case 5: return H_n1(x, xi, t, 2);                                       /* Hd_val */
    case 6: return kappa_s * H_n2(x, xi, t, 3, -1.0);                       /* Hu_sea */
    case 7: return kappa_s * H_n2(x, xi, t, 3, -1.0);                       /* Hd_sea */
    }
    return 0.0;
}
/* f90wrap/model.f90:16-32 etc.: f(ix,ixi,it), Fortran order */
void tko_gk_grid(int which, int nx, int nxi, int nt, const double *x, const double *xi, const double *t, double *f)
{
    for (int it = 0; it < nt; it++)
        for (int ixi = 0; ixi < nxi; ixi++)
            for (int ix = 0; ix < nx; ix++)
                f[(size_t)ix + (size_t)nx * ((size_t)ixi + (size_t)nxi * it)] = tko_gk(which, x[ix], xi[ixi], t[it]);
}
