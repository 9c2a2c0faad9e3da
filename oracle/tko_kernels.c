/* tko_kernels.c -- oracle restatement of the BFM splitting kernels.
 *
 * TEST INFRASTRUCTURE ONLY (see tko_fp.h).
 * Follows, expression by expression and in source evaluation order:
 *   f90src/helpers/specfun.f90:24-79          (dilog, dilog_half)
 *   f90src/kernels/kernels_common.f90:37-399  (steps, logs, f_a/f_b/f_c, h, hbar)
 *   f90src/kernels/kernels_lo.f90:48-221      (LO pieces)
 *   f90src/kernels/kernels_nlo.f90:45-982     (NLO two-variable pieces)
 * Fortran semantics kept on purpose: default-kind literals such as 11./3. are
 * float32 quotients (R32 below); x**2 is x*x; left-to-right evaluation of * and /;
 * theta(0) = 1/2; pieces are always evaluated and then multiplied by their step
 * function (so 0*Inf = NaN behaves as in the reference).
 */
#include "tko.h"
#include "tko_fp.h"

#define CF TKO_CF
#define CA TKO_CA
#define TF TKO_TF
#define R32 TKO_R32
#define SQ(a) ((a) * (a))

/* constants.f90:21-24: pi = acos(-1), zeta2 = pi**2/6, zeta3 = 1.2020569031 */
const double tko_pi    = 3.141592653589793;
const double tko_zeta2 = (3.141592653589793 * 3.141592653589793) / 6.0;
const double tko_zeta3 = 1.2020569031;
#define PI    3.141592653589793
#define PI2   (PI * PI)
#define ZETA2 (PI2 / 6.0)
#define ZETA3 1.2020569031

/* ------------------------------------------------------------------ specfun */

/* specfun.f90:53-79 */
static double dilog_half(double x)
{
    static const double p[7] = {
        0.9999999999999999502e+0, -2.6883926818565423430e+0, 2.6477222699473109692e+0,
        -1.1538559607887416355e+0, 2.0886077795020607837e-1, -1.0859777134152463084e-2, 0.0 };
    static const double q[7] = {
        1.0000000000000000000e+0, -2.9383926818565635485e+0, 3.2712093293018635389e+0,
        -1.7076702173954289421e+0, 4.1596017228400603836e-1, -3.9801343754084482956e-2,
        8.2743668974466659035e-4 };
    double xpow[7];
    xpow[0] = 1.0;
    xpow[1] = x;
    xpow[2] = x * x;
    xpow[3] = x * xpow[2];
    xpow[4] = xpow[2] * xpow[2];
    xpow[5] = xpow[4] * x;
    xpow[6] = xpow[3] * xpow[3];
    double px = 0.0, qx = 0.0;
    for (int i = 0; i < 7; i++) px = px + xpow[i] * p[i];
    for (int i = 0; i < 7; i++) qx = qx + xpow[i] * q[i];
    return x * px / qx;
}

/* specfun.f90:24-51.  `log(2.)` at :33 is a float32 log of a float32 literal. */
double tko_dilog(double z)
{
    if (z == -1.0) return -PI2 / 12.0;
    if (z == 0.0) return 0.0;
    if (z == 0.5) {
        const float l2 = 0.693147182464599609375f;   /* correctly rounded logf(2.0f) */
        const float h = (l2 * l2) / 2.0f;
        return PI2 / 12.0 - (double)h;
    }
    if (z == 1.0) return PI2 / 6.0;
    if (z == 2.0) return PI2 / 4.0;
    if (z < -1.0) {
        double l = tko_log(1.0 - z);
        return l * (0.5 * l - tko_log(-z)) - PI2 / 6.0 + tko_dilog(1.0 / (1.0 - z));
    }
    if (z < 0.0) {
        double l = tko_log(1.0 - z);
        return -tko_dilog(z / (z - 1.0)) - 0.5 * SQ(l);
    }
    if (z > 2.0) {
        double l = tko_log(z);
        return -tko_dilog(1.0 / z) + PI2 / 3.0 - 0.5 * SQ(l);
    }
    if (z > 1.0) {
        double l = tko_log(z);
        return PI2 / 6.0 - l * (tko_log(1.0 - 1.0 / z) + 0.5 * l) + tko_dilog(1.0 - 1.0 / z);
    }
    if (z > 0.5)
        return -tko_dilog(1.0 - z) + PI2 / 6.0 - tko_log(z) * tko_log(1.0 - z);
    return dilog_half(z);     /* also reached by NaN, as in the reference's final else */
}

/* ----------------------------------------------------------- kernels_common */

/* kernels_common.f90:37-46 */
double tko_theta_step(double X)
{
    double theta = 0.0;
    if (X > 0.0) theta = 1.0;
    else if (X == 0.0) theta = 0.5;
    return theta;
}
/* kernels_common.f90:48-53; sign(1.0_dp,Y) is copysign */
double tko_rho_step(double X, double Y)
{
    return tko_theta_step(1.0 - X / Y) * tko_theta_step(X / Y) * copysign(1.0, Y);
}
/* kernels_common.f90:55-61 */
double tko_erbl_step(double X) { return tko_theta_step(X) * tko_theta_step(1.0 - X); }
/* kernels_common.f90:66-73 */
double tko_abslog(double x) { return tko_log(fabs(x)); }
/* kernels_common.f90:75-83 */
double tko_log2(double x)
{
    double l = tko_log(fabs(x));
    double y = l * l;
    if (x < 0.0) y = y - PI2;
    return y;
}
/* kernels_common.f90:85-93 */
double tko_logprod(double x1, double x2)
{
    double y = tko_log(fabs(x1)) * tko_log(fabs(x2));
    if (x1 < 0.0 && x2 < 0.0) y = y - PI2;
    return y;
}

#define theta_step tko_theta_step
#define rho_step tko_rho_step
#define erbl_step tko_erbl_step
#define abslog tko_abslog
#define log2_ tko_log2
#define logprod tko_logprod
#define dilog tko_dilog

/* kernels_common.f90:100-138 */
static double QQ_f_a(double X, double Y) { return X / Y; }
static double QQ_f_b(double X, double Y) { return X / Y / (Y - X); }
/* kernels_common.f90:125-148 (QQ_h, QQ_hbar) */
static double QQ_h(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double QQf = QQ_f_a(X, Y) + QQ_f_b(X, Y);
    double QQfbar = QQ_f_a(Xbar, Ybar) + QQ_f_b(Xbar, Ybar);
    return 2.0 * QQfbar * logprod(Xbar, Y) - 2.0 * QQf * (dilog(X) + dilog(Ybar));
}
static double QQ_hbar(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double QQf = QQ_f_a(X, Y) + QQ_f_b(X, Y);
    double QQfbar = QQ_f_a(Xbar, Ybar) + QQ_f_b(Xbar, Ybar);
    return (-QQfbar + QQf) * (2.0 * dilog(1.0 - X / Y) + log2_(Y))
         + 2.0 * QQf * (dilog(Ybar) - logprod(X, Y))
         + 2.0 * QQfbar * dilog(Xbar);
}

/* kernels_common.f90:152-171 */
static double QG_f_a(double X, double Y) { return X / SQ(Y); }
static double QG_f_c(double X, double Y)
{
    double Xbar = 1.0 - X;
    return X / SQ(Y) * (2.0 * Xbar * Y - X);
}
/* kernels_common.f90:173-234 */
static double QG_h_V(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = -QG_f_a(X, Y) - 2.0 * QG_f_c(X, Y);
    double fbar = -QG_f_a(Xbar, Ybar) - 2.0 * QG_f_c(Xbar, Ybar);
    return -2.0 * fbar * logprod(Xbar, Y) - 2.0 * f * (dilog(X) + dilog(Ybar));
}
static double QG_hbar_V(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = -QG_f_a(X, Y) - 2.0 * QG_f_c(X, Y);
    double fbar = -QG_f_a(Xbar, Ybar) - 2.0 * QG_f_c(Xbar, Ybar);
    return (fbar + f) * (2.0 * dilog(1.0 - X / Y) + log2_(Y))
         + 2.0 * f * (dilog(Ybar) - logprod(X, Y))
         - 2.0 * fbar * dilog(Xbar);
}
static double QG_h_A(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = -QG_f_a(X, Y);
    double fbar = -QG_f_a(Xbar, Ybar);
    return -2.0 * fbar * logprod(Xbar, Y) - 2.0 * f * (dilog(X) + dilog(Ybar));
}
static double QG_hbar_A(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = -QG_f_a(X, Y);
    double fbar = -QG_f_a(Xbar, Ybar);
    return (fbar + f) * (2.0 * dilog(1.0 - X / Y) + log2_(Y))
         + 2.0 * f * (dilog(Ybar) - logprod(X, Y))
         - 2.0 * fbar * dilog(Xbar);
}

/* kernels_common.f90:238-258 */
static double GQ_f_a(double X, double Y) { return SQ(X) / Y; }
static double GQ_f_c(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    return SQ(X) / Y * (2.0 * Xbar * Y - Ybar);
}
/* kernels_common.f90:260-321 */
static double GQ_h_V(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = GQ_f_a(X, Y) + 2.0 * GQ_f_c(X, Y);
    double fbar = GQ_f_a(Xbar, Ybar) + 2.0 * GQ_f_c(Xbar, Ybar);
    return -2.0 * fbar * logprod(Xbar, Y) - 2.0 * f * (dilog(X) + dilog(Ybar));
}
static double GQ_hbar_V(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = GQ_f_a(X, Y) + 2.0 * GQ_f_c(X, Y);
    double fbar = GQ_f_a(Xbar, Ybar) + 2.0 * GQ_f_c(Xbar, Ybar);
    return (fbar + f) * (2.0 * dilog(1.0 - X / Y) + log2_(Y))
         + 2.0 * f * (dilog(Ybar) - logprod(X, Y))
         - 2.0 * fbar * dilog(Xbar);
}
static double GQ_h_A(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = GQ_f_a(X, Y);
    double fbar = GQ_f_a(Xbar, Ybar);
    return -2.0 * fbar * logprod(Xbar, Y) - 2.0 * f * (dilog(X) + dilog(Ybar));
}
static double GQ_hbar_A(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = GQ_f_a(X, Y);
    double fbar = GQ_f_a(Xbar, Ybar);
    return (fbar + f) * (2.0 * dilog(1.0 - X / Y) + log2_(Y))
         + 2.0 * f * (dilog(Ybar) - logprod(X, Y))
         - 2.0 * fbar * dilog(Xbar);
}

/* kernels_common.f90:325-341 */
static double GG_f_a(double X, double Y) { return SQ(X) / SQ(Y); }
static double GG_f_b(double X, double Y) { return SQ(X) / SQ(Y) / (Y - X); }
static double GG_f_c(double X, double Y)
{
    double Xbar = 1.0 - X;
    return SQ(X) / SQ(Y) * (2.0 * Xbar * Y + Y - X);
}
/* kernels_common.f90:343-399 */
static double GG_h_V(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y) + 2.0 * GG_f_c(X, Y);
    double fbar = 2.0 * GG_f_a(Xbar, Ybar) + GG_f_b(Xbar, Ybar) + 2.0 * GG_f_c(Xbar, Ybar);
    return 2.0 * fbar * logprod(Xbar, Y) - 2.0 * f * (dilog(X) + dilog(Ybar));
}
static double GG_hbar_V(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y) + 2.0 * GG_f_c(X, Y);
    double fbar = 2.0 * GG_f_a(Xbar, Ybar) + GG_f_b(Xbar, Ybar) + 2.0 * GG_f_c(Xbar, Ybar);
    return (-fbar + f) * (2.0 * dilog(1.0 - X / Y) + log2_(Y))
         + 2.0 * f * (dilog(Ybar) - logprod(X, Y))
         + 2.0 * fbar * dilog(Xbar);
}
static double GG_h_A(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y);
    double fbar = 2.0 * GG_f_a(Xbar, Ybar) + GG_f_b(Xbar, Ybar);
    return 2.0 * fbar * logprod(Xbar, Y) - 2.0 * f * (dilog(X) + dilog(Ybar));
}
static double GG_hbar_A(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double f = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y);
    double fbar = 2.0 * GG_f_a(Xbar, Ybar) + GG_f_b(Xbar, Ybar);
    return (-fbar + f) * (2.0 * dilog(1.0 - X / Y) + log2_(Y))
         + 2.0 * f * (dilog(Ybar) - logprod(X, Y))
         + 2.0 * fbar * dilog(Xbar);
}

/* ---------------------------------------------------------------- kernels_lo */

#define XY_VARS                                   \
    double X1 = 0.5 * (1.0 + x / xi);             \
    double X2 = 0.5 * (1.0 - x / xi);             \
    double Y1 = 0.5 * (1.0 + y / xi);             \
    double Y2 = 0.5 * (1.0 - y / xi)

/* kernels_lo.f90:48-62 */
double tko_K0_qq_pls(double x, double y, double xi)
{
    XY_VARS;
    double K = (QQ_f_a(X1, Y1) + QQ_f_b(X1, Y1)) * rho_step(X1, Y1)
             + (QQ_f_a(X2, Y2) + QQ_f_b(X2, Y2)) * rho_step(X2, Y2);
    return CF * 0.5 * K / xi;
}

/* kernels_lo.f90:64-79 */
double tko_K0_qq_cst(double x, double xi)
{
    double K;
    if (x > xi)
        K = 1.5 + 2.0 * tko_log(1.0 - x)
          + 0.5 * ((x - xi) * tko_log((x - xi) * (1.0 + xi)) - (x + xi) * tko_log((x + xi) * (1.0 - xi))) / xi;
    else if (x < xi && x > -xi)
        K = 1.5 + tko_log((1.0 - SQ(x)) / (1.0 + xi))
          + 0.5 * ((x - xi) * tko_log(xi - x) - (x + xi) * tko_log(x + xi)) / xi;
    else if (x < -xi)
        K = 1.5 + 2.0 * tko_log(1.0 + x)
          + 0.5 * ((-x - xi) * tko_log((-x - xi) * (1.0 + xi)) - (-x + xi) * tko_log((-x + xi) * (1.0 - xi))) / xi;
    else
        K = 0.0;
    return K * CF;
}

/* kernels_lo.f90:84-100 */
double tko_KA0_qG_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = -QG_f_a(X1, Y1) * rho_step(X1, Y1) + QG_f_a(X2, Y2) * rho_step(X2, Y2);
    K = 2.0 * TF * 0.5 * K / xi;
    K = K / (2.0 * xi);
    return K;
}
/* kernels_lo.f90:102-118 */
double tko_KVmA0_qG_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = -2.0 * QG_f_c(X1, Y1) * rho_step(X1, Y1) + 2.0 * QG_f_c(X2, Y2) * rho_step(X2, Y2);
    K = 2.0 * TF * 0.5 * K / xi;
    K = K / (2.0 * xi);
    return K;
}
/* kernels_lo.f90:123-136 */
double tko_KA0_Gq_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = GQ_f_a(X1, Y1) * rho_step(X1, Y1) - GQ_f_a(X2, Y2) * rho_step(X2, Y2);
    K = CF * 0.5 * K / xi;
    K = K * (2.0 * xi);
    return K;
}
/* kernels_lo.f90:138-151 */
double tko_KVmA0_Gq_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = 2.0 * GQ_f_c(X1, Y1) * rho_step(X1, Y1) - 2.0 * GQ_f_c(X2, Y2) * rho_step(X2, Y2);
    K = CF * 0.5 * K / xi;
    K = K * (2.0 * xi);
    return K;
}
/* kernels_lo.f90:156-169 */
double tko_KA0_GG_pls(double x, double y, double xi)
{
    XY_VARS;
    double K = (2.0 * GG_f_a(X1, Y1) + GG_f_b(X1, Y1)) * rho_step(X1, Y1)
             + (2.0 * GG_f_a(X2, Y2) + GG_f_b(X2, Y2)) * rho_step(X2, Y2);
    return CA * 0.5 * K / xi;
}
/* kernels_lo.f90:171-195; `11./3.` is a float32 quotient */
double tko_KA0_GG_cst(double x, double xi)
{
    double K = 0.0;
    if (x > xi)
        K = -2.0 * x * (1.0 - x) / (1.0 - SQ(xi)) + tko_log(SQ(1.0 - x) / (1.0 - SQ(xi)));
    else if (x < -xi)
        K = 2.0 * x * (1.0 + x) / (1.0 - SQ(xi)) + tko_log(SQ(1.0 + x) / (1.0 - SQ(xi)));
    else
        K = -2.0 * SQ(x) / (1.0 + xi) / xi + tko_log((1.0 - SQ(x)) / SQ(1.0 + xi));
    K = K * CA;
    double beta0 = -R32(11, 3) * CA;
    K = K - beta0 / 2.0;
    return K;
}
/* kernels_lo.f90:197-207 */
double tko_KA0_GG_nfl(double x, double xi)
{
    (void)x; (void)xi;
    double beta0 = R32(4, 3) * TF;
    return -beta0 / 2.0;
}
/* kernels_lo.f90:209-221 */
double tko_KVmA0_GG_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = 2.0 * GG_f_c(X1, Y1) * rho_step(X1, Y1) + 2.0 * GG_f_c(X2, Y2) * rho_step(X2, Y2);
    return CA * 0.5 * K / xi;
}

/* --------------------------------------------------------------- kernels_nlo */

/* kernels_nlo.f90:113-142 */
static double KV1_qq_half(double X, double Y, double Xbar, double Ybar, double zsign)
{
    double QQf = QQ_f_a(X, Y) + QQ_f_b(X, Y);
    double QQfbar = QQ_f_a(Xbar, Ybar) + QQ_f_b(Xbar, Ybar);
    double QQh = QQ_h(X, Y);
    double QQhbar = QQ_hbar(Xbar, Y);
    double beta0 = -R32(11, 3) * CA;
    double piece1 = SQ(CF) * ((R32(4, 3) - 2.0 * ZETA2) * QQf + 3.0 * X / Y
        - (1.5 * QQf - 0.5 * X / Ybar) * abslog(X / Y)
        - (QQf - QQfbar) * logprod(X / Y, 1.0 - X / Y)
        + (QQf + 0.5 * X / Ybar) * log2_(X / Y));
    double piece2 = -0.5 * CF * beta0 * (R32(5, 3) * QQf + X / Y + QQf * abslog(X / Y));
    double piece3 = -CF * (CF - 0.5 * CA) * (R32(4, 3) * QQf + 2.0 * X / Y + QQh - zsign * QQhbar);
    double piece4 = -0.5 * X / Ybar * (abslog(X) + log2_(X) - 2.0 * logprod(X, Xbar)) * SQ(CF);
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * rho_step(X, Y);
    piece3 = piece3 * rho_step(X, Y);
    piece4 = piece4 * erbl_step(X);
    return piece1 + piece2 + piece3 + piece4;
}
/* kernels_nlo.f90:45-58 */
double tko_KV1_NSp_pls(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_qq_half(X1, Y1, X2, Y2, 1.0) + KV1_qq_half(X2, Y2, X1, Y1, 1.0);
    return 0.5 * K / xi;
}
/* kernels_nlo.f90:80-93 */
double tko_KV1_NSm_pls(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_qq_half(X1, Y1, X2, Y2, -1.0) + KV1_qq_half(X2, Y2, X1, Y1, -1.0);
    return 0.5 * K / xi;
}

/* kernels_nlo.f90:213-228 (QQh, QQhbar are computed but unused there; zsign unused) */
static double KV1_qq_half_nfl(double X, double Y, double Xbar, double Ybar, double zsign)
{
    (void)Xbar; (void)Ybar; (void)zsign;
    double QQf = QQ_f_a(X, Y) + QQ_f_b(X, Y);
    double beta0 = R32(4, 3) * TF;
    double K = -0.5 * CF * beta0 * (R32(5, 3) * QQf + X / Y + QQf * abslog(X / Y));
    return K * rho_step(X, Y);
}
/* kernels_nlo.f90:147-160 */
double tko_KV1_NSp_pls_nfl(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_qq_half_nfl(X1, Y1, X2, Y2, 1.0) + KV1_qq_half_nfl(X2, Y2, X1, Y1, 1.0);
    return 0.5 * K / xi;
}
/* kernels_nlo.f90:180-193 */
double tko_KV1_NSm_pls_nfl(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_qq_half_nfl(X1, Y1, X2, Y2, -1.0) + KV1_qq_half_nfl(X2, Y2, X1, Y1, -1.0);
    return 0.5 * K / xi;
}

/* kernels_nlo.f90:244-268 */
static double KV1_qq_half_sin(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double piece1 = X * (3.0 - 8.0 * X * Ybar) / Y + X * (5.0 - 8.0 * X) / Ybar * abslog(X / Y)
        + (X / Ybar - 4.0 * X * Xbar) * log2_(X / Y)
        + 8.0 * X * Xbar * (dilog(Xbar) - dilog(Ybar) + logprod(Xbar, Y));
    double piece2 = -R32(290, 9) * X * Xbar
        - (X * (5.0 - 8.0 * X) / Ybar + 2.0 * X * (9.0 - 19.0 * Xbar) / 3.0) * abslog(X)
        - (X / Ybar - 4.0 * X * Xbar) * log2_(X)
        + 4.0 * X * Xbar * logprod(X, Xbar);
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * erbl_step(X);
    return 2.0 * CF * TF * (piece1 + piece2);
}
/* kernels_nlo.f90:230-242 */
double tko_KV1_qq_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_qq_half_sin(X1, Y1) + KV1_qq_half_sin(X2, Y2);
    return 0.5 * K / xi;
}
/* kernels_nlo.f90:284-301 */
static double KA1_qq_half_sin(double X, double Y)
{
    double Ybar = 1.0 - Y;
    double piece1 = X / Ybar * (abslog(X / Y) - log2_(X / Y));
    double piece2 = -X / Ybar * (abslog(X) - log2_(X));
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * erbl_step(X);
    return -2.0 * CF * TF * (piece1 + piece2);
}
/* kernels_nlo.f90:270-282 */
double tko_KA1_qq_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = KA1_qq_half_sin(X1, Y1) + KA1_qq_half_sin(X2, Y2);
    return 0.5 * K / xi;
}

/* kernels_nlo.f90:321-368 */
static double KV1_qG_half(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double QGfV = -QG_f_a(X, Y) - 2.0 * QG_f_c(X, Y);
    double QGfA = -QG_f_a(X, Y);
    double QGfbarV = -QG_f_a(Xbar, Ybar) - 2.0 * QG_f_c(Xbar, Ybar);
    double QGfbarA = -QG_f_a(Xbar, Ybar);
    double QGhV = QG_h_V(X, Y);
    double QGhbarV = QG_hbar_V(Xbar, Y);
    double piece1 = TF * CF * (
        2.0 * (5.0 - 2.0 * ZETA2) * QGfV - 4.0 * QGfA + X / (Y * Ybar)
        + 2.0 * (QGfV + QGfbarV + (2.0 * X - 2.0 * Y + X * Y) / (2.0 * Y * SQ(Ybar))) * abslog(X / Y)
        - 2.0 * (QGfV + QGfbarV + 1.0 / (Y * Ybar)) * abslog(1.0 - X / Y)
        + QGfbarA * log2_(X / Y) + (QGfV + QGfbarV) * log2_(Y / X - 1.0));
    double al = abslog(1.0 - X / Y);
    double piece2 = TF * CA * (
        6.0 * X * (3.0 - 4.0 * X) / Ybar - 2.0 * X * (11.0 - 16.0 * X) / Y + 4.0 * X * (1.0 - 3.0 * X) / SQ(Y)
        - 2.0 * (QGfV - 5.0 * QGfbarV + (5.0 - 7.0 * X) / SQ(Ybar)
                 + 2.0 * X * (3.0 - 2.0 * X - 12.0 * Xbar * Y) / (Y * Ybar)) * abslog(X / Y)
        + 2.0 * (QGfV + QGfbarV + 1.0 / (Y * Ybar)) * abslog(1.0 - X / Y)
        + (QGfV - QGfbarV + (1.0 - 4.0 * X) / SQ(Ybar)) * log2_(X / Y)
        - (QGfV + QGfbarV) * SQ(al) - QGhV + QGhbarV);
    double piece3 = TF * CF * (
        -2.0 * (QGfV + QGfbarV + QGfbarA + X * (Y + 4.0 * SQ(Ybar)) / (2.0 * Y * SQ(Ybar))) * abslog(X)
        - 2.0 * QGfV * logprod(Xbar, X)
        + (QGfV - QGfbarV - QGfbarA) * log2_(X));
    double piece4 = TF * CA * (
        2.0 * (QGfV - 5.0 * QGfbarV + (5.0 - 7.0 * X) / SQ(Ybar)
               + X * (5.0 - 6.0 * X - 22.0 * Y + 28.0 * X * Y) / (Y * Ybar)) * abslog(X)
        - (QGfV - QGfbarV + (1.0 - 4.0 * X) / SQ(Ybar)) * log2_(X));
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * rho_step(X, Y);
    piece3 = piece3 * erbl_step(X);
    piece4 = piece4 * erbl_step(X);
    return piece1 + piece2 + piece3 + piece4;
}
/* kernels_nlo.f90:306-319 */
double tko_KV1_qG_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_qG_half(X1, Y1) - KV1_qG_half(X2, Y2);
    K = 0.5 * K / xi;
    K = K / (2.0 * xi);
    return K;
}
/* kernels_nlo.f90:385-432 */
static double KA1_qG_half(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double QGfA = -QG_f_a(X, Y);
    double QGfbarA = -QG_f_a(Xbar, Ybar);
    double QGhA = QG_h_A(X, Y);
    double QGhbarA = QG_hbar_A(Xbar, Y);
    double piece1 = TF * CF * (
        2.0 * (5.0 - 2.0 * ZETA2) * QGfA + 5.0 * X / (Y * Ybar)
        + (2.0 * QGfA - 3 * X / SQ(Ybar) + 2.0 * X / (Y * Ybar)) * abslog(X / Y)
        - 2.0 * (QGfA + QGfbarA + 1.0 / (Y * Ybar)) * abslog(1.0 - X / Y)
        - QGfbarA * log2_(X / Y) + (QGfA + QGfbarA) * log2_(Y / X - 1.0));
    double piece2 = TF * CA * (
        -4.0 * QGfA - 6.0 * X / (Y * Ybar) - 2.0 * (QGfA - 2.0 * X / SQ(Ybar)) * abslog(X / Y)
        + 2.0 * (QGfA + QGfbarA + 1.0 / (Y * Ybar)) * abslog(1.0 - X / Y)
        + (QGfA - 3.0 * X / SQ(Ybar)) * log2_(X / Y) - (QGfA + QGfbarA) * log2_(1.0 - X / Y)
        - QGhA - QGhbarA);
    double piece3 = TF * CF * (
        -QGfA * (2.0 * abslog(X) + 2.0 * logprod(X, Xbar) - log2_(X))
        + 3.0 * X / SQ(Ybar) * abslog(X));
    double piece4 = TF * CA * (
        (6.0 * X * Ybar - 4.0 * X * Y) / (Y * SQ(Ybar)) * abslog(X)
        + QGfA * (2.0 * abslog(X) - log2_(X)) + 3.0 * X / SQ(Ybar) * log2_(X));
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * rho_step(X, Y);
    piece3 = piece3 * erbl_step(X);
    piece4 = piece4 * erbl_step(X);
    return piece1 + piece2 + piece3 + piece4;
}
/* kernels_nlo.f90:370-383 */
double tko_KA1_qG_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = KA1_qG_half(X1, Y1) - KA1_qG_half(X2, Y2);
    K = 0.5 * K / xi;
    K = K / (2.0 * xi);
    return K;
}

/* kernels_nlo.f90:451-511; `abslog(x)` at :482 is abslog(X) (case-insensitive) */
static double KV1_Gq_half(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = -R32(11, 3) * CA;
    double GQfV = GQ_f_a(X, Y) + 2.0 * GQ_f_c(X, Y);
    double GQfA = GQ_f_a(X, Y);
    double GQfbarV = GQ_f_a(Xbar, Ybar) + 2.0 * GQ_f_c(Xbar, Ybar);
    double GQfbarA = GQ_f_a(Xbar, Ybar);
    double GQhV = GQ_h_V(X, Y);
    double GQhbarV = GQ_hbar_V(Xbar, Y);
    double piece1 = SQ(CF) * (
        3.0 * GQfA - 4.5 * X / Y - (0.5 * X * (2.0 + X) / Ybar - X * Xbar / (Y * Ybar)) * abslog(X / Y)
        - (1.5 * GQfV + 1.5 * GQfbarV + X * Xbar / (Y * Ybar)) * abslog(1.0 - X / Y)
        - 0.5 * GQfA * log2_(X / Y)
        - 0.5 * (GQfV + GQfbarV) * log2_(1.0 - X / Y));
    double piece2 = -0.5 * CF * beta0 * (
        R32(10, 3) * GQfV + 2.0 * X * Xbar / Y + (GQfV + GQfbarV) * abslog(1.0 - X / Y)
        - GQfA * abslog(X) + GQfbarA * abslog(Xbar));
    double piece3 = CF * CA * (
        (R32(4, 9) - 2.0 * ZETA2) * GQfV - R32(26, 9) * SQ(X) / Y + X * (3.0 + 4.0 * SQ(X)) / Y
        - X / 9.0 * (105.0 - 246.0 * X + 188.0 * SQ(X))
        - (X * Xbar / (Y * Ybar) - X * (6.0 - 11.0 * X + 8.0 * SQ(X)) / Ybar + 4.0 * X * Xbar * (Xbar - X)) * abslog(X / Y)
        + X * Xbar / (Y * Ybar) * abslog(1.0 - X / Y)
        + 0.5 * (2.0 * GQfV + 2.0 * GQfA + GQfbarA - 3.0 * SQ(X) / Ybar - 2.0) * log2_(X / Y)
        + 0.5 * (2.0 - GQfA - GQfbarA) * log2_(Y / X - 1.0)
        - 0.5 * GQhV - 0.5 * GQhbarV);
    double piece4 = SQ(CF) * (
        -0.5 * (X * (2.0 - 5.0 * X) / Y - 3.0 * SQ(X) / Ybar) * abslog(X)
        + 0.5 * GQfA * log2_(X));
    double piece5 = CF * CA * (
        (X * Xbar / (Y * Ybar) - X * (6.0 - 11.0 * X + 8.0 * SQ(X)) / Ybar
         + X * (18.0 - 63.0 * X + 62.0 * SQ(X)) / 3.0) * abslog(X)
        - GQfV * logprod(X, Xbar) - SQ(X) * (1.0 - 4.0 * Y) / (2.0 * Y * Ybar) * log2_(X));
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * rho_step(X, Y);
    piece3 = piece3 * rho_step(X, Y);
    piece4 = piece4 * erbl_step(X);
    piece5 = piece5 * erbl_step(X);
    return piece1 + piece2 + piece3 + piece4 + piece5;
}
/* kernels_nlo.f90:437-449 */
double tko_KV1_Gq_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_Gq_half(X1, Y1) - KV1_Gq_half(X2, Y2);
    K = 0.5 * K / xi;
    K = K * (2.0 * xi);
    return K;
}
/* kernels_nlo.f90:527-591; `GqfbarA` at :550 is GQfbarA */
static double KA1_Gq_half(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = -R32(11, 3) * CA;
    double GQfA = GQ_f_a(X, Y);
    double GQfbarA = GQ_f_a(Xbar, Ybar);
    double GQhA = GQ_h_A(X, Y);
    double GQhbarA = GQ_hbar_A(Xbar, Y);
    double piece1 = SQ(CF) * (
        -3.0 * X * (1.0 + X) / Y + (SQ(X) * Y + 2.0 * X * (Xbar - X) * Ybar) / (Y * Ybar) * abslog(X / Y)
        - (X * (Xbar + Ybar) + Xbar * Y) / (Y * Ybar) * abslog(1.0 - X / Y)
        + GQfA * log2_(X / Y) - (GQfA + GQfbarA) * log2_(1.0 - X / Y));
    double piece2 = -0.5 * CF * beta0 * (
        -R32(2, 3) * GQfA + 2.0 * X / Y + (GQfA + GQfbarA) * abslog(1.0 - X / Y));
    double piece3 = CF * CA * (
        (R32(13, 3) - 2.0 * ZETA2) * GQfA - X / Y
        - (X * (Xbar - X) / Y + 5.0 * X * Xbar / Ybar) * abslog(X / Y)
        + ((Xbar - 2.0 * SQ(Xbar)) / Ybar + (X - 2 * SQ(X)) / Y) * abslog(1.0 - X / Y)
        - 0.5 * (GQfbarA + 3.0 * SQ(X) / Ybar) * log2_(X / Y) + 0.5 * (GQfA + GQfbarA) * log2_(Y / X - 1.0)
        - 0.5 * GQhA + 0.5 * GQhbarA);
    double piece4 = SQ(CF) * (
        -(SQ(X) * Y + (1.0 - SQ(Xbar)) * Ybar) / (Y * Ybar) * abslog(X)
        - GQfA * log2_(X));
    double piece5 = -0.5 * CF * beta0 * (GQfA * abslog(X));
    double piece6 = CF * CA * (
        (5.0 * X * Xbar / Ybar + X / Y) * abslog(X)
        + GQfA * (-2.0 * abslog(X) + 0.5 * log2_(X) - logprod(X, Xbar))
        + 1.5 * SQ(X) / Ybar * log2_(X));
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * rho_step(X, Y);
    piece3 = piece3 * rho_step(X, Y);
    piece4 = piece4 * erbl_step(X);
    piece5 = piece5 * erbl_step(X);
    piece6 = piece6 * erbl_step(X);
    return piece1 + piece2 + piece3 + piece4 + piece5 + piece6;
}
/* kernels_nlo.f90:513-525 */
double tko_KA1_Gq_reg(double x, double y, double xi)
{
    XY_VARS;
    double K = KA1_Gq_half(X1, Y1) - KA1_Gq_half(X2, Y2);
    K = 0.5 * K / xi;
    K = K * (2.0 * xi);
    return K;
}
/* kernels_nlo.f90:610-635 */
static double KV1_Gq_half_nfl(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = R32(4, 3) * TF;
    double GQfV = GQ_f_a(X, Y) + 2.0 * GQ_f_c(X, Y);
    double GQfA = GQ_f_a(X, Y);
    double GQfbarV = GQ_f_a(Xbar, Ybar) + 2.0 * GQ_f_c(Xbar, Ybar);
    double GQfbarA = GQ_f_a(Xbar, Ybar);
    double K = -0.5 * CF * beta0 * (
        R32(10, 3) * GQfV + 2.0 * X * Xbar / Y + (GQfV + GQfbarV) * abslog(1.0 - X / Y)
        - GQfA * abslog(X) + GQfbarA * abslog(Xbar));
    return K * rho_step(X, Y);
}
/* kernels_nlo.f90:596-608 */
double tko_KV1_Gq_reg_nfl(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_Gq_half_nfl(X1, Y1) - KV1_Gq_half_nfl(X2, Y2);
    K = 0.5 * K / xi;
    K = K * (2.0 * xi);
    return K;
}
/* kernels_nlo.f90:651-671 */
static double KA1_Gq_half_nfl(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = R32(4, 3) * TF;
    double GQfA = GQ_f_a(X, Y);
    double GQfbarA = GQ_f_a(Xbar, Ybar);
    double piece2 = -0.5 * CF * beta0 * (
        -R32(2, 3) * GQfA + 2.0 * X / Y + (GQfA + GQfbarA) * abslog(1.0 - X / Y));
    double piece5 = -0.5 * CF * beta0 * (GQfA * abslog(X));
    piece2 = piece2 * rho_step(X, Y);
    piece5 = piece5 * erbl_step(X);
    return piece2 + piece5;
}
/* kernels_nlo.f90:637-649 */
double tko_KA1_Gq_reg_nfl(double x, double y, double xi)
{
    XY_VARS;
    double K = KA1_Gq_half_nfl(X1, Y1) - KA1_Gq_half_nfl(X2, Y2);
    K = 0.5 * K / xi;
    K = K * (2.0 * xi);
    return K;
}

/* kernels_nlo.f90:709-753; `GG_f_C` is GG_f_c */
static double KV1_GG_half(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = -R32(11, 3) * CA;
    double GGfV = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y) + 2.0 * GG_f_c(X, Y);
    double GGfbarV = 2.0 * GG_f_a(Xbar, Ybar) + GG_f_b(Xbar, Ybar) + 2.0 * GG_f_c(Xbar, Ybar);
    double GGfc = GG_f_c(X, Y);
    double GGhV = GG_h_V(X, Y);
    double GGhbarV = GG_hbar_V(Xbar, Y);
    double piece1 = SQ(CA) * (
        (R32(2, 3) - 2.0 * ZETA2) * GGfV - 0.5 * GGfc + X * (4.0 + 5.0 * X) / (4.0 * SQ(Y))
        - X * (10.0 - 19.0 * X + 16.0 * SQ(X)) / Y + X * (6.0 - 13.0 * X + 8.0 * SQ(X)) / Ybar
        + (2.0 * X * (1.0 - 2.0 * Xbar * (X * Y + Xbar * Ybar)) / (Y * Ybar)
           + X * (2.0 - 3.0 * X - 8.0 * SQ(Xbar)) / SQ(Ybar)) * abslog(X / Y)
        + (GGfV + 2.0 * SQ(X) / SQ(Ybar)) * log2_(X / Y)
        - (GGfV - GGfbarV) * logprod(X / Y, 1.0 - X / Y)
        - 0.5 * GGhV - 0.5 * GGhbarV);
    double piece2 = 0.5 * CA * beta0 * (
        -R32(5, 3) * GGfV - R32(13, 3) * GGfc - R32(11, 2) * SQ(X) / SQ(Y) + X * (X + Ybar) / (SQ(Y) * Ybar)
        + SQ(X) / SQ(Ybar) * abslog(Y) + SQ(Xbar) / SQ(Y) * abslog(Xbar));
    double piece4 = SQ(CA) * (
        -2.0 * SQ(X) / SQ(Ybar) * log2_(X)
        + (Xbar * (SQ(X) + SQ(Xbar) - 2.0 * Xbar * (1.0 + 2.0 * X) * Ybar) + Ybar) / SQ(Ybar) * logprod(X, Xbar)
        + ((3.0 * X * (SQ(X) + SQ(Xbar)) * (Ybar - Y) - 4.0 * SQ(X) * Y) / (Y * Ybar)
           - X * (2.0 - 3.0 * X - 8.0 * SQ(Xbar)) / SQ(Ybar)) * abslog(X));
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * rho_step(X, Y);
    double piece3 = 0.0;
    piece4 = piece4 * erbl_step(X);
    double piece5 = 0.0;
    return piece1 + piece2 + piece3 + piece4 + piece5;
}
/* kernels_nlo.f90:676-689 */
double tko_KV1_GG_pls(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_GG_half(X1, Y1) + KV1_GG_half(X2, Y2);
    return 0.5 * K / xi;
}
/* kernels_nlo.f90:790-831 */
static double KA1_GG_half(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = -R32(11, 3) * CA;
    double GGfA = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y);
    double GGfbarA = 2.0 * GG_f_a(Xbar, Ybar) + GG_f_b(Xbar, Ybar);
    double GGhA = GG_h_A(X, Y);
    double GGhbarA = GG_hbar_A(Xbar, Y);
    double piece1 = SQ(CA) * (
        (R32(2, 3) - 2.0 * ZETA2) * GGfA - R32(15, 4) * SQ(X) / SQ(Y) - 6.0 * X * Xbar / (Y * Ybar)
        + X * (X + Ybar) / (SQ(Y) * Ybar)
        - (2.0 * X / (Y * Ybar) - 6.0 * X * Xbar / SQ(Ybar) + SQ(X) / SQ(Ybar)) * abslog(X / Y)
        + (GGfA + 2.0 * SQ(X) / SQ(Ybar)) * log2_(X / Y) - (GGfA - GGfbarA) * logprod(X / Y, 1.0 - X / Y)
        - 0.5 * GGhA + 0.5 * GGhbarA);
    double piece2 = 0.5 * CA * beta0 * (
        -R32(5, 3) * GGfA - R32(5, 2) * SQ(X) / SQ(Y) + X * (Ybar + X * Y) / (SQ(Y) * Ybar)
        + SQ(X) / SQ(Ybar) * abslog(Y) + SQ(Xbar) / SQ(Y) * abslog(Xbar));
    double piece4 = SQ(CA) * (
        (3.0 * (Xbar - X) * X / (Y * Ybar) - 6.0 * X * Xbar / SQ(Ybar) + SQ(X) / SQ(Ybar)) * abslog(X)
        - 2.0 * SQ(X) / SQ(Ybar) * (log2_(X) + logprod(X, Xbar)) + (3.0 * X - Y) / SQ(Ybar) * logprod(X, Xbar));
    piece1 = piece1 * rho_step(X, Y);
    piece2 = piece2 * rho_step(X, Y);
    double piece3 = 0.0;
    piece4 = piece4 * erbl_step(X);
    double piece5 = 0.0;
    return piece1 + piece2 + piece3 + piece4 + piece5;
}
/* kernels_nlo.f90:755-768 */
double tko_KA1_GG_pls(double x, double y, double xi)
{
    XY_VARS;
    double K = KA1_GG_half(X1, Y1) + KA1_GG_half(X2, Y2);
    return 0.5 * K / xi;
}
/* kernels_nlo.f90:872-907 */
static double KV1_GG_half_nfl(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = R32(4, 3) * TF;
    double GGfV = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y) + 2.0 * GG_f_c(X, Y);
    double GGfc = GG_f_c(X, Y);
    double piece2 = 0.5 * CA * beta0 * (
        -R32(5, 3) * GGfV - R32(13, 3) * GGfc - R32(11, 2) * SQ(X) / SQ(Y) + X * (X + Ybar) / (SQ(Y) * Ybar)
        + SQ(X) / SQ(Ybar) * abslog(Y) + SQ(Xbar) / SQ(Y) * abslog(Xbar));
    double piece3 = CF * TF * (
        -R32(20, 3) * GGfc - 12.0 * SQ(X) / SQ(Y) + 2.0 * X * (2.0 + Y + 3.0 * X * Y - 4.0 * SQ(Y)) / (SQ(Y) * Ybar)
        - 2.0 * X * (X - Y + 2.0 * X * Y) / (Y * SQ(Ybar)) * abslog(X / Y) - SQ(X) / SQ(Ybar) * log2_(X / Y));
    double piece5 = CF * TF * (
        -2.0 * X * (1.0 - 3.0 * X - 2.0 * X * Ybar) / SQ(Ybar) * abslog(X) + SQ(X) / SQ(Ybar) * log2_(X));
    double piece1 = 0.0;
    piece2 = piece2 * rho_step(X, Y);
    piece3 = piece3 * rho_step(X, Y);
    double piece4 = 0.0;
    piece5 = piece5 * erbl_step(X);
    return piece1 + piece2 + piece3 + piece4 + piece5;
}
/* kernels_nlo.f90:836-849 */
double tko_KV1_GG_pls_nfl(double x, double y, double xi)
{
    XY_VARS;
    double K = KV1_GG_half_nfl(X1, Y1) + KV1_GG_half_nfl(X2, Y2);
    return 0.5 * K / xi;
}
/* kernels_nlo.f90:946-982 */
static double KA1_GG_half_nfl(double X, double Y)
{
    double Xbar = 1.0 - X, Ybar = 1.0 - Y;
    double beta0 = R32(4, 3) * TF;
    double GGfA = 2.0 * GG_f_a(X, Y) + GG_f_b(X, Y);
    double piece2 = 0.5 * CA * beta0 * (
        -R32(5, 3) * GGfA - R32(5, 2) * SQ(X) / SQ(Y) + X * (Ybar + X * Y) / (SQ(Y) * Ybar)
        + SQ(X) / SQ(Ybar) * abslog(Y) + SQ(Xbar) / SQ(Y) * abslog(Xbar));
    double piece3 = CF * TF * (
        2.0 * X * (1.0 - 2.0 * X * Ybar) / (SQ(Y) * Ybar) + 2.0 * SQ(X) / (Y * Ybar) * abslog(X / Y)
        - X / SQ(Ybar) * (2.0 * abslog(X / Y) + X * log2_(X / Y)));
    double piece5 = CF * TF * (X / SQ(Ybar) * (2.0 * abslog(X) + X * log2_(X)));
    double piece1 = 0.0;
    piece2 = piece2 * rho_step(X, Y);
    piece3 = piece3 * rho_step(X, Y);
    double piece4 = 0.0;
    piece5 = piece5 * erbl_step(X);
    return piece1 + piece2 + piece3 + piece4 + piece5;
}
/* kernels_nlo.f90:909-923 */
double tko_KA1_GG_pls_nfl(double x, double y, double xi)
{
    XY_VARS;
    double K = KA1_GG_half_nfl(X1, Y1) + KA1_GG_half_nfl(X2, Y2);
    return 0.5 * K / xi;
}

/* ------------------------------------------- NLO plus-prescription constants */

typedef struct { tko_k3 pls; double x, xi; } cst_ctx;
/* the `integrand` closures of kernels_nlo.f90:71-77 etc. */
static double cst_integrand(double y, void *vctx)
{
    cst_ctx *c = (cst_ctx *)vctx;
    return c->pls(c->x, y, c->xi) - c->pls(y, c->x, c->xi);
}
static const tko_k3 cst_pls[7] = {
    tko_KV1_NSp_pls, tko_KV1_NSm_pls, tko_KV1_NSp_pls_nfl,
    tko_KV1_GG_pls, tko_KA1_GG_pls, tko_KV1_GG_pls_nfl, tko_KA1_GG_pls_nfl };

double tko_cst_integral(int which, double x, double xi, int last[3], int ifail[3])
{
    cst_ctx c = { cst_pls[which], x, xi };
    return tko_adaptive_integrate(cst_integrand, &c, x, xi, last, ifail);
}
static double cst_int(tko_k3 pls, double x, double xi)
{
    cst_ctx c = { pls, x, xi };
    return tko_adaptive_integrate(cst_integrand, &c, x, xi, 0, 0);
}

/* kernels_nlo.f90:60-78 */
double tko_KV1_NSp_cst(double x, double xi)
{
    double K = CF * (CF - 0.5 * CA) * (6.5 - 6.0 * ZETA2 + 4.0 * ZETA3);
    return K + cst_int(tko_KV1_NSp_pls, x, xi);
}
/* kernels_nlo.f90:95-111 */
double tko_KV1_NSm_cst(double x, double xi) { return cst_int(tko_KV1_NSm_pls, x, xi); }
/* kernels_nlo.f90:162-178 */
double tko_KV1_NSp_cst_nfl(double x, double xi) { return cst_int(tko_KV1_NSp_pls_nfl, x, xi); }
/* kernels_nlo.f90:195-211 */
double tko_KV1_NSm_cst_nfl(double x, double xi) { return cst_int(tko_KV1_NSm_pls_nfl, x, xi); }
/* kernels_nlo.f90:691-707 (no explicit constant) */
double tko_KV1_GG_cst(double x, double xi) { return cst_int(tko_KV1_GG_pls, x, xi); }
/* kernels_nlo.f90:770-788: the explicit constant at :777 is overwritten at :779 (reference bug, kept) */
double tko_KA1_GG_cst(double x, double xi)
{
    double K = SQ(CA) * (R32(95, 27) - R32(14, 3) * ZETA2 + 2.0 * ZETA3);
    K = cst_int(tko_KA1_GG_pls, x, xi);
    return K;
}
/* kernels_nlo.f90:851-870 */
double tko_KV1_GG_cst_nfl(double x, double xi)
{
    double K = -R32(1, 108) * (35.0 * CA + 74.0 * CF);
    return K + cst_int(tko_KV1_GG_pls_nfl, x, xi);
}
/* kernels_nlo.f90:925-944 */
double tko_KA1_GG_cst_nfl(double x, double xi)
{
    double K = R32(1, 54) * (29.0 * CA - 28.0 * CF);
    return K + cst_int(tko_KA1_GG_pls_nfl, x, xi);
}

/* ------------------------------------------------------------- index tables */

const char *const tko_kernel_names[23] = {
    "K0_qq_pls", "KA0_qG_reg", "KVmA0_qG_reg", "KA0_Gq_reg", "KVmA0_Gq_reg", "KA0_GG_pls", "KVmA0_GG_reg",
    "KV1_NSp_pls", "KV1_NSm_pls", "KV1_NS_pls_nfl",
    "KV1_qq_reg", "KA1_qq_reg", "KV1_qG_reg", "KA1_qG_reg",
    "KV1_Gq_reg", "KA1_Gq_reg", "KV1_Gq_reg_nfl", "KA1_Gq_reg_nfl",
    "KV1_GG_pls", "KA1_GG_pls", "KV1_GG_pls_nfl", "KA1_GG_pls_nfl", "KV1_NSm_pls_nfl" };

double tko_kernel_by_index(int idx, double x, double y, double xi)
{
    static const tko_k3 tab[23] = {
        tko_K0_qq_pls, tko_KA0_qG_reg, tko_KVmA0_qG_reg, tko_KA0_Gq_reg, tko_KVmA0_Gq_reg,
        tko_KA0_GG_pls, tko_KVmA0_GG_reg,
        tko_KV1_NSp_pls, tko_KV1_NSm_pls, tko_KV1_NSp_pls_nfl,
        tko_KV1_qq_reg, tko_KA1_qq_reg, tko_KV1_qG_reg, tko_KA1_qG_reg,
        tko_KV1_Gq_reg, tko_KA1_Gq_reg, tko_KV1_Gq_reg_nfl, tko_KA1_Gq_reg_nfl,
        tko_KV1_GG_pls, tko_KA1_GG_pls, tko_KV1_GG_pls_nfl, tko_KA1_GG_pls_nfl, tko_KV1_NSm_pls_nfl };
    return tab[idx](x, y, xi);
}
