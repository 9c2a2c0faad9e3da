/* tko_fp.h -- floating-point primitives of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product (tiktaalik_b200/) never includes
 * or links anything from this directory.
 *
 * Why own log/pow instead of libm: the reference calls the Fortran intrinsics
 * log and ** (kernels_common.f90:66-93, gridspace.f90:36-41, integration.f90:214).
 * At small xi the V-type NLO gluon kernels amplify a 1-ulp difference in log by
 * up to 1e14, so "GPU vs oracle <= 1e-9" is only reachable
 * if both sides round every log/pow identically.  glibc and the CUDA math library
 * differ in the last bit, therefore both the oracle (this file, plain C) and the
 * CUDA path (tiktaalik_b200/csrc/tkb_fp.cuh, written separately) implement the
 * SAME published algorithms with the SAME operation order using only IEEE-754
 * correctly rounded operations (+ - * / sqrt fma):
 *   - log: table-driven, x = 2^k m, rho = m invc_i - 1 formed exactly, L_i = -log(invc_i)
 *     tabulated as hi + lo, error-free sums of the large pieces, one final rounding:
 *     correctly rounded in practice (see the comment above tko_log_core).
 *   - pow(x>0, y): table-driven double-double log, exact product with y, table-driven
 *     exponential of the hi/lo pair with one final rounding: correctly rounded in
 *     practice (see the comment above tko_log_dd).
 * Compile with -ffp-contract=off (no implicit FMA) and without -ffast-math.
 */
#ifndef TKO_FP_H
#define TKO_FP_H

#include <stdint.h>
#include <string.h>
#include <math.h>   /* sqrt, fabs, floor, fma: exact / correctly rounded primitives only */

static inline uint64_t tko_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double   tko_from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

#define TKO_INF  (tko_from_bits(0x7ff0000000000000ULL))
#define TKO_NAN  (tko_from_bits(0x7ff8000000000000ULL))
#define TKO_HUGE (tko_from_bits(0x7fefffffffffffffULL))   /* Fortran huge(1d0)    */
#define TKO_TINY (tko_from_bits(0x0010000000000000ULL))   /* Fortran tiny(1d0)    */
#define TKO_EPS  (tko_from_bits(0x3cb0000000000000ULL))   /* Fortran epsilon(1d0) */

/* Fortran min/max on two reals, written as explicit compares so that host and
 * device agree even when a NaN is involved (result = second argument then). */
static inline double tko_min(double a, double b) { return (a < b) ? a : b; }
static inline double tko_max(double a, double b) { return (a > b) ? a : b; }

/* a float32 quotient of two default-kind literals, e.g. Fortran `11./3.` */
#define TKO_R32(a, b) ((double)((float)(a) / (float)(b)))

static const double TKO_LN2_HI = 6.93147180369123816490e-01; /* 0x3fe62e42fee00000 */
static const double TKO_LN2_LO = 1.90821492927058770002e-10; /* 0x3dea39ef35793c76 */

/* Argument reduction shared by log and log_dd: returns k and m in [sqrt(1/2), sqrt(2))
 * with x = 2^k m.  x must be positive, finite and non-zero on entry. */
static inline double tko_log_reduce(double x, int *kout)
{
    int k = 0;
    uint64_t ix = tko_bits(x);
    uint32_t hx = (uint32_t)(ix >> 32);
    if (hx < 0x00100000u) {          /* subnormal: scale up by 2^54 */
        k -= 54;
        x = x * 18014398509481984.0;
        ix = tko_bits(x);
        hx = (uint32_t)(ix >> 32);
    }
    hx += 0x3ff00000u - 0x3fe6a09eu;
    k += (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    ix = ((uint64_t)hx << 32) | (ix & 0xffffffffULL);
    *kout = k;
    return tko_from_bits(ix);
}

/* log: table-driven (tko_pow_tables.h), ~2^-63 relative before the one final rounding: 0.002 % of the logarithms of
 * arguments in (0, 2) and 0.03 % of those within 1e-2 of 1 differ from the correctly rounded value -- glibc's log, which a
 * real build of the reference calls, misses 0.1 % / 0.02 %.  Rounds 1-2 used fdlibm's algorithm (< 1 ulp, but 7.6 % of the
 * logarithms of arguments in (0, 2) one ulp off).
 *   x = 2^k m, m in [sqrt(1/2), sqrt(2)); 129 intervals of 2^13 steps of m's high word, placed so that m = 1 is the centre c
 *   of its interval (log c = 0 and rho = m - 1 exactly there), c = the midpoint elsewhere; invc_i = 1 / c_i to 21 bits (a high
 *   word), L_i = -log(invc_i) as a double + a high-word correction: a 16-byte entry; rho = m invc_i - 1 = a + b exactly (b = the
 *   fma residual), |rho| <= 2^-8,
 *   log x = [k LN2_HI + L_i.hi + a]  (error-free fast two-sums: every partial sum dominates the next term)
 *         + [two-sum errors + k LN2_LO + L_i.lo + b (1 - a) - a^2/2 + a^3/3 - a^4/4 + ... - a^8/8].
 * The CUDA side (tkb_fp.cuh: log_core) performs the same operations in the same order. */
#include "tko_pow_tables.h"

static inline double tko_log_core(double m, int k)
{
    uint32_t i = ((uint32_t)(tko_bits(m) >> 32) - 0x3fe6a09eu + 0x109eu) >> 13;      /* 0 <= i <= 128 */
    double invc = tko_from_bits((uint64_t)TKO_LOG_TAB[i].invc_hi << 32);
    double Lh = TKO_LOG_TAB[i].Lh;
    double Ll = tko_from_bits((uint64_t)TKO_LOG_TAB[i].Ll_hi << 32);
    double ph = m * invc;
    double b = fma(m, invc, -ph);
    double a = ph - 1.0;
    double sqh = a * a;
    /* a^3 (1/3 - a/4 + a^2/5 - a^3/6 + a^4/7 - a^5/8) */
    double q = fma(a, fma(a, fma(a, fma(a, fma(a, -0.125, 1.0 / 7.0), -1.0 / 6.0), 0.2), -0.25), 1.0 / 3.0);
    double t = (a * sqh) * q;
    double dk = (double)k;
    double A = dk * TKO_LN2_HI;                       /* exact: LN2_HI has 33 significant bits */
    double h1 = A + Lh;
    double e1 = Lh - (h1 - A);                        /* fast two-sum: A = 0 or |A| >= 0.69 > |Lh| */
    double h2 = h1 + a;
    double e2 = a - (h2 - h1);                        /* h1 = 0 (k = 0, c = 1) or |h1| >= 0.0078 > |a| */
    double low = (e1 + e2) + (fma(dk, TKO_LN2_LO, Ll) + (fma(-a, b, b) + fma(-0.5, sqh, t)));
    return h2 + low;
}

static inline double tko_log(double x)
{
    uint64_t ix = tko_bits(x);
    if ((ix << 1) == 0) return -TKO_INF;              /* log(+-0) = -inf  */
    if (ix >> 63) return TKO_NAN;                     /* log(negative)    */
    if ((ix >> 52) >= 0x7ff) return x;                /* +inf or NaN      */
    int k;
    double m = tko_log_reduce(x, &k);
    return tko_log_core(m, k);
}

/* ---- pow: table-driven, double-double intermediates (tables: tko_pow_tables.h, tools/make_pow_tables.py) ----
 * The grid-type-2 node map is x = +-(xi^2)^(1 -+ eta) (gridspace.f90:36-41): every node of the outer halves of the grid
 * is a pow result, so pow has to return what the libm behind a real build returns.  glibc's pow is correctly rounded in
 * all but ~0.1 % of the cases; a ~0.6 ulp pow (the fdlibm-style one used until round 2) disagrees with it in 10 %.
 * This one keeps ~2^-70 relative accuracy in the logarithm and ~2^-70 in the exponential before the single final
 * rounding, i.e. it is correctly rounded except when the exact result lies within ~2^-11 ulp of a rounding boundary.
 *   log x = k ln2 + L_i + log1p(rho),  x = 2^k m,  i = round(128 m) - 90,  rho = m invc_i - 1 EXACTLY as a + b
 *           (a = fl(m invc) - 1, b = the fma residual of the product),  L_i = -log(invc_i) tabulated as hi + lo;
 *           log1p(rho) = a - a^2/2 (exact product) + [b - a b + a^3 (1/3 - a/4 + ... - a^7/10)];  all large pieces are
 *           added with error-free two-sums, the small ones collected in one double.
 *   p = y log x as an exact product of y with the hi part plus y times the lo part.
 *   exp p = 2^e T_j (1 + expm1(r)),  128 p / ln2 = 128 e + j + (128 / ln2) r,  |r| <= ln2 / 256,  T_j = 2^(j/128) as hi + lo,
 *           expm1(r) = r + r^2/2 + ... + r^6/720 with r = rh + rl; T_j rh is formed exactly (fma), the rest in one double.
 * The CUDA side (tkb_fp.cuh: pow_) performs the same operations in the same order on the same tables. */
static inline void tko_two_sum(double a, double b, double *s, double *e)   /* Knuth: a + b = s + e exactly */
{
    double t = a + b;
    double bb = t - a;
    *e = (a - (t - bb)) + (b - bb);
    *s = t;
}

/* log(x) = hi + lo, x positive finite non-zero; ~2^-70 relative */
static inline void tko_log_dd(double x, double *hi, double *lo)
{
    int k;
    double m = tko_log_reduce(x, &k);
    int i = (int)(m * 128.0 + 0.5) - 90;              /* m*128 exact; 0 <= i <= 92 */
    double invc = TKO_POW_LOG[i][0];
    double ph = m * invc;
    double b = fma(m, invc, -ph);                     /* m invc = ph + b exactly */
    double a = ph - 1.0;                              /* exact (ph within [1/2, 2]) */
    double sqh = a * a;
    double sql = fma(a, a, -sqh);                     /* a^2 = sqh + sql exactly */
    double q = 1.0 / 3.0 + a * (-0.25 + a * (0.2 + a * (-1.0 / 6.0 + a * (1.0 / 7.0 + a * (-0.125 + a * (1.0 / 9.0 + a * -0.1))))));
    double t = (a * sqh) * q;                         /* a^3/3 - a^4/4 + ... - a^10/10 */
    double dk = (double)k;
    double h1, e1, h2, e2, h3, e3, h4, e4;
    tko_two_sum(dk * TKO_LN2_HI, TKO_POW_LOG[i][1], &h1, &e1);   /* k LN2_HI is exact (33 significant bits) */
    tko_two_sum(h1, a, &h2, &e2);
    tko_two_sum(h2, -0.5 * sqh, &h3, &e3);
    tko_two_sum(h3, t, &h4, &e4);
    double low = ((e1 + e2) + (e3 + e4)) + ((dk * TKO_LN2_LO + TKO_POW_LOG[i][2]) + ((b - a * b) - 0.5 * sql));
    double H = h4 + low;
    *lo = low - (H - h4);
    *hi = H;
}

/* log(x) from the double-double logarithm: correctly rounded in practice.  For the grid maps (gridspace.f90:49-61,
 * 95-99: eta = log(|x| / xi^2) / (2 log xi), Jacobian 2 log(xi) (xi^2)^(1 -+ eta)), where the value behind a real build
 * is libm's; the kernels' logarithms keep the cheaper tko_log. */
static inline double tko_log_cr(double x)
{
    uint64_t ix = tko_bits(x);
    if ((ix << 1) == 0) return -TKO_INF;
    if (ix >> 63) return TKO_NAN;
    if ((ix >> 52) >= 0x7ff) return x;
    double hi, lo;
    tko_log_dd(x, &hi, &lo);
    return hi;
}

/* exp(ph + pl), |pl| << |ph| */
static inline double tko_exp_dd(double ph, double pl)
{
    if (ph != ph) return ph;
    if (ph > 709.782712893384) return TKO_INF;
    if (ph < -745.2) return 0.0;
    double kf = floor(ph * TKO_POW_INV_LN2_128 + 0.5);
    double rh0 = ph - kf * (TKO_LN2_HI * 0.0078125);  /* exact: kf LN2_HI / 128 has < 53 bits, Sterbenz */
    double rl0 = pl - kf * (TKO_LN2_LO * 0.0078125);
    double rh, rl;
    tko_two_sum(rh0, rl0, &rh, &rl);
    int kk = (int)kf;
    int j = kk & 127;
    int k = (kk - j) / 128;
    double Th = TKO_POW_EXP[j][0], Tl = TKO_POW_EXP[j][1];
    double ql = rl + (rh * rh) * (0.5 + rh * (1.0 / 6.0 + rh * (1.0 / 24.0 + rh * (1.0 / 120.0 + rh * (1.0 / 720.0)))));
    double p1 = Th * rh;
    double p1l = fma(Th, rh, -p1);                    /* Th rh = p1 + p1l exactly */
    double s, e;
    tko_two_sum(Th, p1, &s, &e);
    double y = s + ((e + p1l) + (Tl + (Th * ql + Tl * rh)));
    if (k > 1023) {                                   /* y*2^k may still be finite */
        y = y * tko_from_bits((uint64_t)(1023 + 1023) << 52);
        k -= 1023;
        if (k > 1023) return TKO_INF;
    } else if (k < -1021) {                           /* gradual underflow */
        y = y * tko_from_bits((uint64_t)(1023 - 1000) << 52);
        k += 1000;
        if (k < -1021) return 0.0;
    }
    return y * tko_from_bits((uint64_t)(1023 + k) << 52);
}

/* Fortran x**y for real y (gridspace.f90:36-41,95-99; integration.f90:214) */
static inline double tko_pow(double x, double y)
{
    if (y == 0.0) return 1.0;
    if (x != x || y != y) return TKO_NAN;
    if (x == 1.0) return 1.0;
    if (x < 0.0) return TKO_NAN;                      /* non-integer powers only */
    if (x == 0.0) return (y > 0.0) ? 0.0 : TKO_INF;
    if (x == TKO_INF) return (y > 0.0) ? TKO_INF : 0.0;
    if (y == TKO_INF) return (x > 1.0) ? TKO_INF : 0.0;
    if (y == -TKO_INF) return (x > 1.0) ? 0.0 : TKO_INF;
    double lh, ll;
    tko_log_dd(x, &lh, &ll);
    double ph = y * lh;
    double pl = fma(y, lh, -ph) + y * ll;
    return tko_exp_dd(ph, pl);
}

/* x**1.5 as QUADPACK's error estimate needs it (integration.f90, qk21: (200*abserr/resasc)**1.5_dp under min(1, .)).
 * For 1e-150 < t < 1:  sqrt(t) = s + e with e from the exact residual t - s*s, so t**1.5 = t*s + t*e to ~2^-104:
 * correctly rounded in practice, like the libm pow behind the reference (and far cheaper than tko_pow).  t >= 1 gives a value >= 1, which min(1, .) turns into 1.  Everything else (NaN, zero,
 * below 1e-150) goes through tko_pow. */
static inline double tko_pow15_capped(double t)
{
    if (t >= 1.0) return 1.0;
    if (t > 1e-150) {
        double s = sqrt(t);
        double e = fma(-s, s, t) / (2.0 * s);
        double hi = t * s;
        double lo = fma(t, s, -hi) + t * e;
        return hi + lo;
    }
    return tko_pow(t, 1.5);
}

/* TIMING-ONLY build (libtko_fast.so, bench.py's CPU legs): libm's log / pow and the reference's own compiler flags
 * (f90src/CMakeLists.txt:12-14: -Ofast -march=native -fopenmp; here -march=x86-64-v3, AVX2 + FMA, because the built
 * library travels to another host), i.e. what a real build of the Fortran spends per
 * evaluation rather than what the bit-exact restatement above spends.  Its numbers differ from the oracle's in the
 * last bits (and QUADPACK may take a few more or fewer bisections); it is never used as a checker. */
#ifdef TKO_LIBM
#define tko_log(x) log(x)
#define tko_log_cr(x) log(x)
#define tko_pow(x, y) pow((x), (y))
#define tko_pow15_capped(t) (((t) >= 1.0) ? 1.0 : pow((t), 1.5))
#endif

#endif
