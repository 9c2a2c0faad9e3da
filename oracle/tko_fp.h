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
 *   - log: the classic fdlibm/musl reduction  x = 2^k (1+f), s = f/(2+f),
 *     log(1+f) = f - hfsq + s (hfsq + R(s^2)), degree-14 minimax R (Lg1..Lg7);
 *     error < 1 ulp.
 *   - pow(x>0, y): double-double log (two-sum + fma residuals), exact product
 *     with y, fdlibm-style exp of the hi/lo pair; error ~ 1 ulp.
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
static const double TKO_LG1 = 6.666666666666735130e-01;
static const double TKO_LG2 = 3.999999999940941908e-01;
static const double TKO_LG3 = 2.857142874366239149e-01;
static const double TKO_LG4 = 2.222219843214978396e-01;
static const double TKO_LG5 = 1.818357216161805012e-01;
static const double TKO_LG6 = 1.531383769920937332e-01;
static const double TKO_LG7 = 1.479819860511658591e-01;

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

static inline double tko_log(double x)
{
    uint64_t ix = tko_bits(x);
    if ((ix << 1) == 0) return -TKO_INF;              /* log(+-0) = -inf  */
    if (ix >> 63) return TKO_NAN;                     /* log(negative)    */
    if ((ix >> 52) >= 0x7ff) return x;                /* +inf or NaN      */
    if (ix == 0x3ff0000000000000ULL) return 0.0;
    int k;
    double m = tko_log_reduce(x, &k);
    double f = m - 1.0;
    double hfsq = 0.5 * f * f;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    /* fdlibm e_log.c with the Horner steps and the final sums fused; the library's log_core fuses the same ones */
    double t1 = w * fma(w, fma(w, TKO_LG6, TKO_LG4), TKO_LG2);
    double t2 = z * fma(w, fma(w, fma(w, TKO_LG7, TKO_LG5), TKO_LG3), TKO_LG1);
    double R = t2 + t1;
    double dk = (double)k;
    double u = fma(s, hfsq + R, dk * TKO_LN2_LO);
    return fma(dk, TKO_LN2_HI, f - (hfsq - u));
}

/* log(x) = hi + lo with ~2^-60 relative accuracy, x positive finite non-zero */
static inline void tko_log_dd(double x, double *hi, double *lo)
{
    int k;
    double m = tko_log_reduce(x, &k);
    double f = m - 1.0;                               /* exact */
    double d = 2.0 + f;
    double dl = f - (d - 2.0);                        /* exact: d + dl = 2 + f */
    double sh = f / d;
    double r = fma(-sh, d, f);
    r = r - sh * dl;
    double sl = r / d;                                /* s = sh + sl */
    double z = sh * sh;
    double w = z * z;
    double t1 = w * (TKO_LG2 + w * (TKO_LG4 + w * TKO_LG6));
    double t2 = z * (TKO_LG1 + w * (TKO_LG3 + w * (TKO_LG5 + w * TKO_LG7)));
    double R = t2 + t1;                               /* log(m) = 2 s + s R */
    double dk = (double)k;
    double A = dk * TKO_LN2_HI;                       /* exact */
    double B = 2.0 * sh;
    double h = A + B;                                 /* two-sum (Knuth) */
    double bb = h - A;
    double l = (A - (h - bb)) + (B - bb);
    double tail = l + ((2.0 * sl + sh * R) + dk * TKO_LN2_LO);
    double H = h + tail;
    *lo = tail - (H - h);
    *hi = H;
}

static const double TKO_INVLN2 = 1.44269504088896338700e+00;
static const double TKO_EP1 =  1.66666666666666019037e-01;
static const double TKO_EP2 = -2.77777777770155933842e-03;
static const double TKO_EP3 =  6.61375632143793436117e-05;
static const double TKO_EP4 = -1.65339022054652515390e-06;
static const double TKO_EP5 =  4.13813679705723846039e-08;

/* exp(ph + pl), |pl| << |ph| */
static inline double tko_exp_dd(double ph, double pl)
{
    if (ph != ph) return ph;
    if (ph > 709.782712893384) return TKO_INF;
    if (ph < -745.2) return 0.0;
    double kf = floor(ph * TKO_INVLN2 + 0.5);
    double hi = ph - kf * TKO_LN2_HI;
    double lo = kf * TKO_LN2_LO - pl;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (TKO_EP1 + t * (TKO_EP2 + t * (TKO_EP3 + t * (TKO_EP4 + t * TKO_EP5))));
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    int k = (int)kf;
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
 * correctly rounded in practice, like the libm pow behind the reference (tko_pow is good to ~0.6 ulp only, and
 * far more expensive).  t >= 1 gives a value >= 1, which min(1, .) turns into 1.  Everything else (NaN, zero,
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
#define tko_pow(x, y) pow((x), (y))
#define tko_pow15_capped(t) (((t) >= 1.0) ? 1.0 : pow((t), 1.5))
#endif

#endif
