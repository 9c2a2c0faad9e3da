// tkb_kernels.cuh -- the BFM splitting kernels as sm_100a device code.
//
// What it computes (reference file:line):
//   specfun.f90:24-79          real dilogarithm
//   kernels_common.f90:37-399  step functions, log helpers, f_a/f_b/f_c, h / hbar
//   kernels_lo.f90:48-221      LO pieces
//   kernels_nlo.f90:45-982     NLO two-variable pieces ("half kernels")
//
// How it is organised (not like the reference): the reference evaluates each of its 23
// kernel functions separately, every one recomputing X/Y, the logarithms and up to ten
// dilogarithms per "half".  Here one call, eval_half<MASK>, evaluates every requested
// half kernel at one (X, Y) from ONE shared set of transcendental values (7 logs, at most
// 7 dilogs).  The assembly kernel runs it with all 22 bits set, one thread per
// (quadrature node, half); the QAGS kernel instantiates it with a single bit.
// Every shared value is formed by exactly the expression the reference uses, in its
// evaluation order, and the file is compiled with -fmad=false, so each half kernel
// rounds identically to a straight evaluation of the reference formula (the oracle checks
// this bit for bit).  Fortran default-kind literals such as 11./3. are float32 quotients.
#pragma once
#include "tkb_fp.cuh"

namespace tkb {

constexpr double CF = 4.0 / 3.0;
constexpr double CA = 3.0;
constexpr double TF = 0.5;
constexpr double PI = 3.141592653589793;           // acos(-1)
constexpr double PI2 = PI * PI;
constexpr double ZETA2 = PI2 / 6.0;                // constants.f90:22
constexpr double ZETA3 = 1.2020569031;             // constants.f90:23 (10 digits, as the reference)

// float32 rationals of kernels_lo.f90 / kernels_nlo.f90
#define TKB_R32(a, b) ((double)((float)(a) / (float)(b)))
constexpr double R11_3 = TKB_R32(11, 3);
constexpr double R4_3 = TKB_R32(4, 3);
constexpr double R5_3 = TKB_R32(5, 3);
constexpr double R10_3 = TKB_R32(10, 3);
constexpr double R2_3 = TKB_R32(2, 3);
constexpr double R13_3 = TKB_R32(13, 3);
constexpr double R4_9 = TKB_R32(4, 9);
constexpr double R26_9 = TKB_R32(26, 9);
constexpr double R290_9 = TKB_R32(290, 9);
constexpr double R20_3 = TKB_R32(20, 3);
constexpr double R11_2 = TKB_R32(11, 2);
constexpr double R5_2 = TKB_R32(5, 2);
constexpr double R15_4 = TKB_R32(15, 4);
constexpr double R95_27 = TKB_R32(95, 27);
constexpr double R14_3 = TKB_R32(14, 3);
constexpr double R1_108 = TKB_R32(1, 108);
constexpr double R1_54 = TKB_R32(1, 54);

// ---------------------------------------------------------------------------------------
// dilogarithm (specfun.f90:24-79).  The reference recurses once into [0, 1/2]; the inner
// call can only see 0, 1/2 (float32 log(2.) constant, specfun.f90:33) or the rational core.

static __constant__ double c_dlp[6] = { 0.9999999999999999502e+0, -2.6883926818565423430e+0, 2.6477222699473109692e+0,
                                 -1.1538559607887416355e+0, 2.0886077795020607837e-1, -1.0859777134152463084e-2 };
static __constant__ double c_dlq[7] = { 1.0000000000000000000e+0, -2.9383926818565635485e+0, 3.2712093293018635389e+0,
                                 -1.7076702173954289421e+0, 4.1596017228400603836e-1, -3.9801343754084482956e-2,
                                 8.2743668974466659035e-4 };

TKB_DEV double dilog_half(double x)
{
    const double x2 = x * x;
    const double x3 = x * x2;
    const double x4 = x2 * x2;
    const double x5 = x4 * x;
    const double x6 = x3 * x3;
    double px = c_dlp[0];
    px = px + x * c_dlp[1];
    px = px + x2 * c_dlp[2];
    px = px + x3 * c_dlp[3];
    px = px + x4 * c_dlp[4];
    px = px + x5 * c_dlp[5];
    px = px + x6 * 0.0;
    double qx = c_dlq[0];
    qx = qx + x * c_dlq[1];
    qx = qx + x2 * c_dlq[2];
    qx = qx + x3 * c_dlq[3];
    qx = qx + x4 * c_dlq[4];
    qx = qx + x5 * c_dlq[5];
    qx = qx + x6 * c_dlq[6];
    return x * px / qx;
}

// value of dilog at the float32-flavoured special point z = 1/2
TKB_DEV double dilog_at_half()
{
    const float l2 = 0.693147182464599609375f;              // float32 log(2.)
    const float h = __fdiv_rn(__fmul_rn(l2, l2), 2.0f);
    return PI2 / 12.0 - (double)h;
}

// Branch-uniform evaluation: every range of the reference's if-chain needs at most two logarithms
// and one reduced dilogarithm, so the arguments are selected first and log_ / dilog_half are called
// unconditionally -- lanes of a warp that fall into different ranges do not diverge.  Each range's
// value is formed by exactly the reference's expression (specfun.f90:36-49).
#ifdef TKB_DILOG_INLINE
TKB_DEV double dilog(double z)
#else
static __device__ __noinline__ double dilog(double z)
#endif
{
    // range id: 0 z<-1, 1 -1<z<0, 2 z>2, 3 1<z<2, 4 1/2<z<1, 5 rational core (0<z<1/2, NaN), 6 special value
    int rg;
    if (z == -1.0 || z == 0.0 || z == 0.5 || z == 1.0 || z == 2.0) rg = 6;
    else if (z < -1.0) rg = 0;
    else if (z < 0.0) rg = 1;
    else if (z > 2.0) rg = 2;
    else if (z > 1.0) rg = 3;
    else if (z > 0.5) rg = 4;
    else rg = 5;
    const double omz = 1.0 - z;
    // reduced argument w and the two log arguments
    double w, la, lb;
    switch (rg) {
    case 0: w = 1.0 / omz; la = omz; lb = -z; break;
    case 1: w = z / (z - 1.0); la = omz; lb = 1.0; break;
    case 2: w = 1.0 / z; la = z; lb = 1.0; break;
    case 3: w = 1.0 - 1.0 / z; la = z; lb = w; break;
    case 4: w = omz; la = z; lb = omz; break;
    default: w = z; la = 1.0; lb = 1.0; break;
    }
    const double l1 = log_or_zero(la);
    const double l2 = log_or_zero(lb);
    double D = dilog_half(w);
    if (rg != 5 && rg != 6) {                 // the inner call of the reference sees the special points too
        if (w == 0.0) D = 0.0;
        else if (w == 0.5) D = dilog_at_half();
        else if (w == 1.0) D = PI2 / 6.0;
        else if (w == -1.0) D = -PI2 / 12.0;
    }
    switch (rg) {
    case 0: return l1 * (0.5 * l1 - l2) - PI2 / 6.0 + D;
    case 1: return -D - 0.5 * (l1 * l1);
    case 2: return -D + PI2 / 3.0 - 0.5 * (l1 * l1);
    case 3: return PI2 / 6.0 - l1 * (l2 + 0.5 * l1) + D;
    case 4: return -D + PI2 / 6.0 - l1 * l2;
    case 5: return D;
    }
    if (z == -1.0) return -PI2 / 12.0;
    if (z == 0.0) return 0.0;
    if (z == 0.5) return dilog_at_half();
    if (z == 1.0) return PI2 / 6.0;
    return PI2 / 4.0;
}

// ---------------------------------------------------------------------------------------
// steps and log helpers (kernels_common.f90:37-93)

TKB_DEV double theta_step(double X) { return (X > 0.0) ? 1.0 : ((X == 0.0) ? 0.5 : 0.0); }
TKB_DEV double rho_step(double r, double omr, double Y)      // r = X/Y, omr = 1.-X/Y
{
    return theta_step(omr) * theta_step(r) * copysign(1.0, Y);
}
// log2(x) from l = log|x|
TKB_DEV double sqlog(double l, double x)
{
    double y = l * l;
    if (x < 0.0) y = y - PI2;
    return y;
}
// logprod(x1,x2) from l1 = log|x1|, l2 = log|x2|
TKB_DEV double prodlog(double l1, double x1, double l2, double x2)
{
    double y = l1 * l2;
    if (x1 < 0.0 && x2 < 0.0) y = y - PI2;
    return y;
}

// ---------------------------------------------------------------------------------------
// rational building blocks (kernels_common.f90:100-341)

TKB_DEV double qq_fab(double X, double Y) { return X / Y + X / Y / (Y - X); }       // f_a + f_b
TKB_DEV double qg_fa(double X, double Y) { return X / (Y * Y); }
TKB_DEV double qg_fc(double X, double Y) { return X / (Y * Y) * (2.0 * (1.0 - X) * Y - X); }
TKB_DEV double qg_fV(double X, double Y) { return -qg_fa(X, Y) - 2.0 * qg_fc(X, Y); }
TKB_DEV double qg_fA(double X, double Y) { return -qg_fa(X, Y); }
TKB_DEV double gq_fa(double X, double Y) { return (X * X) / Y; }
TKB_DEV double gq_fc(double X, double Y) { return (X * X) / Y * (2.0 * (1.0 - X) * Y - (1.0 - Y)); }
TKB_DEV double gq_fV(double X, double Y) { return gq_fa(X, Y) + 2.0 * gq_fc(X, Y); }
TKB_DEV double gg_fa(double X, double Y) { return (X * X) / (Y * Y); }
TKB_DEV double gg_fb(double X, double Y) { return (X * X) / (Y * Y) / (Y - X); }
TKB_DEV double gg_fc(double X, double Y) { return (X * X) / (Y * Y) * (2.0 * (1.0 - X) * Y + Y - X); }
TKB_DEV double gg_fV(double X, double Y) { return 2.0 * gg_fa(X, Y) + gg_fb(X, Y) + 2.0 * gg_fc(X, Y); }
TKB_DEV double gg_fA(double X, double Y) { return 2.0 * gg_fa(X, Y) + gg_fb(X, Y); }

// h and hbar of Eq. (174) from shared transcendental values.
//   h(X,Y)    needs  LP = logprod(1-X, Y),  DS = dilog(X) + dilog(1-Y)
//   hbar(A,Y) needs  G = 2 dilog(1-A/Y) + log2(Y),  DYb = dilog(1-Y),  LPA = logprod(A,Y),
//                    DAb = dilog(1-A);   f = f(A,Y), fbar = f(1-A, 1-Y)
// "pp": QQ and GG sign pattern (kernels_common.f90:125-148, 343-399);
// "mm": QG and GQ sign pattern (kernels_common.f90:173-234, 260-321).
TKB_DEV double h_pp(double f, double fbar, double LP, double DS) { return 2.0 * fbar * LP - 2.0 * f * DS; }
TKB_DEV double h_mm(double f, double fbar, double LP, double DS) { return -2.0 * fbar * LP - 2.0 * f * DS; }
TKB_DEV double hbar_pp(double f, double fbar, double G, double DYb, double LPA, double DAb)
{
    return (-fbar + f) * G + 2.0 * f * (DYb - LPA) + 2.0 * fbar * DAb;
}
TKB_DEV double hbar_mm(double f, double fbar, double G, double DYb, double LPA, double DAb)
{
    return (fbar + f) * G + 2.0 * f * (DYb - LPA) - 2.0 * fbar * DAb;
}

// ---------------------------------------------------------------------------------------
// kernel indices (order shared with the C-ABI, include/tiktaalik_b200.h)

enum KernelId : int {
    K_K0QQ = 0, K_A0QG, K_VMA0QG, K_A0GQ, K_VMA0GQ, K_A0GG, K_VMA0GG,
    K_NSP, K_NSM, K_NSN,
    K_V1QQ, K_A1QQ, K_V1QG, K_A1QG,
    K_V1GQ, K_A1GQ, K_V1GQN, K_A1GQN,
    K_V1GG, K_A1GG, K_V1GGN, K_A1GGN,
    NKERN
};
constexpr unsigned MASK_ALL = (1u << NKERN) - 1u;
__host__ __device__ constexpr unsigned bit(int k) { return 1u << k; }

// ---------------------------------------------------------------------------------------
// support classes of a half kernel (used by K1 and K2 to skip halves that are exact zeros)

// every denominator / log argument that any half kernel of the node's four tasks can meet is non-zero
// and finite, so a piece multiplied by a vanishing support step is an exact zero (not 0*Inf = NaN).
// (a1, a2) = (0.5 (1 + x/xi), 0.5 (1 - x/xi)), (b1, b2) the same for y.
TKB_DEV bool node_is_safe(double a1, double a2, double b1, double b2)
{
    const double big = huge_();
    bool ok = (fabs(a1) <= big) && (fabs(a2) <= big) && (fabs(b1) <= big) && (fabs(b2) <= big);
    const double v[4] = { a1, a2, b1, b2 };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double c = 1.0 - v[k];                 // Xbar / Ybar / 1 - Xo
        ok = ok && (v[k] != 0.0) && (c != 0.0) && (1.0 - c != 0.0);
    }
    // pairs (X, Y) that appear as Y - X, Ybar - Xbar, Y - Xo, Ybar - (1 - Xo), Y - Xbar, Ybar - Xbb, Y/X - 1:
    // any first-argument value against any second-argument value, in both directions
    const double eps4 = 4.0 * eps_();
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 2; j < 4; j++) {
            const double p = v[i], q = v[j];
            const double pb = 1.0 - p, qb = 1.0 - q;
            ok = ok && (fabs(q - p) > eps4 * fabs(p)) && (fabs(p - q) > eps4 * fabs(q)) && (qb != pb) &&
                 (q != pb) && (p != qb) && (qb != 1.0 - pb) && (pb != 1.0 - qb);
        }
    ok = ok && (a2 != a1) && (b2 != b1);             // Yo - Xo of the non-singlet halves: (a1,a2) vs (b1,b2) handled above
    return ok;
}

// support class of the half kernel at (X, Y): 2 full, 1 erbl-only, 0 dead.
// The evaluation forms r = X / Y, rho_step(r, 1 - r, Y) = theta(1 - r) theta(r) sign(Y) and
// erbl_step = theta(X) theta(1 - X) with theta(0) = 1/2.  For a safe node (X, Y finite, non-zero, X != Y:
// node_is_safe) the class follows from the operands without the division: a correctly rounded quotient has
// the sign of X Y and is <= 1 exactly when |X| <= |Y| (two different doubles differ by more than half an ulp of
// their quotient, so it cannot round to 1), hence rho != 0  <=>  sign X = sign Y and |X| <= |Y|; r = 0 or
// 1 - r = 0 would need X = 0 or X = Y.  An unsafe node is always evaluated in full.
TKB_DEV int task_class(double X, double Y, bool safe)
{
    if (!safe) return 2;
    if (((X < 0.0) == (Y < 0.0)) && fabs(X) <= fabs(Y)) return 2;
    return (X >= 0.0 && 1.0 - X >= 0.0) ? 1 : 0;
}

// How the two halves combine into K(x,y,xi) (kernels_lo.f90 / kernels_nlo.f90 wrappers):
//   sign  +1: K = h(X1,Y1) + h(X2,Y2);  -1: K = h(X1,Y1) - h(X2,Y2);  -2: K = -h(X1,Y1) + h(X2,Y2)
//   pref  colour factor c in  K = c*0.5*K/xi  (LO) or K = 0.5*K/xi (NLO, c = 0 marks "no colour")
//   xipow -1: K = K/(2 xi) (qG);  +1: K = K*(2 xi) (Gq);  0: nothing
template <int KID>
TKB_DEV double combine_halves(double h1, double h2, double xi)
{
    double K;
    if (KID == K_A0QG || KID == K_VMA0QG) K = -h1 + h2;
    else if (KID == K_A0GQ || KID == K_VMA0GQ || KID == K_V1QG || KID == K_A1QG || KID == K_V1GQ ||
             KID == K_A1GQ || KID == K_V1GQN || KID == K_A1GQN) K = h1 - h2;
    else K = h1 + h2;
    if (KID == K_K0QQ || KID == K_A0GQ || KID == K_VMA0GQ) K = CF * 0.5 * K / xi;
    else if (KID == K_A0QG || KID == K_VMA0QG) K = 2.0 * TF * 0.5 * K / xi;
    else if (KID == K_A0GG || KID == K_VMA0GG) K = CA * 0.5 * K / xi;
    else K = 0.5 * K / xi;
    if (KID == K_A0QG || KID == K_VMA0QG || KID == K_V1QG || KID == K_A1QG) K = K / (2.0 * xi);
    if (KID == K_A0GQ || KID == K_VMA0GQ || KID == K_V1GQ || KID == K_A1GQ || KID == K_V1GQN || KID == K_A1GQN)
        K = K * (2.0 * xi);
    return K;
}
__device__ inline double combine_halves_rt(int kid, double h1, double h2, double xi)
{
    switch (kid) {
#define TKB_CASE(k) case k: return combine_halves<k>(h1, h2, xi);
        TKB_CASE(K_K0QQ) TKB_CASE(K_A0QG) TKB_CASE(K_VMA0QG) TKB_CASE(K_A0GQ) TKB_CASE(K_VMA0GQ)
        TKB_CASE(K_A0GG) TKB_CASE(K_VMA0GG) TKB_CASE(K_NSP) TKB_CASE(K_NSM) TKB_CASE(K_NSN)
        TKB_CASE(K_V1QQ) TKB_CASE(K_A1QQ) TKB_CASE(K_V1QG) TKB_CASE(K_A1QG) TKB_CASE(K_V1GQ)
        TKB_CASE(K_A1GQ) TKB_CASE(K_V1GQN) TKB_CASE(K_A1GQN) TKB_CASE(K_V1GG) TKB_CASE(K_A1GG)
        TKB_CASE(K_V1GGN) TKB_CASE(K_A1GGN)
#undef TKB_CASE
    }
    return 0.0;
}

// ---------------------------------------------------------------------------------------
// one half: every requested half kernel at (X, Y); (Xo, Yo) is the other half's pair, used
// only by the non-singlet qq half kernels which take it as their "Xbar, Ybar"
// (kernels_nlo.f90:56, 113).  out[k] is written for every k in MASK.

// WITH_RHO = false evaluates only the pieces that carry erbl_step (the caller has established that
// rho_step(X,Y) == 0 and that every skipped piece is finite, so the skipped terms are exact zeros);
// it is implemented for the kernels the QAGS constants need.
constexpr unsigned MASK_QAGS = bit(K_NSP) | bit(K_NSM) | bit(K_NSN) | bit(K_V1GG) | bit(K_A1GG) | bit(K_V1GGN) | bit(K_A1GGN);

//
// FIX: transcendental values that depend on one argument only are the same for every quadrature node of a
// QAGS instance (its fixed point x is X in K(x,y) and Y in K(y,x)), so the K2 kernels compute them once per
// instance with the same functions on the same arguments -- bit-identical values -- and pass them in:
//   FIX = 1: X (and Xo) belong to the fixed point; FIX = 2: Y does.  preS / preO are the tables
//   { log|U|, log|1-U|, dilog(U), dilog(1-U), dilog(1-(1-U)) } of U = the fixed argument of this half
//   (X resp. Y) and of the other half (Xo).  FIX = 0: everything is evaluated here.
enum { PRE_L = 0, PRE_LB = 1, PRE_D = 2, PRE_DB = 3, PRE_DBB = 4, PRE_N = 5 };
TKB_DEV double pre_value(int which, double U)
{
    const double Ub = 1.0 - U;
    const double arg = (which == PRE_L || which == PRE_D) ? U : ((which == PRE_DBB) ? 1.0 - Ub : Ub);
    return (which >= PRE_D) ? dilog(arg) : abslog(arg);
}

// FIX = 4 (the fixed-quadrature kernel, which evaluates the 22 kernels of a node in several passes of a few kernels
// each): X belongs to the fixed point as with FIX = 1, and the transcendental values that depend on the node -- the
// same for every kernel -- were computed once by node_values() and are passed in `tv`.
enum { TV_LR = 0, TV_LOMR, TV_LYX1, TV_LY, TV_DYB, TV_GD, TV_GDO, TV_N };

// exactly the expressions eval_half uses for these values (same operations, same order)
TKB_DEV void node_values(double X, double Y, double Xo, double* __restrict__ tv)
{
    const double Xb = 1.0 - X, Yb = 1.0 - Y;
    const double r = X / Y;
    const double omr = 1.0 - r;
    const double yx1 = Y / X - 1.0;
    tv[TV_LR] = abslog(r);
    tv[TV_LOMR] = abslog(omr);
    tv[TV_LYX1] = abslog(yx1);
    tv[TV_LY] = abslog(Y);
    tv[TV_DYB] = dilog(Yb);
    tv[TV_GD] = dilog(1.0 - Xb / Y);
    tv[TV_GDO] = dilog(1.0 - Xo / Y);
}

template <unsigned MASK, bool WITH_RHO = true, int FIX = 0>
__device__ __forceinline__ void eval_half(double X, double Y, double Xo, double Yo, double* __restrict__ out,
                                          const double* __restrict__ preS = nullptr,
                                          const double* __restrict__ preO = nullptr, int fix_rt = 0,
                                          const double* __restrict__ tv = nullptr)
{
    // FIX = 3: the side of the fixed point is a run-time (warp-uniform) argument, one copy of the formulas serves both
    constexpr bool TV = (FIX == 4);
    const bool F1 = (FIX == 1) || TV || (FIX == 3 && fix_rt == 1);
    const bool F2 = (FIX == 2) || (FIX == 3 && fix_rt == 2);
    static_assert(WITH_RHO || (MASK & ~MASK_QAGS) == 0, "erbl-only evaluation exists for the QAGS kernels only");
    constexpr bool LO_ANY = (MASK & 0x7fu) != 0;
    constexpr bool NS = (MASK & (bit(K_NSP) | bit(K_NSM))) != 0;
    constexpr bool V1QQ = (MASK & bit(K_V1QQ)) != 0, A1QQ = (MASK & bit(K_A1QQ)) != 0;
    constexpr bool V1QG = (MASK & bit(K_V1QG)) != 0, A1QG = (MASK & bit(K_A1QG)) != 0;
    constexpr bool V1GQ = (MASK & bit(K_V1GQ)) != 0, A1GQ = (MASK & bit(K_A1GQ)) != 0;
    constexpr bool V1GQN = (MASK & bit(K_V1GQN)) != 0, A1GQN = (MASK & bit(K_A1GQN)) != 0;
    constexpr bool V1GG = (MASK & bit(K_V1GG)) != 0, A1GG = (MASK & bit(K_A1GG)) != 0;
    constexpr bool V1GGN = (MASK & bit(K_V1GGN)) != 0, A1GGN = (MASK & bit(K_A1GGN)) != 0;
    constexpr bool NSN = (MASK & bit(K_NSN)) != 0;
    // which shared values are needed
    constexpr bool HV = V1QG || A1QG || V1GQ || A1GQ || V1GG || A1GG;          // own-Xbar h / hbar users
    constexpr bool N_lr = WITH_RHO && (NS || NSN || V1QQ || A1QQ || V1QG || A1QG || V1GQ || A1GQ || V1GG || A1GG || V1GGN || A1GGN);
    constexpr bool N_lomr = WITH_RHO && (NS || V1QG || A1QG || V1GQ || A1GQ || V1GQN || A1GQN || V1GG || A1GG);
    constexpr bool N_lyx1 = WITH_RHO && (V1QG || A1QG || V1GQ || A1GQ);
    constexpr bool N_lX = (WITH_RHO || NS || V1GG || A1GG || V1GGN || A1GGN) && (NS || V1QQ || A1QQ || V1QG || A1QG || V1GQ || A1GQ || V1GQN || A1GQN || V1GG || A1GG || V1GGN || A1GGN);
    constexpr bool N_lXb = WITH_RHO ? (NS || V1QQ || HV || V1GQN || V1GGN || A1GGN) : (V1GG || A1GG);
    constexpr bool N_lY = WITH_RHO && (NS || V1QQ || HV || V1GGN || A1GGN);
    constexpr bool N_DX = WITH_RHO && (NS || HV);
    constexpr bool N_DYb = WITH_RHO && (NS || V1QQ || HV);
    constexpr bool N_DXb = WITH_RHO && V1QQ;

    const double Xb = 1.0 - X, Yb = 1.0 - Y;
    const double r = X / Y;
    const double omr = 1.0 - r;
    const double rho = rho_step(r, omr, Y);
    const double erbl = theta_step(X) * theta_step(Xb);

    double lr = 0, lomr = 0, lyx1 = 0, yx1 = 0, lX = 0, lXb = 0, lY = 0, DX = 0, DYb = 0, DXb = 0;
    if (N_lr) lr = TV ? tv[TV_LR] : abslog(r);
    if (N_lomr) lomr = TV ? tv[TV_LOMR] : abslog(omr);
    if (N_lyx1) { yx1 = Y / X - 1.0; lyx1 = TV ? tv[TV_LYX1] : abslog(yx1); }
    if (N_lX) lX = F1 ? preS[PRE_L] : abslog(X);
    if (N_lXb) lXb = F1 ? preS[PRE_LB] : abslog(Xb);
    if (N_lY) lY = TV ? tv[TV_LY] : (F2 ? preS[PRE_L] : abslog(Y));
    if (N_DX) DX = F1 ? preS[PRE_D] : dilog(X);
    if (N_DYb) DYb = TV ? tv[TV_DYB] : (F2 ? preS[PRE_DB] : dilog(Yb));
    if (N_DXb) DXb = F1 ? preS[PRE_DB] : dilog(Xb);
    const double L2r = sqlog(lr, r);                      // log2(X/Y)
    const double L2X = sqlog(lX, X);                      // log2(X)
    const double LPXbY = prodlog(lXb, Xb, lY, Y);         // logprod(Xbar, Y)
    const double DS = DX + DYb;                           // dilog(X) + dilog(Ybar)
    // hbar(Xbar, Y) with the half's own Xbar = 1 - X
    double G = 0, DXbb = 0, Xbb = 0;
    if (HV && WITH_RHO) {
        Xbb = 1.0 - Xb;
        G = 2.0 * (TV ? tv[TV_GD] : dilog(1.0 - Xb / Y)) + sqlog(lY, Y);
        DXbb = F1 ? preS[PRE_DBB] : dilog(Xbb);
    }

    // ---------------- LO (kernels_lo.f90:48-221) ----------------
    if (LO_ANY) {
        if (MASK & bit(K_K0QQ)) out[K_K0QQ] = qq_fab(X, Y) * rho;
        if (MASK & bit(K_A0QG)) out[K_A0QG] = qg_fa(X, Y) * rho;
        if (MASK & bit(K_VMA0QG)) out[K_VMA0QG] = 2.0 * qg_fc(X, Y) * rho;
        if (MASK & bit(K_A0GQ)) out[K_A0GQ] = gq_fa(X, Y) * rho;
        if (MASK & bit(K_VMA0GQ)) out[K_VMA0GQ] = 2.0 * gq_fc(X, Y) * rho;
        if (MASK & bit(K_A0GG)) out[K_A0GG] = (2.0 * gg_fa(X, Y) + gg_fb(X, Y)) * rho;
        if (MASK & bit(K_VMA0GG)) out[K_VMA0GG] = 2.0 * gg_fc(X, Y) * rho;
    }

    // ---------------- non-singlet qq (kernels_nlo.f90:113-142, 213-228) ----------------
    if (NS || NSN) {
        const double QQf = qq_fab(X, Y);
        if (NSN) {
            const double beta0 = R4_3 * TF;
            out[K_NSN] = WITH_RHO ? (-0.5 * CF * beta0 * (R5_3 * QQf + X / Y + QQf * lr)) * rho : 0.0;
        }
        if (NS) {
            const double lXo = F1 ? preO[PRE_L] : abslog(Xo);
            const double piece4 = -0.5 * X / Yo * (lX + L2X - 2.0 * prodlog(lX, X, lXo, Xo)) * (CF * CF);
            const double p4 = piece4 * erbl;
            if (!WITH_RHO) {
                if (MASK & bit(K_NSP)) out[K_NSP] = p4;
                if (MASK & bit(K_NSM)) out[K_NSM] = p4;
            } else {
                const double QQfbar = qq_fab(Xo, Yo);
                const double QQh = h_pp(QQf, qq_fab(Xb, Yb), LPXbY, DS);
                // QQ_hbar(Xo, Y): its X is Xo, its Xbar is 1 - Xo
                const double Xob = 1.0 - Xo;
                const double Go = 2.0 * (TV ? tv[TV_GDO] : dilog(1.0 - Xo / Y)) + sqlog(lY, Y);
                const double QQhbar = hbar_pp(qq_fab(Xo, Y), qq_fab(Xob, Yb), Go, DYb, prodlog(lXo, Xo, lY, Y),
                                               F1 ? preO[PRE_DB] : dilog(Xob));
                const double beta0 = -R11_3 * CA;
                const double piece1 = (CF * CF) * ((R4_3 - 2.0 * ZETA2) * QQf + 3.0 * X / Y
                    - (1.5 * QQf - 0.5 * X / Yo) * lr
                    - (QQf - QQfbar) * prodlog(lr, r, lomr, omr)
                    + (QQf + 0.5 * X / Yo) * L2r);
                const double piece2 = -0.5 * CF * beta0 * (R5_3 * QQf + X / Y + QQf * lr);
                const double base3 = R4_3 * QQf + 2.0 * X / Y + QQh;
                const double p12 = piece1 * rho + piece2 * rho;
                if (MASK & bit(K_NSP)) {
                    const double piece3 = -CF * (CF - 0.5 * CA) * (base3 - 1.0 * QQhbar);
                    out[K_NSP] = p12 + piece3 * rho + p4;
                }
                if (MASK & bit(K_NSM)) {
                    const double piece3 = -CF * (CF - 0.5 * CA) * (base3 - (-1.0) * QQhbar);
                    out[K_NSM] = p12 + piece3 * rho + p4;
                }
            }
        }
    }

    // ---------------- singlet qq (kernels_nlo.f90:244-301) ----------------
    if (V1QQ) {
        const double piece1 = X * (3.0 - 8.0 * X * Yb) / Y + X * (5.0 - 8.0 * X) / Yb * lr
            + (X / Yb - 4.0 * X * Xb) * L2r
            + 8.0 * X * Xb * (DXb - DYb + LPXbY);
        const double piece2 = -R290_9 * X * Xb
            - (X * (5.0 - 8.0 * X) / Yb + 2.0 * X * (9.0 - 19.0 * Xb) / 3.0) * lX
            - (X / Yb - 4.0 * X * Xb) * L2X
            + 4.0 * X * Xb * prodlog(lX, X, lXb, Xb);
        out[K_V1QQ] = 2.0 * CF * TF * (piece1 * rho + piece2 * erbl);
    }
    if (A1QQ) {
        const double piece1 = X / Yb * (lr - L2r);
        const double piece2 = -X / Yb * (lX - L2X);
        out[K_A1QQ] = -2.0 * CF * TF * (piece1 * rho + piece2 * erbl);
    }

    // ---------------- qG (kernels_nlo.f90:321-432) ----------------
    if (V1QG || A1QG) {
        const double QGfV = qg_fV(X, Y), QGfA = qg_fA(X, Y);
        const double QGfbarV = qg_fV(Xb, Yb), QGfbarA = qg_fA(Xb, Yb);
        const double L2yx1 = sqlog(lyx1, yx1);
        if (V1QG) {
            const double QGhV = h_mm(QGfV, QGfbarV, LPXbY, DS);
            const double QGhbarV = hbar_mm(qg_fV(Xb, Y), qg_fV(Xbb, Yb), G, DYb, LPXbY, DXbb);
            const double piece1 = TF * CF * (
                2.0 * (5.0 - 2.0 * ZETA2) * QGfV - 4.0 * QGfA + X / (Y * Yb)
                + 2.0 * (QGfV + QGfbarV + (2.0 * X - 2.0 * Y + X * Y) / (2.0 * Y * (Yb * Yb))) * lr
                - 2.0 * (QGfV + QGfbarV + 1.0 / (Y * Yb)) * lomr
                + QGfbarA * L2r + (QGfV + QGfbarV) * L2yx1);
            const double piece2 = TF * CA * (
                6.0 * X * (3.0 - 4.0 * X) / Yb - 2.0 * X * (11.0 - 16.0 * X) / Y + 4.0 * X * (1.0 - 3.0 * X) / (Y * Y)
                - 2.0 * (QGfV - 5.0 * QGfbarV + (5.0 - 7.0 * X) / (Yb * Yb)
                         + 2.0 * X * (3.0 - 2.0 * X - 12.0 * Xb * Y) / (Y * Yb)) * lr
                + 2.0 * (QGfV + QGfbarV + 1.0 / (Y * Yb)) * lomr
                + (QGfV - QGfbarV + (1.0 - 4.0 * X) / (Yb * Yb)) * L2r
                - (QGfV + QGfbarV) * (lomr * lomr) - QGhV + QGhbarV);
            const double piece3 = TF * CF * (
                -2.0 * (QGfV + QGfbarV + QGfbarA + X * (Y + 4.0 * (Yb * Yb)) / (2.0 * Y * (Yb * Yb))) * lX
                - 2.0 * QGfV * prodlog(lXb, Xb, lX, X)
                + (QGfV - QGfbarV - QGfbarA) * L2X);
            const double piece4 = TF * CA * (
                2.0 * (QGfV - 5.0 * QGfbarV + (5.0 - 7.0 * X) / (Yb * Yb)
                       + X * (5.0 - 6.0 * X - 22.0 * Y + 28.0 * X * Y) / (Y * Yb)) * lX
                - (QGfV - QGfbarV + (1.0 - 4.0 * X) / (Yb * Yb)) * L2X);
            out[K_V1QG] = piece1 * rho + piece2 * rho + piece3 * erbl + piece4 * erbl;
        }
        if (A1QG) {
            const double QGhA = h_mm(QGfA, QGfbarA, LPXbY, DS);
            const double QGhbarA = hbar_mm(qg_fA(Xb, Y), qg_fA(Xbb, Yb), G, DYb, LPXbY, DXbb);
            const double piece1 = TF * CF * (
                2.0 * (5.0 - 2.0 * ZETA2) * QGfA + 5.0 * X / (Y * Yb)
                + (2.0 * QGfA - 3.0 * X / (Yb * Yb) + 2.0 * X / (Y * Yb)) * lr
                - 2.0 * (QGfA + QGfbarA + 1.0 / (Y * Yb)) * lomr
                - QGfbarA * L2r + (QGfA + QGfbarA) * L2yx1);
            const double piece2 = TF * CA * (
                -4.0 * QGfA - 6.0 * X / (Y * Yb) - 2.0 * (QGfA - 2.0 * X / (Yb * Yb)) * lr
                + 2.0 * (QGfA + QGfbarA + 1.0 / (Y * Yb)) * lomr
                + (QGfA - 3.0 * X / (Yb * Yb)) * L2r - (QGfA + QGfbarA) * sqlog(lomr, omr)
                - QGhA - QGhbarA);
            const double piece3 = TF * CF * (
                -QGfA * (2.0 * lX + 2.0 * prodlog(lX, X, lXb, Xb) - L2X)
                + 3.0 * X / (Yb * Yb) * lX);
            const double piece4 = TF * CA * (
                (6.0 * X * Yb - 4.0 * X * Y) / (Y * (Yb * Yb)) * lX
                + QGfA * (2.0 * lX - L2X) + 3.0 * X / (Yb * Yb) * L2X);
            out[K_A1QG] = piece1 * rho + piece2 * rho + piece3 * erbl + piece4 * erbl;
        }
    }

    // ---------------- Gq (kernels_nlo.f90:451-671) ----------------
    if (V1GQ || A1GQ || V1GQN || A1GQN) {
        const double GQfV = gq_fV(X, Y), GQfA = gq_fa(X, Y);
        const double GQfbarV = gq_fV(Xb, Yb), GQfbarA = gq_fa(Xb, Yb);
        const double L2omr = sqlog(lomr, omr);
        const double L2yx1 = sqlog(lyx1, yx1);
        if (V1GQ || V1GQN) {
            // the beta0 bracket is shared by KV1_Gq_half (piece2) and KV1_Gq_half_nfl
            const double br = R10_3 * GQfV + 2.0 * X * Xb / Y + (GQfV + GQfbarV) * lomr - GQfA * lX + GQfbarA * lXb;
            if (V1GQN) {
                const double beta0 = R4_3 * TF;
                out[K_V1GQN] = (-0.5 * CF * beta0 * br) * rho;
            }
            if (V1GQ) {
                const double beta0 = -R11_3 * CA;
                const double GQhV = h_mm(GQfV, GQfbarV, LPXbY, DS);
                const double GQhbarV = hbar_mm(gq_fV(Xb, Y), gq_fV(Xbb, Yb), G, DYb, LPXbY, DXbb);
                const double piece1 = (CF * CF) * (
                    3.0 * GQfA - 4.5 * X / Y - (0.5 * X * (2.0 + X) / Yb - X * Xb / (Y * Yb)) * lr
                    - (1.5 * GQfV + 1.5 * GQfbarV + X * Xb / (Y * Yb)) * lomr
                    - 0.5 * GQfA * L2r
                    - 0.5 * (GQfV + GQfbarV) * L2omr);
                const double piece2 = -0.5 * CF * beta0 * br;
                const double piece3 = CF * CA * (
                    (R4_9 - 2.0 * ZETA2) * GQfV - R26_9 * (X * X) / Y + X * (3.0 + 4.0 * (X * X)) / Y
                    - X / 9.0 * (105.0 - 246.0 * X + 188.0 * (X * X))
                    - (X * Xb / (Y * Yb) - X * (6.0 - 11.0 * X + 8.0 * (X * X)) / Yb + 4.0 * X * Xb * (Xb - X)) * lr
                    + X * Xb / (Y * Yb) * lomr
                    + 0.5 * (2.0 * GQfV + 2.0 * GQfA + GQfbarA - 3.0 * (X * X) / Yb - 2.0) * L2r
                    + 0.5 * (2.0 - GQfA - GQfbarA) * L2yx1
                    - 0.5 * GQhV - 0.5 * GQhbarV);
                const double piece4 = (CF * CF) * (
                    -0.5 * (X * (2.0 - 5.0 * X) / Y - 3.0 * (X * X) / Yb) * lX
                    + 0.5 * GQfA * L2X);
                const double piece5 = CF * CA * (
                    (X * Xb / (Y * Yb) - X * (6.0 - 11.0 * X + 8.0 * (X * X)) / Yb
                     + X * (18.0 - 63.0 * X + 62.0 * (X * X)) / 3.0) * lX
                    - GQfV * prodlog(lX, X, lXb, Xb) - (X * X) * (1.0 - 4.0 * Y) / (2.0 * Y * Yb) * L2X);
                out[K_V1GQ] = piece1 * rho + piece2 * rho + piece3 * rho + piece4 * erbl + piece5 * erbl;
            }
        }
        if (A1GQ || A1GQN) {
            const double br2 = -R2_3 * GQfA + 2.0 * X / Y + (GQfA + GQfbarA) * lomr;
            const double br5 = GQfA * lX;
            if (A1GQN) {
                const double beta0 = R4_3 * TF;
                out[K_A1GQN] = (-0.5 * CF * beta0 * br2) * rho + (-0.5 * CF * beta0 * br5) * erbl;
            }
            if (A1GQ) {
                const double beta0 = -R11_3 * CA;
                const double GQhA = h_mm(GQfA, GQfbarA, LPXbY, DS);
                const double GQhbarA = hbar_mm(gq_fa(Xb, Y), gq_fa(Xbb, Yb), G, DYb, LPXbY, DXbb);
                const double piece1 = (CF * CF) * (
                    -3.0 * X * (1.0 + X) / Y + ((X * X) * Y + 2.0 * X * (Xb - X) * Yb) / (Y * Yb) * lr
                    - (X * (Xb + Yb) + Xb * Y) / (Y * Yb) * lomr
                    + GQfA * L2r - (GQfA + GQfbarA) * L2omr);
                const double piece2 = -0.5 * CF * beta0 * br2;
                const double piece3 = CF * CA * (
                    (R13_3 - 2.0 * ZETA2) * GQfA - X / Y
                    - (X * (Xb - X) / Y + 5.0 * X * Xb / Yb) * lr
                    + ((Xb - 2.0 * (Xb * Xb)) / Yb + (X - 2.0 * (X * X)) / Y) * lomr
                    - 0.5 * (GQfbarA + 3.0 * (X * X) / Yb) * L2r + 0.5 * (GQfA + GQfbarA) * L2yx1
                    - 0.5 * GQhA + 0.5 * GQhbarA);
                const double piece4 = (CF * CF) * (
                    -((X * X) * Y + (1.0 - (Xb * Xb)) * Yb) / (Y * Yb) * lX
                    - GQfA * L2X);
                const double piece5 = -0.5 * CF * beta0 * br5;
                const double piece6 = CF * CA * (
                    (5.0 * X * Xb / Yb + X / Y) * lX
                    + GQfA * (-2.0 * lX + 0.5 * L2X - prodlog(lX, X, lXb, Xb))
                    + 1.5 * (X * X) / Yb * L2X);
                out[K_A1GQ] = piece1 * rho + piece2 * rho + piece3 * rho + piece4 * erbl + piece5 * erbl + piece6 * erbl;
            }
        }
    }

    // ---------------- GG (kernels_nlo.f90:709-982) ----------------
    if ((V1GG || V1GGN) && !WITH_RHO) {
        if (V1GG) {
            const double piece4 = (CA * CA) * (
                -2.0 * (X * X) / (Yb * Yb) * L2X
                + (Xb * ((X * X) + (Xb * Xb) - 2.0 * Xb * (1.0 + 2.0 * X) * Yb) + Yb) / (Yb * Yb) * prodlog(lX, X, lXb, Xb)
                + ((3.0 * X * ((X * X) + (Xb * Xb)) * (Yb - Y) - 4.0 * (X * X) * Y) / (Y * Yb)
                   - X * (2.0 - 3.0 * X - 8.0 * (Xb * Xb)) / (Yb * Yb)) * lX);
            out[K_V1GG] = piece4 * erbl;
        }
        if (V1GGN) {
            const double piece5 = CF * TF * (
                -2.0 * X * (1.0 - 3.0 * X - 2.0 * X * Yb) / (Yb * Yb) * lX + (X * X) / (Yb * Yb) * L2X);
            out[K_V1GGN] = piece5 * erbl;
        }
    }
    if ((V1GG || V1GGN) && WITH_RHO) {
        const double GGfV = gg_fV(X, Y);
        const double GGfc = gg_fc(X, Y);
        // the beta0 bracket shared by KV1_GG_half (piece2) and KV1_GG_half_nfl (piece2)
        const double br = -R5_3 * GGfV - R13_3 * GGfc - R11_2 * (X * X) / (Y * Y) + X * (X + Yb) / ((Y * Y) * Yb)
            + (X * X) / (Yb * Yb) * lY + (Xb * Xb) / (Y * Y) * lXb;
        if (V1GG) {
            const double beta0 = -R11_3 * CA;
            const double GGfbarV = gg_fV(Xb, Yb);
            const double GGhV = h_pp(GGfV, GGfbarV, LPXbY, DS);
            const double GGhbarV = hbar_pp(gg_fV(Xb, Y), gg_fV(Xbb, Yb), G, DYb, LPXbY, DXbb);
            const double piece1 = (CA * CA) * (
                (R2_3 - 2.0 * ZETA2) * GGfV - 0.5 * GGfc + X * (4.0 + 5.0 * X) / (4.0 * (Y * Y))
                - X * (10.0 - 19.0 * X + 16.0 * (X * X)) / Y + X * (6.0 - 13.0 * X + 8.0 * (X * X)) / Yb
                + (2.0 * X * (1.0 - 2.0 * Xb * (X * Y + Xb * Yb)) / (Y * Yb)
                   + X * (2.0 - 3.0 * X - 8.0 * (Xb * Xb)) / (Yb * Yb)) * lr
                + (GGfV + 2.0 * (X * X) / (Yb * Yb)) * L2r
                - (GGfV - GGfbarV) * prodlog(lr, r, lomr, omr)
                - 0.5 * GGhV - 0.5 * GGhbarV);
            const double piece2 = 0.5 * CA * beta0 * br;
            const double piece4 = (CA * CA) * (
                -2.0 * (X * X) / (Yb * Yb) * L2X
                + (Xb * ((X * X) + (Xb * Xb) - 2.0 * Xb * (1.0 + 2.0 * X) * Yb) + Yb) / (Yb * Yb) * prodlog(lX, X, lXb, Xb)
                + ((3.0 * X * ((X * X) + (Xb * Xb)) * (Yb - Y) - 4.0 * (X * X) * Y) / (Y * Yb)
                   - X * (2.0 - 3.0 * X - 8.0 * (Xb * Xb)) / (Yb * Yb)) * lX);
            out[K_V1GG] = piece1 * rho + piece2 * rho + 0.0 + piece4 * erbl + 0.0;
        }
        if (V1GGN) {
            const double beta0 = R4_3 * TF;
            const double piece2 = 0.5 * CA * beta0 * br;
            const double piece3 = CF * TF * (
                -R20_3 * GGfc - 12.0 * (X * X) / (Y * Y) + 2.0 * X * (2.0 + Y + 3.0 * X * Y - 4.0 * (Y * Y)) / ((Y * Y) * Yb)
                - 2.0 * X * (X - Y + 2.0 * X * Y) / (Y * (Yb * Yb)) * lr - (X * X) / (Yb * Yb) * L2r);
            const double piece5 = CF * TF * (
                -2.0 * X * (1.0 - 3.0 * X - 2.0 * X * Yb) / (Yb * Yb) * lX + (X * X) / (Yb * Yb) * L2X);
            out[K_V1GGN] = 0.0 + piece2 * rho + piece3 * rho + 0.0 + piece5 * erbl;
        }
    }
    if ((A1GG || A1GGN) && !WITH_RHO) {
        if (A1GG) {
            const double LPX = prodlog(lX, X, lXb, Xb);
            const double piece4 = (CA * CA) * (
                (3.0 * (Xb - X) * X / (Y * Yb) - 6.0 * X * Xb / (Yb * Yb) + (X * X) / (Yb * Yb)) * lX
                - 2.0 * (X * X) / (Yb * Yb) * (L2X + LPX) + (3.0 * X - Y) / (Yb * Yb) * LPX);
            out[K_A1GG] = piece4 * erbl;
        }
        if (A1GGN) {
            const double piece5 = CF * TF * (X / (Yb * Yb) * (2.0 * lX + X * L2X));
            out[K_A1GGN] = piece5 * erbl;
        }
    }
    if ((A1GG || A1GGN) && WITH_RHO) {
        const double GGfA = gg_fA(X, Y);
        const double br = -R5_3 * GGfA - R5_2 * (X * X) / (Y * Y) + X * (Yb + X * Y) / ((Y * Y) * Yb)
            + (X * X) / (Yb * Yb) * lY + (Xb * Xb) / (Y * Y) * lXb;
        if (A1GG) {
            const double beta0 = -R11_3 * CA;
            const double GGfbarA = gg_fA(Xb, Yb);
            const double GGhA = h_pp(GGfA, GGfbarA, LPXbY, DS);
            const double GGhbarA = hbar_pp(gg_fA(Xb, Y), gg_fA(Xbb, Yb), G, DYb, LPXbY, DXbb);
            const double piece1 = (CA * CA) * (
                (R2_3 - 2.0 * ZETA2) * GGfA - R15_4 * (X * X) / (Y * Y) - 6.0 * X * Xb / (Y * Yb)
                + X * (X + Yb) / ((Y * Y) * Yb)
                - (2.0 * X / (Y * Yb) - 6.0 * X * Xb / (Yb * Yb) + (X * X) / (Yb * Yb)) * lr
                + (GGfA + 2.0 * (X * X) / (Yb * Yb)) * L2r - (GGfA - GGfbarA) * prodlog(lr, r, lomr, omr)
                - 0.5 * GGhA + 0.5 * GGhbarA);
            const double piece2 = 0.5 * CA * beta0 * br;
            const double LPX = prodlog(lX, X, lXb, Xb);
            const double piece4 = (CA * CA) * (
                (3.0 * (Xb - X) * X / (Y * Yb) - 6.0 * X * Xb / (Yb * Yb) + (X * X) / (Yb * Yb)) * lX
                - 2.0 * (X * X) / (Yb * Yb) * (L2X + LPX) + (3.0 * X - Y) / (Yb * Yb) * LPX);
            out[K_A1GG] = piece1 * rho + piece2 * rho + 0.0 + piece4 * erbl + 0.0;
        }
        if (A1GGN) {
            const double beta0 = R4_3 * TF;
            const double piece2 = 0.5 * CA * beta0 * br;
            const double piece3 = CF * TF * (
                2.0 * X * (1.0 - 2.0 * X * Yb) / ((Y * Y) * Yb) + 2.0 * (X * X) / (Y * Yb) * lr
                - X / (Yb * Yb) * (2.0 * lr + X * L2r));
            const double piece5 = CF * TF * (X / (Yb * Yb) * (2.0 * lX + X * L2X));
            out[K_A1GGN] = 0.0 + piece2 * rho + piece3 * rho + 0.0 + piece5 * erbl;
        }
    }
}

// ---------------------------------------------------------------------------------------
// analytic LO constants (kernels_lo.f90:64-79, 171-207)

TKB_DEV double K0_qq_cst(double x, double xi)
{
    double K;
    if (x > xi)
        K = 1.5 + 2.0 * log_(1.0 - x)
            + 0.5 * ((x - xi) * log_((x - xi) * (1.0 + xi)) - (x + xi) * log_((x + xi) * (1.0 - xi))) / xi;
    else if (x < xi && x > -xi)
        K = 1.5 + log_((1.0 - x * x) / (1.0 + xi)) + 0.5 * ((x - xi) * log_(xi - x) - (x + xi) * log_(x + xi)) / xi;
    else if (x < -xi)
        K = 1.5 + 2.0 * log_(1.0 + x)
            + 0.5 * ((-x - xi) * log_((-x - xi) * (1.0 + xi)) - (-x + xi) * log_((-x + xi) * (1.0 - xi))) / xi;
    else
        K = 0.0;
    return K * CF;
}

TKB_DEV double KA0_GG_cst(double x, double xi)
{
    double K;
    if (x > xi)
        K = -2.0 * x * (1.0 - x) / (1.0 - xi * xi) + log_(((1.0 - x) * (1.0 - x)) / (1.0 - xi * xi));
    else if (x < -xi)
        K = 2.0 * x * (1.0 + x) / (1.0 - xi * xi) + log_(((1.0 + x) * (1.0 + x)) / (1.0 - xi * xi));
    else
        K = -2.0 * (x * x) / (1.0 + xi) / xi + log_((1.0 - x * x) / ((1.0 + xi) * (1.0 + xi)));
    K = K * CA;
    const double beta0 = -R11_3 * CA;
    return K - beta0 / 2.0;
}

TKB_DEV double KA0_GG_nfl_cst()
{
    const double beta0 = R4_3 * TF;
    return -beta0 / 2.0;
}

// explicit delta-function constants of the NLO kernels (kernels_nlo.f90:68, 858, 932);
// KV1_NSm, KV1_NS*_nfl, KV1_GG have none and KA1_GG's is overwritten in the reference (:777-779)
TKB_DEV double nlo_explicit_cst(int c)      // c: 0 NSp 1 NSm 2 NS_nfl 3 VGG 4 AGG 5 VGG_nfl 6 AGG_nfl
{
    switch (c) {
    case 0: return CF * (CF - 0.5 * CA) * (6.5 - 6.0 * ZETA2 + 4.0 * ZETA3);
    case 5: return -R1_108 * (35.0 * CA + 74.0 * CF);
    case 6: return R1_54 * (29.0 * CA - 28.0 * CF);
    }
    return 0.0;
}

}  // namespace tkb
