// qags_step.cuh -- per-instance state of a QUADPACK QAGS run and its bookkeeping steps (integration.f90:246-611)
// as device functions.  Included by assemble.cu after the scalar routines (qsort_, qsort_bs_, qextr_, qk21_reduce).
//
// One QAGS instance = one sub-integral of one NLO plus-prescription constant (adaptive_integrate,
// integration.f90:51-64).  Its scalar state lives in a QSlot, the first QHEAD entries of its interval lists in a
// QHead (both in shared memory), the tail of the lists in a QLists record in global memory (the reference's limit
// is 500 intervals).  qags_first consumes the 21 integrand values of the first approximation, qags_step the 2 x 21
// values of one bisection; each returns QS_CONTINUE with the next request (a1, b1, a2, b2) in the slot, or QS_DONE
// with result / last / ifail.  Every comparison is QUADPACK's, in its order, so the bisection path is the reference's.
#pragma once

namespace tkb {

#ifndef TKB_QHEAD
#define TKB_QHEAD 48
#endif
constexpr int QHEAD = TKB_QHEAD;            // list entries (1-based index < QHEAD) kept in shared memory

struct QSlot {                       // scalar QUADPACK state of one instance (integration.f90:374-381)
    double x, xi, xx1, xx2, a, b;
    double a1, b1, a2, b2;
    double area, errsum, errmax, erlarg, erlast, ertest, errbnd, small, correc, result, abserr, defabs;
    double res3la[4];
    unsigned long long inst;
    int maxerr, nrmax, nres, numrl2, ktmin, iroff1, iroff2, iroff3, ierro, ifail, ksgn, last;
    int extrap, noext;
    int nanl;                        // an error estimate was a NaN: the ordered list may be unsorted (see qsort_bs_)
    int nint;                        // 0 idle, 1 first qk21 pending, 2 bisection pending
    int defer;                       // qags_step<true> posted the next request before the extrapolation: qags_deferred is due
    int term;                        // the deferred extrapolation ended the integration: result / last / ifail are final
    double erlarg_pre;               // erlarg as the extrapolation step sees it (the posted state already carries errsum)
};

struct QLists {                      // per-slot interval lists in global memory (1-based)
    double alist[QLIMIT + 1], blist[QLIMIT + 1], rlist[QLIMIT + 1], elist[QLIMIT + 1];
    double rlist2[56];
    int iord[QLIMIT + 3];
};

struct QHead {                        // shared-memory head of one instance's interval lists
    double alist[QHEAD], blist[QHEAD], rlist[QHEAD], elist[QHEAD];
    double rlist2[56];
    int iord[QHEAD];
};
using BDList = SpillList<double, QHEAD>;
using BIList = SpillList<int, QHEAD>;

// outcome of one bookkeeping step
enum { QS_CONTINUE = 0, QS_DONE = 1 };

// first approximation (integration.f90:398-441): consumes fv[0..20]
static __device__ __noinline__ int qags_first(QSlot& S, QHead& H, QLists& L, const double* fv)
{
    const BDList alist{ H.alist, L.alist }, blist{ H.blist, L.blist }, rlist{ H.rlist, L.rlist }, elist{ H.elist, L.elist };
    const BIList iord{ H.iord, L.iord };
    double* rlist2 = H.rlist2;
    const double epsabs = 1.0e-12, epsrel = 1.0e-06;
    double result, abserr, defabs, resabs;
    alist[1] = S.a; blist[1] = S.b; rlist[1] = 0.0; elist[1] = 0.0;
    qk21_reduce(fv, S.a, S.b, result, abserr, defabs, resabs);
    const double dres = fabs(result);
    const double errbnd = fmax2(epsabs, epsrel * dres);
    S.last = 1;
    S.nanl = 0;
    S.defer = 0;
    S.term = 0;
    rlist[1] = result;
    elist[1] = abserr;
    iord[1] = 1;
    S.ifail = 0;
    if (abserr <= 1.0e+02 * eps_() * defabs && abserr > errbnd) S.ifail = 2;
    S.result = result;
    S.abserr = abserr;
    S.defabs = defabs;
    if (S.ifail != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) return QS_DONE;   // goto 140
    rlist2[1] = result;
    S.errmax = abserr;
    S.maxerr = 1;
    S.area = result;
    S.errsum = abserr;
    S.abserr = huge_();
    S.nrmax = 1; S.nres = 0; S.numrl2 = 2; S.ktmin = 0; S.extrap = 0; S.noext = 0;
    S.iroff1 = 0; S.iroff2 = 0; S.iroff3 = 0; S.ierro = 0;
    S.correc = 0; S.erlarg = 0; S.small = 0; S.ertest = 0; S.errbnd = errbnd; S.erlast = 0;
    S.ksgn = (dres >= (1.0 - 5.0e+01 * eps_()) * defabs) ? 1 : -1;
    S.last = 2;                                      // the loop variable of `do last = 2, limit`
    return QS_CONTINUE;
}

// final result selection (integration.f90:579-607); status 1 = Fortran `exit`, 2 = goto 115
static __device__ __noinline__ void qags_finish(QSlot& S, QHead& H, QLists& L, int status)
{
    const BDList rlist{ H.rlist, L.rlist };
    double result = S.result, abserr = S.abserr;
    int ifail = S.ifail;
    const double area = S.area, errsum = S.errsum;
    if (status == 2) goto L115;
    if (abserr == huge_()) goto L115;
    if (ifail + S.ierro == 0) goto L110;
    if (S.ierro == 3) abserr = abserr + S.correc;
    if (ifail == 0) ifail = 3;
    if (result != 0.0 && area != 0.0) goto L105;
    if (abserr > errsum) goto L115;
    if (area == 0.0) goto L130;
    goto L110;
L105:
    if (abserr / fabs(result) > errsum / fabs(area)) goto L115;
L110:
    if (S.ksgn == (-1) && fmax2(fabs(result), fabs(area)) <= S.defabs * 1.0e-02) goto L130;
    if (1.0e-02 > (result / area) || (result / area) > 1.0e+02 || errsum > fabs(area)) ifail = 6;
    goto L130;
L115:
    result = 0.0;
    for (int k = 1; k <= S.last; k++) result = result + rlist[k];
    abserr = errsum;
L130:
    if (ifail > 2) ifail = ifail - 1;
    S.result = result;
    S.abserr = abserr;
    S.ifail = ifail;
}

// one pass of the `do last = 2, limit` body after its two qk21 calls (integration.f90:455-577).
// (area1, error1, defab1), (area2, error2, defab2) are the qk21 results of the request (a1,b1), (a2,b2).
// DEFER: the epsilon-algorithm step (qextr, integration.f90:540-541: the longest piece of the bookkeeping chain, ~9 k
// cycles) is taken off the path to the next request.  Everything QUADPACK does after it -- restart from the interval
// with the largest error, halve `small`, erlarg = errsum -- does not depend on its numbers unless it ENDS the
// integration, so that state is set and the request posted first; qags_deferred then runs the extrapolation and its
// tests while the evaluators work on the request, and if it ends the integration the values of that one request are
// dropped (S.term).  Same operations on the same numbers in QUADPACK's order per instance: result, last, ifail unchanged.
template <bool DEFER = false>
static __device__ __noinline__ int qags_step(QSlot& S, QHead& H, QLists& L, double area1, double error1, double defab1,
                                             double area2, double error2, double defab2)
{
    const double epsabs = 1.0e-12, epsrel = 1.0e-06;
    const int limit = QLIMIT;
    const BDList alist{ H.alist, L.alist }, blist{ H.blist, L.blist }, rlist{ H.rlist, L.rlist }, elist{ H.elist, L.elist };
    const BIList iord{ H.iord, L.iord };
    double* rlist2 = H.rlist2;
    // scalars in registers for the duration of the step
    int maxerr = S.maxerr, nrmax = S.nrmax, nres = S.nres, numrl2 = S.numrl2, ktmin = S.ktmin, iroff1 = S.iroff1,
        iroff2 = S.iroff2, iroff3 = S.iroff3, ierro = S.ierro, ifail = S.ifail;
    const int last = S.last;
    bool extrap = S.extrap != 0, noext = S.noext != 0;
    double area = S.area, errsum = S.errsum, errmax = S.errmax, erlarg = S.erlarg, ertest = S.ertest, errbnd = S.errbnd,
           small = S.small, correc = S.correc, result = S.result, abserr = S.abserr;
    const double erlast = S.erlast, a1 = S.a1, b1 = S.b1, a2 = S.a2, b2 = S.b2;
    double area12, erro12, abseps, reseps;
    int id, jupbnd, k, status = 0;

    area12 = area1 + area2;
    erro12 = error1 + error2;
    errsum = errsum + erro12 - errmax;
    area = area + area12 - rlist[maxerr];
    if (defab1 == error1 || defab2 == error2) goto L15;
    if (fabs(rlist[maxerr] - area12) > (double)1.0e-05f * fabs(area12) || erro12 < 9.9e-01 * errmax) goto L10;
    if (extrap) iroff2 = iroff2 + 1; else iroff1 = iroff1 + 1;
L10:
    if (last > 10 && erro12 > errmax) iroff3 = iroff3 + 1;
L15:
    rlist[maxerr] = area1;
    rlist[last] = area2;
    errbnd = fmax2(epsabs, epsrel * fabs(area));
    if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ifail = 2;
    if (iroff2 >= 5) ierro = 3;
    if (last == limit) ifail = 1;
    if (fmax2(fabs(a1), fabs(b2)) <= (1.0 + 1.0e+03 * eps_()) * (fabs(a2) + 1.0e+03 * tiny_())) ifail = 4;
    if (error2 <= error1) {
        alist[last] = a2;
        blist[maxerr] = b1;
        blist[last] = b2;
        elist[maxerr] = error1;
        elist[last] = error2;
    } else {
        alist[maxerr] = a2;
        alist[last] = a1;
        blist[last] = b1;
        rlist[maxerr] = area2;
        rlist[last] = area1;
        elist[maxerr] = error2;
        elist[last] = error1;
    }
#ifdef TKB_K2_STATS
    long long tq0_ = clock64();
#endif
    if (!(error1 == error1) || !(error2 == error2)) S.nanl = 1;
    if (S.nanl) qsort_(limit, last, maxerr, errmax, elist, iord, nrmax);
    else qsort_bs_(limit, last, maxerr, errmax, elist, iord, nrmax);
#ifdef TKB_K2_STATS
    atomicAdd(&g_k2stats[8], (unsigned long long)(clock64() - tq0_)); atomicAdd(&g_k2stats[9], 1ULL);
#endif
    if (errsum <= errbnd) { status = 2; goto Lend; }
    if (ifail != 0) { status = 1; goto Lend; }
    if (last == 2) goto L80;
    if (noext) goto Lend;
    erlarg = erlarg - erlast;
    if (fabs(b1 - a1) > small) erlarg = erlarg + erro12;
    if (extrap) goto L40;
    if (fabs(blist[maxerr] - alist[maxerr]) > small) goto Lend;
    extrap = true;
    nrmax = 2;
L40:
    if (ierro != 3 && erlarg > ertest) {
        id = nrmax;
        jupbnd = last;
        if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
        for (k = id; k <= jupbnd; k++) {
            maxerr = iord[nrmax];
            errmax = elist[maxerr];
            if (fabs(blist[maxerr] - alist[maxerr]) > small) goto Lend;
            nrmax = nrmax + 1;
        }
    }
    numrl2 = numrl2 + 1;
    rlist2[numrl2] = area;
    if (DEFER) {
        S.defer = 1;
        S.erlarg_pre = erlarg;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = false;
        small = small * 5.0e-01;
        erlarg = errsum;
        goto Lend;
    }
#ifdef TKB_K2_STATS
    tq0_ = clock64();
#endif
    qextr_(numrl2, rlist2, reseps, abseps, S.res3la, nres);
#ifdef TKB_K2_STATS
    atomicAdd(&g_k2stats[10], (unsigned long long)(clock64() - tq0_)); atomicAdd(&g_k2stats[11], 1ULL);
#endif
    ktmin = ktmin + 1;
    if (ktmin > 5 && abserr < 1.0e-03 * errsum) ifail = 5;
    if (abseps < abserr) {
        ktmin = 0;
        abserr = abseps;
        result = reseps;
        correc = erlarg;
        ertest = fmax2(epsabs, epsrel * fabs(reseps));
        if (abserr <= ertest) { status = 1; goto Lend; }
    }
    if (numrl2 == 1) noext = true;
    if (ifail == 5) { status = 1; goto Lend; }
    maxerr = iord[1];
    errmax = elist[maxerr];
    nrmax = 1;
    extrap = false;
    small = small * 5.0e-01;
    erlarg = errsum;
    goto Lend;
L80:
    small = fabs(S.b - S.a) * 3.75e-01;
    erlarg = errsum;
    ertest = errbnd;
    rlist2[2] = area;
Lend:
    S.maxerr = maxerr; S.nrmax = nrmax; S.nres = nres; S.numrl2 = numrl2; S.ktmin = ktmin; S.iroff1 = iroff1;
    S.iroff2 = iroff2; S.iroff3 = iroff3; S.ierro = ierro; S.ifail = ifail;
    S.extrap = extrap ? 1 : 0; S.noext = noext ? 1 : 0;
    S.area = area; S.errsum = errsum; S.errmax = errmax; S.erlarg = erlarg; S.ertest = ertest; S.errbnd = errbnd;
    S.small = small; S.correc = correc; S.result = result; S.abserr = abserr;
    if (status != 0) {
        qags_finish(S, H, L, status);
        return QS_DONE;
    }
    if (last == limit) {              // cannot happen (last == limit sets ifail = 1), kept as a guard
        qags_finish(S, H, L, 1);
        return QS_DONE;
    }
    // next pass of the loop: bisect the interval with the nrmax-th largest error (integration.f90:447-452)
    S.last = last + 1;
    S.a1 = alist[maxerr];
    S.b1 = 5.0e-01 * (alist[maxerr] + blist[maxerr]);
    S.a2 = S.b1;
    S.b2 = blist[maxerr];
    S.erlast = errmax;
    return QS_CONTINUE;
}

// the extrapolation step that qags_step<true> left behind (integration.f90:540-556), run after the next request is posted
static __device__ __noinline__ void qags_deferred(QSlot& S, QHead& H, QLists& L)
{
    const double epsabs = 1.0e-12, epsrel = 1.0e-06;
    int numrl2 = S.numrl2, nres = S.nres, ktmin = S.ktmin, ifail = S.ifail;
    double abserr = S.abserr, result = S.result, correc = S.correc, ertest = S.ertest;
    const double errsum = S.errsum;
    double reseps, abseps;
    int status = 0;
#ifdef TKB_K2_STATS
    const long long tq0_ = clock64();
#endif
    qextr_(numrl2, H.rlist2, reseps, abseps, S.res3la, nres);
#ifdef TKB_K2_STATS
    atomicAdd(&g_k2stats[10], (unsigned long long)(clock64() - tq0_)); atomicAdd(&g_k2stats[11], 1ULL);
#endif
    ktmin = ktmin + 1;
    if (ktmin > 5 && abserr < 1.0e-03 * errsum) ifail = 5;
    if (abseps < abserr) {
        ktmin = 0;
        abserr = abseps;
        result = reseps;
        correc = S.erlarg_pre;
        ertest = fmax2(epsabs, epsrel * fabs(reseps));
        if (abserr <= ertest) status = 1;
    }
    if (status == 0) {
        if (numrl2 == 1) S.noext = 1;
        if (ifail == 5) status = 1;
    }
    S.numrl2 = numrl2; S.nres = nres; S.ktmin = ktmin; S.ifail = ifail;
    S.abserr = abserr; S.result = result; S.correc = correc; S.ertest = ertest;
    if (status != 0) {
        S.last = S.last - 1;                 // the step that ran the extrapolation, not the request posted after it
        qags_finish(S, H, L, status);
        S.term = 1;
    }
}

}  // namespace tkb
