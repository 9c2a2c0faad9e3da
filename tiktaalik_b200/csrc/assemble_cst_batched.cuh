// assemble_cst_batched.cuh -- K2 in its batched form: one persistent block per SM advances QB QAGS
// instances in lock-step rounds.  Included by assemble.cu (after the QUADPACK helpers).
//
// Why: with one instance per warp (the first design, kept below it for comparison) the SM's
// instruction cache thrashes -- at any moment its warps are spread over integrand evaluation,
// classification and QUADPACK bookkeeping, about 100 KB of code; ncu showed a 66 % hit rate and the GPC
// instruction cache at 65 % of its request rate -- the bookkeeping runs on one lane of 32, and the last
// evaluation pass of every step is partly empty.  Here a round has uniform phases for the whole block:
//   S  one thread per instance: consume the 2 x 21 integrand values of its last request (qk21 sums,
//      QUADPACK's bisection bookkeeping: integration.f90:404-601), issue the next request or finish
//      and pull a new instance from the global queue;
//   N  node arguments for all pending requests;
//   C  support class of every (instance, node, direction, half) task, block-wide ballot compaction;
//   E  evaluation of the live half kernels, all threads, full passes;
//   K  K(x,y) - K(y,x) per node.
// The arithmetic of every instance is unchanged (same qk21 sums in the same order, same comparisons),
// so results and bisection paths are identical to the per-warp version and to the oracle.
#pragma once

namespace tkb {

#ifndef TKB_QB
#define TKB_QB 16
#endif
#ifndef TKB_QT
#define TKB_QT 384
#endif
constexpr int QB = TKB_QB;           // instances per block
constexpr int QT = TKB_QT;           // threads per block
#ifndef TKB_QBPS
#define TKB_QBPS 2
#endif
constexpr int QBLOCKS_PER_SM = TKB_QBPS;   // blocks per SM run out of phase: one block's bookkeeping overlaps the others' evaluation
constexpr int QHEAD = 48;            // list entries (1-based index < QHEAD) kept in shared memory

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
};

struct QLists {                      // per-slot interval lists in global memory (1-based)
    double alist[QLIMIT + 1], blist[QLIMIT + 1], rlist[QLIMIT + 1], elist[QLIMIT + 1];
    double rlist2[56];
    int iord[QLIMIT + 3];
};
size_t qags_lists_bytes() { return sizeof(QLists) * QB; }

struct QHead {                        // shared-memory head of one instance's interval lists
    double alist[QHEAD], blist[QHEAD], rlist[QHEAD], elist[QHEAD];
    double rlist2[56];
    int iord[QHEAD];
};
using BDList = SpillList<double, QHEAD>;
using BIList = SpillList<int, QHEAD>;

struct QBatchShared {
    QSlot slot[QB];
    QHead head[QB];
    double ny[QB][42][2];
    double hv[QB][168];
    double fv[QB][42];
    unsigned short live[QB * 168];
    int nfull, nerbl;
};

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

// node_is_safe / task_class: tkb_kernels.cuh

template <int KID>
__global__ void __launch_bounds__(QT, QBLOCKS_PER_SM)
assemble_cst_batched_kernel(int c, int n, int nxi, const double* __restrict__ xi_arr, int grid_type,
                            QLists* __restrict__ lists, unsigned long long* __restrict__ counter,
                            double* __restrict__ cst_res, int* __restrict__ cst_last, int* __restrict__ cst_ifail)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QBatchShared& sh = *reinterpret_cast<QBatchShared*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    constexpr bool HAS_ERBL = (KID != K_NSN);
    bool exhausted = false;
    if (tid < QB) sh.slot[tid].nint = 0;
    __syncthreads();

#ifdef TKB_K2_STATS
    long long tq_ = clock64();
#define QB_ACC(slot) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_k2stats[slot], (unsigned long long)(now_ - tq_)); tq_ = now_; } } while (0)
#else
#define QB_ACC(slot)
#endif
    for (;;) {
        // ---------------- S: bookkeeping ----------------
        // Instances are spread over all warps (QB / NW per warp) so that the latency-bound scalar code of
        // different instances overlaps; each instance uses a lane pair, one qk21 reduction per lane.
        {
            constexpr int NW = QT / 32, SLW = (QB + NW - 1) / NW;      // warps, slots per warp
            const int w = tid >> 5;
            const int cand = (lane >> 1) * NW + w;
            const bool mine = (lane < 2 * SLW) && (cand < QB);
            const int myslot = mine ? cand : 0;
            const int part = lane & 1;
            QSlot& S = sh.slot[myslot];
            QLists* myL = lists + (size_t)blockIdx.x * QB + myslot;
            const int nint0 = mine ? S.nint : 0;
            double r_ = 0, e_ = 0, ra_ = 0, rc_ = 0;
            if (nint0 == 2) qk21_reduce(sh.fv[myslot] + 21 * part, part ? S.a2 : S.a1, part ? S.b2 : S.b1, r_, e_, ra_, rc_);
            const double area2 = __shfl_down_sync(0xffffffffu, r_, 1);
            const double error2 = __shfl_down_sync(0xffffffffu, e_, 1);
            const double defab2 = __shfl_down_sync(0xffffffffu, rc_, 1);
            if (mine && part == 0) {
                int state = QS_DONE;
                if (nint0 == 1) {
                    state = qags_first(S, sh.head[myslot], *myL, sh.fv[myslot]);
                    if (state == QS_CONTINUE) {
                        S.a1 = S.a;
                        S.b1 = 5.0e-01 * (S.a + S.b);
                        S.a2 = S.b1;
                        S.b2 = S.b;
                        S.erlast = S.errmax;
                        S.nint = 2;
                    }
                } else if (nint0 == 2) {
                    state = qags_step(S, sh.head[myslot], *myL, r_, e_, rc_, area2, error2, defab2);
                }
                if (state == QS_DONE) {
                    if (nint0 != 0) {
                        const unsigned long long row = S.inst / 3ULL;
                        const size_t idx = ((size_t)row * 7 + (size_t)c) * 3 + (size_t)(S.inst % 3ULL);
                        cst_res[idx] = S.result;
                        cst_last[idx] = S.last;
                        cst_ifail[idx] = S.ifail;
                        S.nint = 0;
                    }
                    if (!exhausted) {
                        const unsigned long long inst = atomicAdd(counter, 1ULL);
                        if (inst >= total) exhausted = true;
                        else {
                            const int sub = (int)(inst % 3ULL);
                            const unsigned long long row = inst / 3ULL;
                            const int i = (int)(row % (unsigned long long)n) + 1;
                            const int ixi = (int)(row / (unsigned long long)n);
                            const double xi = xi_arr[ixi];
                            const double x = grid_node(i, n, xi, grid_type);
                            const double eps = 1e-15;                        // integration.f90:56-63
                            const double ymin = -1.0 + eps, ymax = 1.0 - eps;
                            double a, b;
                            if (sub == 0) { a = ymin; b = -fabs(x) - eps; }
                            else if (sub == 1) { a = -fabs(x) + eps; b = fabs(x) - eps; }
                            else { a = fabs(x) + eps; b = ymax; }
                            S.inst = inst; S.x = x; S.xi = xi; S.a = a; S.b = b;
                            S.xx1 = 0.5 * (1.0 + x / xi);
                            S.xx2 = 0.5 * (1.0 - x / xi);
                            S.a1 = a; S.b1 = b; S.a2 = a; S.b2 = b;
                            S.nint = 1;
                        }
                    }
                }
            }
        }
        if (tid == 0) { sh.nfull = 0; sh.nerbl = 0; }
        __syncthreads();
        const int any = __syncthreads_or(tid < QB && sh.slot[tid < QB ? tid : 0].nint != 0);
        if (!any) break;
        QB_ACC(0);
        QB_ACC(1);
#ifdef TKB_K2_STATS
        if (tid < QB && sh.slot[tid].nint != 0) atomicAdd(&g_k2stats[6], 1ULL);
        if (tid == 0) atomicAdd(&g_k2stats[5], 1ULL);
#endif

        // ---------------- N + C: node arguments, support classes, compaction ----------------
        // one thread per (instance, node): the four half-kernel tasks of a node are permutations of the
        // same four numbers (xx1, xx2 of the fixed point, yy1, yy2 of the node), so the "no skipped piece
        // can be non-finite" test is done once per node (conservatively: it must hold for all four)
        for (int e = tid; e < QB * 42; e += QT) {
            const int s = e / 42, node = e - 42 * s;
            const QSlot& S = sh.slot[s];
            if (node >= 21 * S.nint) continue;
            const int iv = node / 21, q = node - 21 * iv;
            const double a = iv ? S.a2 : S.a1, b = iv ? S.b2 : S.b1;
            const double centr = 5.0e-01 * (a + b);
            const double hlgth = 5.0e-01 * (b - a);
            const double y = node_eta(centr, hlgth, q);
            const double t = y / S.xi;
            const double yy1 = 0.5 * (1.0 + t), yy2 = 0.5 * (1.0 - t);
            sh.ny[s][node][0] = yy1;
            sh.ny[s][node][1] = yy2;
            const double xx1 = S.xx1, xx2 = S.xx2;
            const bool safe = node_is_safe(xx1, xx2, yy1, yy2);
            int cls[4];
            cls[0] = task_class(xx1, yy1, safe);       // dir 0 (K(x,y)), half 0
            cls[1] = task_class(xx2, yy2, safe);       // dir 0, half 1
            cls[2] = task_class(yy1, xx1, safe);       // dir 1 (K(y,x)), half 0
            cls[3] = task_class(yy2, xx2, safe);       // dir 1, half 1
            int nf = 0, ne = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!HAS_ERBL && cls[k] == 1) cls[k] = 0;
                if (cls[k] == 0) sh.hv[s][4 * node + k] = 0.0;
                nf += (cls[k] == 2);
                ne += (cls[k] == 1);
            }
            int of = nf ? atomicAdd(&sh.nfull, nf) : 0;
            int oe = ne ? atomicAdd(&sh.nerbl, ne) : 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned short id = (unsigned short)(s * 168 + 4 * node + k);
                if (cls[k] == 2) sh.live[of++] = id;
                if (cls[k] == 1) sh.live[QB * 168 - 1 - (oe++)] = id;
            }
        }
        __syncthreads();
        QB_ACC(2);

        // ---------------- E: evaluate the live half kernels ----------------
        const int nfull = sh.nfull, nerbl = sh.nerbl;
        for (int idx = tid; idx < nfull; idx += QT) {
            const int e = sh.live[idx];
            const int s = e / 168, task = e - 168 * s;
            const QSlot& S = sh.slot[s];
            const int node = task >> 2, dir = (task >> 1) & 1, half = task & 1;
            const double yy1 = sh.ny[s][node][0], yy2 = sh.ny[s][node][1];
            const double U1 = dir ? yy1 : S.xx1, U2 = dir ? yy2 : S.xx2;
            const double V1 = dir ? S.xx1 : yy1, V2 = dir ? S.xx2 : yy2;
            double out[NKERN];
            eval_half<bit(KID), true>(half ? U2 : U1, half ? V2 : V1, half ? U1 : U2, half ? V1 : V2, out);
            sh.hv[s][task] = out[KID];
        }
        if (HAS_ERBL) {
            for (int idx = tid; idx < nerbl; idx += QT) {
                const int e = sh.live[QB * 168 - 1 - idx];
                const int s = e / 168, task = e - 168 * s;
                const QSlot& S = sh.slot[s];
                const int node = task >> 2, dir = (task >> 1) & 1, half = task & 1;
                const double yy1 = sh.ny[s][node][0], yy2 = sh.ny[s][node][1];
                const double U1 = dir ? yy1 : S.xx1, U2 = dir ? yy2 : S.xx2;
                const double V1 = dir ? S.xx1 : yy1, V2 = dir ? S.xx2 : yy2;
                double out[NKERN];
                eval_half<bit(KID), false>(half ? U2 : U1, half ? V2 : V1, half ? U1 : U2, half ? V1 : V2, out);
                sh.hv[s][task] = out[KID];
            }
        }
        __syncthreads();
        QB_ACC(3);

        // ---------------- K: K(x,y,xi) - K(y,x,xi) per node ----------------
        for (int e = tid; e < QB * 42; e += QT) {
            const int s = e / 42, node = e - 42 * s;
            const QSlot& S = sh.slot[s];
            if (node < 21 * S.nint) {
                const double* h = sh.hv[s] + 4 * node;
                const double Kxy = combine_halves<KID>(h[0], h[1], S.xi);
                const double Kyx = combine_halves<KID>(h[2], h[3], S.xi);
                sh.fv[s][node] = Kxy - Kyx;
            }
        }
        __syncthreads();
        QB_ACC(4);
    }
}

}  // namespace tkb
