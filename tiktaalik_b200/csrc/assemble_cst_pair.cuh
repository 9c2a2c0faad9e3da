// assemble_cst_pair.cuh -- K2 for TWO constants whose integrands share their transcendental values and most of their
// QUADPACK path: KV1_NSp and KV1_NSm (kernels_nlo.f90:60-78 and 95-111).  Included by assemble.cu after
// assemble_cst_pipe.cuh (same role split, same barriers, same QSlot / qags_* helpers).
//
// The two half kernels differ only in the sign of one term (tkb_kernels.cuh: base3 -+ QQhbar), so ONE evaluation gives
// both values, and their QAGS runs on the same sub-interval bisect the same intervals for ~84 % of their steps (the
// errors are dominated by the common pieces).  Here a block's slot holds both runs of one (xi, i, sub): two QUADPACK
// states (each with its own lists, its own comparisons, its own `last` / `ifail`: the control flow of every run is
// untouched), advanced by two lane pairs of the bookkeeping warp.  Per round each state posts its request; when the two
// requests are the same interval the evaluators classify and evaluate its nodes ONCE (node set 0 serves both states),
// otherwise each request gets its own node set.  A state that finishes first idles until its sibling is done; then the
// slot takes the next (xi, i, sub) for both.  Every state sees bit-identical integrand values in either case, so
// results and paths equal those of two separate launches; the evaluation work of the pair drops by ~40 %.
#pragma once

namespace tkb {

template <int QP>
struct QPairShared {
    QSlot slot[2 * QP * 2];               // [group][pair][state]: s = (g * QP + p) * 2 + st
    QHead head[2 * QP * 2];
    double ny[2 * QP][2][42][2];          // [pair slot][node set][node]: 0.5 (1 +- y/xi)
    double hv[2 * QP][2][2][168];         // [pair slot][node set][state][42 (2 dir + half) + node]: half-kernel values (the K pass reads consecutive nodes)
    double fv[2 * QP * 2][42];            // [state][node]: integrand values handed to the bookkeepers
    double pre[2 * QP][2][PRE_N];         // per pair slot and half: one-argument transcendental values (eval_half, FIX)
    unsigned short live[2][2][QP * 168];  // per group and direction: full tasks from the front, erbl-only from the back
    int nfull[2][2], nerbl[2][2];
    int next[2];
    int act[2];
    unsigned char same[2 * QP];           // per pair slot: the two states asked for the same interval (set 0 serves both)
    unsigned char own0[2 * QP];           // state whose request node set 0 carries (0: state A, 1: state B alone)
};

template <int KA, int KB, int QP, int NT, int BPS = 1>
__global__ void __launch_bounds__(NT, BPS)
assemble_cst_pair_kernel(int ca, int cb, int n, int nxi, const double* __restrict__ xi_arr, int grid_type,
                         QLists* __restrict__ lists, unsigned long long* __restrict__ counter,
                         double* __restrict__ cst_res, int* __restrict__ cst_last, int* __restrict__ cst_ifail)
{
    static_assert(4 * QP <= 32, "one bookkeeping warp per group, one lane pair per state");
    constexpr int NBK = 64;                           // two bookkeeping warps (one per group)
    constexpr int NE = NT - NBK;
    constexpr int NRQ = NE + 32;
    constexpr unsigned MASK = bit(KA) | bit(KB);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QPairShared<QP>& sh = *reinterpret_cast<QPairShared<QP>*>(smem_raw);
    log_tab_init();
    const int tid = threadIdx.x;
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    if (tid < 4 * QP) { sh.slot[tid].nint = 0; sh.slot[tid].term = 0; sh.slot[tid].defer = 0; }
    if (tid < 2) { sh.nfull[tid][0] = sh.nfull[tid][1] = 0; sh.nerbl[tid][0] = sh.nerbl[tid][1] = 0; sh.next[tid] = 0; }
    __syncthreads();

    if (tid < NBK) {
        // ======================= bookkeeper of group g: lane pair lp owns state lp (pair lp / 2, state lp % 2) =======================
        const int lane = tid & 31, part = lane & 1, lp = lane >> 1;
        const int g = tid >> 5;
        const bool mine = lp < 2 * QP;
        const int st = lp & 1;
        const int myslot = g * (2 * QP) + (mine ? lp : 0);
        QSlot& S = sh.slot[myslot];
        QLists* myL = lists + (size_t)blockIdx.x * (4 * QP) + myslot;
        const int c = st ? cb : ca;
        bool exhausted = false;
        for (int round = 0;; round++) {
#ifdef TKB_K2_STATS
            const long long tb0_ = qp_clk();
#endif
            if (round >= 1) qp_bar_sync(QP_BAR_VAL0 + g, NRQ);
#ifdef TKB_K2_STATS
            const long long tb1_ = qp_clk();
#endif
            const int nint0 = mine ? S.nint : 0;
            const bool term = mine && nint0 == 2 && S.term != 0;
            double r_ = 0, e_ = 0, ra_ = 0, rc_ = 0;
            if (nint0 == 2 && !term) qk21_reduce(sh.fv[myslot] + 21 * part, part ? S.a2 : S.a1, part ? S.b2 : S.b1, r_, e_, ra_, rc_);
            const double area2 = __shfl_down_sync(0xffffffffu, r_, 1);
            const double error2 = __shfl_down_sync(0xffffffffu, e_, 1);
            const double defab2 = __shfl_down_sync(0xffffffffu, rc_, 1);
            int idle = 1;
            if (mine && part == 0) {
                int state = QS_DONE;
                if (term) {
                    S.term = 0;
                } else if (nint0 == 1) {
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
                    state = qags_step<true>(S, sh.head[myslot], *myL, r_, e_, rc_, area2, error2, defab2);
                }
                if (state == QS_DONE && nint0 != 0) {
                    const unsigned long long row = S.inst / 3ULL;
                    const size_t idx = ((size_t)row * 7 + (size_t)c) * 3 + (size_t)(S.inst % 3ULL);
                    cst_res[idx] = S.result;
                    cst_last[idx] = S.last;
                    cst_ifail[idx] = S.ifail;
                    S.nint = 0;
                }
                idle = (S.nint == 0) ? 1 : 0;
            }
            // the slot takes its next (xi, i, sub) when BOTH of its runs are done; the lane of state A starts both
            const int sib_idle = __shfl_xor_sync(0xffffffffu, idle, 2);
            if (mine && part == 0 && st == 0 && idle && sib_idle && !exhausted) {
                const unsigned long long inst = atomicAdd(counter, 1ULL);
                if (inst >= total) exhausted = true;
                else {
                    const int sub = (int)(inst % 3ULL);
                    const unsigned long long row = inst / 3ULL;
                    const int ixi = (int)(row / (unsigned long long)n);
                    const double xi = xi_arr[ixi];
                    const double* rowtab = d_pre_tab + (size_t)row * PRE_ROW;       // x_i and 0.5 (1 +- x_i / xi) of the row
                    const double x = rowtab[PRE_X];
                    const double eps = 1e-15;                        // integration.f90:56-63
                    const double ymin = -1.0 + eps, ymax = 1.0 - eps;
                    double a, b;
                    if (sub == 0) { a = ymin; b = -fabs(x) - eps; }
                    else if (sub == 1) { a = -fabs(x) + eps; b = fabs(x) - eps; }
                    else { a = fabs(x) + eps; b = ymax; }
                    const double u1 = rowtab[PRE_U1], u2 = rowtab[PRE_U2];
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        QSlot& T = sh.slot[myslot + k];
                        T.inst = inst; T.x = x; T.xi = xi; T.a = a; T.b = b;
                        T.xx1 = u1; T.xx2 = u2;
                        T.a1 = a; T.b1 = b; T.a2 = a; T.b2 = b;
                        T.defer = 0; T.term = 0;
                        T.nint = 1;
                    }
                }
            }
            __syncwarp();
            const unsigned act = __ballot_sync(0xffffffffu, mine && part == 0 && S.nint != 0);
            if (lane == 0) {
                sh.nfull[g][0] = sh.nfull[g][1] = 0;
                sh.nerbl[g][0] = sh.nerbl[g][1] = 0;
                sh.next[g] = 0;
                sh.act[g] = (act != 0u) ? 1 : 0;
            }
            __syncwarp();
#ifdef TKB_K2_STATS
            const long long tb2_ = qp_clk();
#endif
            qp_bar_arrive(QP_BAR_REQ0 + g, NRQ);
            if (mine && part == 0 && S.nint == 2 && S.defer != 0) {
                S.defer = 0;
                qags_deferred(S, sh.head[myslot], *myL);
            }
            __syncwarp();
#ifdef TKB_K2_STATS
            if (lane == 0) {
                atomicAdd(&g_k2stats[0], (unsigned long long)(tb1_ - tb0_));
                atomicAdd(&g_k2stats[1], (unsigned long long)(tb2_ - tb1_));
                atomicAdd(&g_k2stats[2], (unsigned long long)(qp_clk() - tb2_));
                atomicAdd(&g_k2stats[3], 1ULL);
                atomicAdd(&g_k2stats[13], (unsigned long long)__popc(act));          // active states in this round
            }
#endif
            if (act == 0u) break;
        }
        return;
    }

    // ======================= evaluators =======================
    const int etid = tid - NBK;
    int alive = 3;
    for (int g = 0; alive != 0; g ^= 1) {
        if (!((alive >> g) & 1)) continue;
#ifdef TKB_K2_STATS
        const long long te0_ = qp_clk();
#endif
        qp_bar_sync(QP_BAR_REQ0 + g, NRQ);
#ifdef TKB_K2_STATS
        const long long te1_ = qp_clk();
#endif
        if (*(volatile int*)&sh.act[g] == 0) { alive &= ~(1 << g); continue; }

        // ---- N + C: node arguments, support classes, compaction (as in the single-constant pipeline) ----
        for (int e0 = 0; e0 < QP * 84; e0 += NE) {
            const int etid = qp_tid() - NBK, elane = etid & 31;
            const int e = e0 + etid;
            const bool inr = e < QP * 84;
            const int pl = inr ? e / 84 : 0, rem = inr ? e - 84 * pl : 0;
            const int set = rem / 42, node = rem - 42 * set;
            const int ps = g * QP + pl;
            // which node sets the round needs: set 0 carries state A's request (state B's when A idles), set 1 state B's
            // request when it differs from A's
            const QSlot& A = sh.slot[2 * ps];
            const QSlot& B = sh.slot[2 * ps + 1];
            const int na = A.nint, nb = B.nint;
            const bool same = (na != 0) && (na == nb) && (A.a1 == B.a1) && (A.b1 == B.b1) && (A.a2 == B.a2) && (A.b2 == B.b2);
            const int own0 = (na == 0 && nb != 0) ? 1 : 0;
            const int nset = (set == 0) ? ((na != 0) ? na : nb) : ((na != 0 && nb != 0 && !same) ? nb : 0);
            if (inr && rem == 0) { sh.same[ps] = same ? 1 : 0; sh.own0[ps] = (unsigned char)own0; }     // for the K pass
#ifdef TKB_K2_STATS
            if (inr && rem == 0) {
                if (same) atomicAdd(&g_k2stats[14], 1ULL);                       // pair slots whose states share the node set
                if (na != 0 && nb != 0 && !same) atomicAdd(&g_k2stats[21], 1ULL);  // both active, different requests
                if ((na != 0) != (nb != 0)) atomicAdd(&g_k2stats[22], 1ULL);       // one state waits for its sibling
            }
#endif
            const QSlot& S = (set == 0 && own0 == 0) ? A : B;
            const bool act = inr && node < 21 * nset;
            int cls[4] = { 0, 0, 0, 0 };
            if (act) {
                const int iv = node / 21, q = node - 21 * iv;
                const double a = iv ? S.a2 : S.a1, b = iv ? S.b2 : S.b1;
                const double centr = 5.0e-01 * (a + b);
                const double hlgth = 5.0e-01 * (b - a);
                const double y = node_eta(centr, hlgth, q);
                const double t = y / S.xi;
                const double yy1 = 0.5 * (1.0 + t), yy2 = 0.5 * (1.0 - t);
                sh.ny[ps][set][node][0] = yy1;
                sh.ny[ps][set][node][1] = yy2;
                const double xx1 = S.xx1, xx2 = S.xx2;
                const bool safe = node_is_safe(xx1, xx2, yy1, yy2);
                cls[0] = task_class(xx1, yy1, safe);
                cls[1] = task_class(xx2, yy2, safe);
                cls[2] = task_class(yy1, xx1, safe);
                cls[3] = task_class(yy2, xx2, safe);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (cls[k] == 0) { sh.hv[ps][set][0][42 * k + node] = 0.0; sh.hv[ps][set][1][42 * k + node] = 0.0; }
                // one-argument transcendental values of a slot that starts in this round, from the per-row table
                if (node == 0 && set == 0 && S.nint == 1) {
                    const double* rowtab = d_pre_tab + (size_t)(S.inst / 3ULL) * PRE_ROW;
#pragma unroll
                    for (int v = 0; v < 2 * PRE_N; v++)
                        if (((qp_pre_need<KA>() | qp_pre_need<KB>()) >> (v % PRE_N)) & 1u) sh.pre[ps][v / PRE_N][v % PRE_N] = rowtab[v];
                }
            }
            const unsigned lt = (1u << elane) - 1u;
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const unsigned f0 = __ballot_sync(0xffffffffu, cls[2 * d] == 2), f1 = __ballot_sync(0xffffffffu, cls[2 * d + 1] == 2);
                const unsigned r0 = __ballot_sync(0xffffffffu, cls[2 * d] == 1), r1 = __ballot_sync(0xffffffffu, cls[2 * d + 1] == 1);
                const int nf = __popc(f0) + __popc(f1), ne = __popc(r0) + __popc(r1);
                int bf = 0, be = 0;
                if (elane == 0) {
                    if (nf) bf = atomicAdd(&sh.nfull[g][d], nf);
                    if (ne) be = atomicAdd(&sh.nerbl[g][d], ne);
                }
                bf = __shfl_sync(0xffffffffu, bf, 0);
                be = __shfl_sync(0xffffffffu, be, 0);
                const unsigned short id0 = (unsigned short)((pl * 2 + set) * 168 + 4 * node + 2 * d);
                if (cls[2 * d] == 2) sh.live[g][d][bf + __popc(f0 & lt)] = id0;
                if (cls[2 * d + 1] == 2) sh.live[g][d][bf + __popc(f0) + __popc(f1 & lt)] = id0 + 1;
                if (cls[2 * d] == 1) sh.live[g][d][QP * 168 - 1 - (be + __popc(r0 & lt))] = id0;
                if (cls[2 * d + 1] == 1) sh.live[g][d][QP * 168 - 1 - (be + __popc(r0) + __popc(r1 & lt))] = id0 + 1;
            }
        }
#ifdef TKB_K2_STATS
        const long long tn_ = qp_clk();
#endif
        qp_bar_sync(QP_BAR_EV, NE);
#ifdef TKB_K2_STATS
        const long long tp1_ = qp_clk();
#endif

        // ---- E: the live half kernels, both constants from one evaluation ----
        {
            const int elane = qp_tid() & 31;
#ifdef TKB_K2_HOIST_SEG
                // the segment sizes are re-read per chunk (four shared-memory loads) rather than kept live across the evaluation
                const int n0 = *(volatile int*)&sh.nfull[g][1], n1 = *(volatile int*)&sh.nfull[g][0];
                const int n2 = *(volatile int*)&sh.nerbl[g][1], n3 = *(volatile int*)&sh.nerbl[g][0];
                const int p0 = (n0 + 31) & ~31, p1 = p0 + ((n1 + 31) & ~31), p2 = p1 + ((n2 + 31) & ~31), p3 = p2 + ((n3 + 31) & ~31);
#endif
            for (;;) {
                int chunk = 0;
                if (elane == 0) chunk = atomicAdd(&sh.next[g], 1);
                chunk = __shfl_sync(0xffffffffu, chunk, 0);
#ifndef TKB_K2_HOIST_SEG
                // the segment sizes are re-read per chunk (four shared-memory loads) rather than kept live across the evaluation
                const int n0 = *(volatile int*)&sh.nfull[g][1], n1 = *(volatile int*)&sh.nfull[g][0];
                const int n2 = *(volatile int*)&sh.nerbl[g][1], n3 = *(volatile int*)&sh.nerbl[g][0];
                const int p0 = (n0 + 31) & ~31, p1 = p0 + ((n1 + 31) & ~31), p2 = p1 + ((n2 + 31) & ~31), p3 = p2 + ((n3 + 31) & ~31);
#endif
                const int idx = chunk * 32 + elane;
                if (chunk * 32 >= p3) break;
                const int seg = (idx < p0) ? 0 : ((idx < p1) ? 1 : ((idx < p2) ? 2 : 3));
                const int j = idx - ((seg == 0) ? 0 : ((seg == 1) ? p0 : ((seg == 2) ? p1 : p2)));
                const int cnt = (seg == 0) ? n0 : ((seg == 1) ? n1 : ((seg == 2) ? n2 : n3));
                if (j >= cnt) continue;
                const int dir = (seg & 1) ^ 1;
                const int e = (seg < 2) ? sh.live[g][dir][j] : sh.live[g][dir][QP * 168 - 1 - j];
                const int ss = e / 168, task = e - 168 * ss;
                const int pl = ss >> 1, set = ss & 1, ps = g * QP + pl;
                const QSlot& S = sh.slot[2 * ps];                       // x, xi and 0.5 (1 +- x/xi) are the slot's, equal for both states
                const int node = task >> 2, half = task & 1;
                const double yy1 = sh.ny[ps][set][node][0], yy2 = sh.ny[ps][set][node][1];
                const double U1 = dir ? yy1 : S.xx1, U2 = dir ? yy2 : S.xx2;
                const double V1 = dir ? S.xx1 : yy1, V2 = dir ? S.xx2 : yy2;
                const double X = half ? U2 : U1, Y = half ? V2 : V1, Xo = half ? U1 : U2, Yo = half ? V1 : V2;
                const double* preS = sh.pre[ps][half];
                const double* preO = sh.pre[ps][half ^ 1];
                double out[NKERN];
                if (seg < 2) eval_half<MASK, true, 3>(X, Y, Xo, Yo, out, preS, preO, dir ? 2 : 1);
                else eval_half<MASK, false, 3>(X, Y, Xo, Yo, out, preS, preO, dir ? 2 : 1);
                sh.hv[ps][set][0][42 * (task & 3) + node] = out[KA];
                sh.hv[ps][set][1][42 * (task & 3) + node] = out[KB];
            }
        }
#ifdef TKB_K2_STATS
        const long long tv_ = qp_clk();
#endif
        qp_bar_sync(QP_BAR_EV, NE);
#ifdef TKB_K2_STATS
        const long long tp2_ = qp_clk();
        const int ntask_ = sh.nfull[g][0] + sh.nfull[g][1] + sh.nerbl[g][0] + sh.nerbl[g][1];
#endif

        // ---- K: K(x,y,xi) - K(y,x,xi) per state and node ----
        for (int e = qp_tid() - NBK; e < QP * 84; e += NE) {
            const int sl = e / 42, node = e - 42 * sl;               // sl = state within the group
            const int pl = sl >> 1, st = sl & 1, ps = g * QP + pl;
            const QSlot& S = sh.slot[2 * ps + st];
            if (node < 21 * S.nint) {
                const int set = (st == 0) ? 0 : ((sh.same[ps] || sh.own0[ps]) ? 0 : 1);
                const double* h = sh.hv[ps][set][st] + node;
                const double Kxy = st ? combine_halves<KB>(h[0], h[42], S.xi) : combine_halves<KA>(h[0], h[42], S.xi);
                const double Kyx = st ? combine_halves<KB>(h[84], h[126], S.xi) : combine_halves<KA>(h[84], h[126], S.xi);
                sh.fv[2 * ps + st][node] = Kxy - Kyx;
            }
        }
#ifdef TKB_K2_STATS
        const long long tp3_ = qp_clk();
#endif
        qp_bar_arrive(QP_BAR_VAL0 + g, NRQ);
#ifdef TKB_K2_STATS
        const long long tp4_ = qp_clk();
        if ((qp_tid() & 31) == 0) {                                        // every evaluator warp reports
            atomicAdd(&g_k2stats[4], (unsigned long long)(te1_ - te0_));   // wait for the requests
            atomicAdd(&g_k2stats[16], (unsigned long long)(tn_ - te1_));   // node pass, own work
            atomicAdd(&g_k2stats[7], (unsigned long long)(tp1_ - tn_));    // barrier after the node pass
            atomicAdd(&g_k2stats[17], (unsigned long long)(tv_ - tp1_));   // evaluation, own chunks
            atomicAdd(&g_k2stats[15], (unsigned long long)(tp2_ - tv_));   // barrier after the evaluation
            atomicAdd(&g_k2stats[18], (unsigned long long)(tp4_ - tp2_));  // K pass + hand-over
            atomicAdd(&g_k2stats[6], 1ULL);
            if (etid == 0) atomicAdd(&g_k2stats[19], (unsigned long long)ntask_), atomicAdd(&g_k2stats[5], 1ULL);
        }
#endif
    }
}

}  // namespace tkb
