// assemble_cst_pipe.cuh -- K2 with the QUADPACK bookkeeping taken off the evaluators' critical path.
// Included by assemble.cu after qags_step.cuh (QSlot / QHead / qags_first / qags_step).
//
// In the lock-step form of round 1 (removed) every round ends in a bookkeeping phase in which one
// lane pair per instance runs QUADPACK's scalar code (qk21 sums, list update, qsort, epsilon table: about
// 15-25 k cycles of dependent FP64 latency, the maximum over the block's instances) while all other threads
// wait at the barrier: ncu showed 4.1 of 11.9 warp-cycles per issued instruction spent in barrier stalls.
// Here a block is split by role:
//   * warp 0 is the BOOKKEEPER.  It owns two groups of QG instances, one lane pair per instance;
//   * all other warps are EVALUATORS.  They alternate between the groups: node arguments, support classes,
//     live half kernels, K(x,y) - K(y,x) of group g, then the same for group g^1.
// While the evaluators work on group g^1 the bookkeeper consumes the values of group g and posts its next
// requests, so the evaluators never wait for scalar code unless one bookkeeping step outlasts a whole
// evaluation phase.  The hand-over uses four named barriers (bar.arrive by the producer, bar.sync by the
// consumer) and a fifth one among the evaluators.
// The arithmetic and the control flow of every instance are those of the lock-step form, hence results,
// `last` and `ifail` stay bit-identical to it and to the oracle.
#pragma once

namespace tkb {

template <int QG>
struct QPipeShared {
    QSlot slot[2 * QG];
    QHead head[2 * QG];
    double ny[2 * QG][42][2];
    double hv[2 * QG][168];             // [instance][42 (2 dir + half) + node]: half-kernel values (the K pass reads consecutive nodes)
    double fv[2 * QG][42];
    double pre[2 * QG][2][PRE_N];       // per instance and half: one-argument transcendental values (eval_half, FIX)
    unsigned short live[2][2][QG * 84]; // per group and direction: full tasks from the front, erbl-only from the back
    int nfull[2][2], nerbl[2][2];
    int next[2];                        // E phase: next 32-task chunk of the group
    long long t_exh;                    // TKB_K2_STATS: clock at which this block first found the queue empty
    int act[2][8];                      // per group and bookkeeping warp: it still has a running instance
};

// which entries of the per-instance table (PRE_*) the half kernel of one constant reads (the N_* flags of eval_half)
template <int KID>
__host__ __device__ constexpr unsigned qp_pre_need()
{
    return (KID == K_NSN) ? 0u
         : (KID == K_V1GGN || KID == K_A1GGN) ? ((1u << PRE_L) | (1u << PRE_LB))
         : (KID == K_NSP || KID == K_NSM) ? ((1u << PRE_L) | (1u << PRE_LB) | (1u << PRE_D) | (1u << PRE_DB))
         : 0x1fu;
}

enum { QP_BAR_REQ0 = 1, QP_BAR_REQ1 = 2, QP_BAR_VAL0 = 3, QP_BAR_VAL1 = 4, QP_BAR_EV = 5 };

// barrier ids are immediates so that ptxas reserves only the barriers in use (a register id reserves all 16)
template <int ID>
__device__ __forceinline__ void qp_bar_sync_i(int count)
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void qp_bar_arrive_i(int count)
{
    __threadfence_block();               // the producer's shared-memory writes are visible before it is counted as arrived
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
__device__ __forceinline__ void qp_bar_sync(int id, int count)
{
    switch (id) {
    case 1: qp_bar_sync_i<1>(count); break;
    case 2: qp_bar_sync_i<2>(count); break;
    case 3: qp_bar_sync_i<3>(count); break;
    case 4: qp_bar_sync_i<4>(count); break;
    default: qp_bar_sync_i<5>(count); break;
    }
}
// threadIdx.x re-read behind an opaque asm: index arithmetic derived from it is redone in every phase of every round
// (a few integer instructions) instead of being hoisted out of the round loop, where the evaluation's register pressure
// spills it to local memory -- with ~200 KB of the SM given to shared memory those reloads miss L1 and serialise
// (measured: 9.6 k of 25 k cycles per group-round in the K pass of the paired kernel)
__device__ __forceinline__ int qp_tid()
{
#ifdef TKB_K2_PLAIN_TID
    return (int)threadIdx.x;
#else
    int t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t;
#endif
}
// clock read that keeps its place among the barrier instructions (qp_clk() may be scheduled across an asm barrier)
__device__ __forceinline__ long long qp_clk()
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory");
    return t;
}
__device__ __forceinline__ void qp_bar_arrive(int id, int count)
{
    switch (id) {
    case 1: qp_bar_arrive_i<1>(count); break;
    case 2: qp_bar_arrive_i<2>(count); break;
    case 3: qp_bar_arrive_i<3>(count); break;
    default: qp_bar_arrive_i<4>(count); break;
    }
}

template <int KID, int QG, int NT, int BPS, int BWG>
__global__ void __launch_bounds__(NT, BPS)
assemble_cst_pipe_kernel(int c, int n, int nxi, const double* __restrict__ xi_arr, int grid_type,
                         QLists* __restrict__ lists, unsigned long long* __restrict__ counter,
                         double* __restrict__ cst_res, int* __restrict__ cst_last, int* __restrict__ cst_ifail)
{
        static_assert(BWG >= 1 && QG % BWG == 0 && 2 * (QG / BWG) <= 32, "BWG bookkeeping warps per group, one lane pair per instance");
    constexpr int NBK = 2 * BWG * 32;                 // bookkeeping threads (both groups)
    constexpr int NE = NT - NBK;                      // evaluator threads
    constexpr int NRQ = NE + 32 * BWG;                // participants of a hand-over barrier: a group's bookkeepers + evaluators
    constexpr bool HAS_ERBL = (KID != K_NSN);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QPipeShared<QG>& sh = *reinterpret_cast<QPipeShared<QG>*>(smem_raw);
    log_tab_init();
    const int tid = threadIdx.x;
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    if (tid < 2 * QG) sh.slot[tid].nint = 0;
    if (tid == 0) sh.t_exh = 0;
    if (tid < 2) { sh.nfull[tid][0] = sh.nfull[tid][1] = 0; sh.nerbl[tid][0] = sh.nerbl[tid][1] = 0; sh.next[tid] = 0; }
    __syncthreads();

    if (tid < NBK) {
        // ======================= bookkeepers: warp b of group g owns the instances b, b + BWG, ... =======================
        const int lane = tid & 31, part = lane & 1;
        const int g = (tid >> 5) / BWG, bw = (tid >> 5) % BWG;
        const bool mine = (lane >> 1) < QG / BWG;
        bool exhausted = false;
        for (int round = 0;; round++) {
#ifdef TKB_K2_STATS
            long long tb0_ = qp_clk();
#endif
            if (round >= 1) qp_bar_sync(QP_BAR_VAL0 + g, NRQ);      // the first round only pulls instances
#ifdef TKB_K2_STATS
            long long tb1_ = qp_clk();
#endif
            const int myslot = g * QG + (mine ? (bw + BWG * (lane >> 1)) : 0);
            QSlot& S = sh.slot[myslot];
            QLists* myL = lists + (size_t)blockIdx.x * (2 * QG) + myslot;
            const int nint0 = mine ? S.nint : 0;
            const bool term = mine && nint0 == 2 && S.term != 0;      // ended by the deferred extrapolation: these values are dropped
            double r_ = 0, e_ = 0, ra_ = 0, rc_ = 0;
            if (nint0 == 2 && !term) qk21_reduce(sh.fv[myslot] + 21 * part, part ? S.a2 : S.a1, part ? S.b2 : S.b1, r_, e_, ra_, rc_);
            const double area2 = __shfl_down_sync(0xffffffffu, r_, 1);
            const double error2 = __shfl_down_sync(0xffffffffu, e_, 1);
            const double defab2 = __shfl_down_sync(0xffffffffu, rc_, 1);
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
#ifdef TKB_K2_STATS
                        if (inst >= total && sh.t_exh == 0) sh.t_exh = qp_clk();
#endif
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
                            S.inst = inst; S.x = x; S.xi = xi; S.a = a; S.b = b;
                            S.xx1 = rowtab[PRE_U1];
                            S.xx2 = rowtab[PRE_U2];
                            S.a1 = a; S.b1 = b; S.a2 = a; S.b2 = b;
                            S.defer = 0; S.term = 0;
                            S.nint = 1;
                        }
                    }
                }
            }
            __syncwarp();
            const unsigned act = __ballot_sync(0xffffffffu, mine && part == 0 && S.nint != 0);
            if (lane == 0) {
                if (bw == 0) {
                    sh.nfull[g][0] = sh.nfull[g][1] = 0;
                    sh.nerbl[g][0] = sh.nerbl[g][1] = 0;
                    sh.next[g] = 0;
                }
                sh.act[g][bw] = (act != 0u) ? 1 : 0;
            }
            __syncwarp();
            // hand-over: the bookkeepers only ARRIVE at the requests barrier (the evaluators wait on it), then run the
            // extrapolation steps they deferred while the evaluators work on the requests
#ifdef TKB_K2_STATS
            long long tb2_ = qp_clk();
#endif
            qp_bar_arrive(QP_BAR_REQ0 + g, NRQ);
            if (mine && part == 0 && S.nint == 2 && S.defer != 0) {
                S.defer = 0;
                qags_deferred(S, sh.head[myslot], *myL);
            }
            __syncwarp();
            if (BWG > 1) {                                           // the group's bookkeeping warps: act[] of the siblings is written
                if (g == 0) qp_bar_sync_i<6>(32 * BWG); else qp_bar_sync_i<7>(32 * BWG);
            }
#ifdef TKB_K2_STATS
            if (lane == 0 && bw == 0) {
                atomicAdd(&g_k2stats[0], (unsigned long long)(tb1_ - tb0_));          // bookkeeper waits for values
                atomicAdd(&g_k2stats[1], (unsigned long long)(tb2_ - tb1_));          // bookkeeping up to the hand-over
                atomicAdd(&g_k2stats[2], (unsigned long long)(qp_clk() - tb2_));     // deferred extrapolation after the hand-over
                atomicAdd(&g_k2stats[3], 1ULL);
            }
#endif
            int any = 0;
#pragma unroll
            for (int b = 0; b < BWG; b++) any |= *(volatile int*)&sh.act[g][b];
            if (!any) break;
        }
        return;
    }

    // ======================= evaluators =======================
    const int etid = tid - NBK;
#ifdef TKB_K2_STATS
    const long long tk0_ = qp_clk();
#endif
    int alive = 3;
    for (int g = 0; alive != 0; g ^= 1) {
        if (!((alive >> g) & 1)) continue;
#ifdef TKB_K2_STATS
        long long te0_ = qp_clk();
#endif
        qp_bar_sync(QP_BAR_REQ0 + g, NRQ);
#ifdef TKB_K2_STATS
        long long te1_ = qp_clk();
#endif
        {
            int any = 0;
#pragma unroll
            for (int b = 0; b < BWG; b++) any |= *(volatile int*)&sh.act[g][b];
            if (!any) { alive &= ~(1 << g); continue; }
        }
        // ---- N + C: node arguments, support classes, compaction; the tasks are
        // listed per direction so that a warp evaluates one direction (one set of pre-computed values) ----
        // (warp-aggregated compaction: one shared-memory atomic per warp and list instead of one per thread)
        for (int e0 = 0; e0 < QG * 42; e0 += NE) {
            const int etid = qp_tid() - NBK;
            const int e = e0 + etid;
            const bool inr = e < QG * 42;
            const int s = g * QG + (inr ? e / 42 : 0), node = inr ? e % 42 : 0;
            const QSlot& S = sh.slot[s];
            const bool act = inr && node < 21 * S.nint;
            int cls[4] = { 0, 0, 0, 0 };
            if (act) {
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
                cls[0] = task_class(xx1, yy1, safe);       // dir 0 (K(x,y)), half 0
                cls[1] = task_class(xx2, yy2, safe);       // dir 0, half 1
                cls[2] = task_class(yy1, xx1, safe);       // dir 1 (K(y,x)), half 0
                cls[3] = task_class(yy2, xx2, safe);       // dir 1, half 1
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (!HAS_ERBL && cls[k] == 1) cls[k] = 0;
                    if (cls[k] == 0) sh.hv[s][42 * k + node] = 0.0;
                }
            }
            const int elane_ = etid & 31;
            const unsigned lt = (1u << elane_) - 1u;
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const unsigned f0 = __ballot_sync(0xffffffffu, cls[2 * d] == 2), f1 = __ballot_sync(0xffffffffu, cls[2 * d + 1] == 2);
                const unsigned r0 = HAS_ERBL ? __ballot_sync(0xffffffffu, cls[2 * d] == 1) : 0u;
                const unsigned r1 = HAS_ERBL ? __ballot_sync(0xffffffffu, cls[2 * d + 1] == 1) : 0u;
                const int nf = __popc(f0) + __popc(f1), ne = __popc(r0) + __popc(r1);
                int bf = 0, be = 0;
                if (elane_ == 0) {
                    if (nf) bf = atomicAdd(&sh.nfull[g][d], nf);
                    if (ne) be = atomicAdd(&sh.nerbl[g][d], ne);
                }
                bf = __shfl_sync(0xffffffffu, bf, 0);
                be = __shfl_sync(0xffffffffu, be, 0);
                const unsigned short id0 = (unsigned short)((s - g * QG) * 168 + 4 * node + 2 * d);
                if (cls[2 * d] == 2) sh.live[g][d][bf + __popc(f0 & lt)] = id0;
                if (cls[2 * d + 1] == 2) sh.live[g][d][bf + __popc(f0) + __popc(f1 & lt)] = id0 + 1;
                if (HAS_ERBL) {
                    if (cls[2 * d] == 1) sh.live[g][d][QG * 84 - 1 - (be + __popc(r0 & lt))] = id0;
                    if (cls[2 * d + 1] == 1) sh.live[g][d][QG * 84 - 1 - (be + __popc(r0) + __popc(r1 & lt))] = id0 + 1;
                }
            }
        }
        // one-argument transcendental values of the instances that start in this round (nint == 1)
        if (qp_pre_need<KID>() != 0u) {
            const int etid = qp_tid() - NBK;
            for (int e = etid; e < QG * 2 * PRE_N; e += NE) {
                const int s = g * QG + e / (2 * PRE_N), v = e % (2 * PRE_N);
                const QSlot& S = sh.slot[s];
                if (S.nint == 1 && ((qp_pre_need<KID>() >> (v % PRE_N)) & 1u))
                    sh.pre[s][v / PRE_N][v % PRE_N] = d_pre_tab[(size_t)(S.inst / 3ULL) * PRE_ROW + v];   // per-row table
            }
        }
        qp_bar_sync(QP_BAR_EV, NE);
#ifdef TKB_K2_STATS
        const long long tp1_ = qp_clk();
#endif

        // ---- E: evaluate the live half kernels.  Four segments (direction x {full, erbl-only}), each padded
        // to whole warps so that a warp runs one variant; the warps draw 32-task chunks from a counter, most
        // expensive segment first (K(y,x) needs more transcendental values than K(x,y)), so that the cheap
        // chunks fill the end of the phase ----
        {
            const int elane = qp_tid() & 31;
#ifdef TKB_K2_HOIST_SEG
                // the segment sizes are re-read per chunk (four shared-memory loads) rather than kept live across the evaluation
                const int n0 = *(volatile int*)&sh.nfull[g][1], n1 = *(volatile int*)&sh.nfull[g][0];
                const int n2 = HAS_ERBL ? *(volatile int*)&sh.nerbl[g][1] : 0, n3 = HAS_ERBL ? *(volatile int*)&sh.nerbl[g][0] : 0;
                const int p0 = (n0 + 31) & ~31, p1 = p0 + ((n1 + 31) & ~31), p2 = p1 + ((n2 + 31) & ~31), p3 = p2 + ((n3 + 31) & ~31);
#endif
            for (;;) {
                int chunk = 0;
                if (elane == 0) chunk = atomicAdd(&sh.next[g], 1);
                chunk = __shfl_sync(0xffffffffu, chunk, 0);
#ifndef TKB_K2_HOIST_SEG
                // the segment sizes are re-read per chunk (four shared-memory loads) rather than kept live across the evaluation
                const int n0 = *(volatile int*)&sh.nfull[g][1], n1 = *(volatile int*)&sh.nfull[g][0];
                const int n2 = HAS_ERBL ? *(volatile int*)&sh.nerbl[g][1] : 0, n3 = HAS_ERBL ? *(volatile int*)&sh.nerbl[g][0] : 0;
                const int p0 = (n0 + 31) & ~31, p1 = p0 + ((n1 + 31) & ~31), p2 = p1 + ((n2 + 31) & ~31), p3 = p2 + ((n3 + 31) & ~31);
#endif
                const int idx = chunk * 32 + elane;
                if (chunk * 32 >= p3) break;
                const int seg = (idx < p0) ? 0 : ((idx < p1) ? 1 : ((idx < p2) ? 2 : 3));
                const int j = idx - ((seg == 0) ? 0 : ((seg == 1) ? p0 : ((seg == 2) ? p1 : p2)));
                const int cnt = (seg == 0) ? n0 : ((seg == 1) ? n1 : ((seg == 2) ? n2 : n3));
                if (j >= cnt) continue;
                const int dir = (seg & 1) ^ 1;
                const int e = (seg < 2) ? sh.live[g][dir][j] : sh.live[g][dir][QG * 84 - 1 - j];
                const int sl = e / 168, task = e - 168 * sl, s = g * QG + sl;
                const QSlot& S = sh.slot[s];
                const int node = task >> 2, half = task & 1;
                const double yy1 = sh.ny[s][node][0], yy2 = sh.ny[s][node][1];
                const double U1 = dir ? yy1 : S.xx1, U2 = dir ? yy2 : S.xx2;
                const double V1 = dir ? S.xx1 : yy1, V2 = dir ? S.xx2 : yy2;
                const double X = half ? U2 : U1, Y = half ? V2 : V1, Xo = half ? U1 : U2, Yo = half ? V1 : V2;
                const double* preS = sh.pre[s][half];
                const double* preO = sh.pre[s][half ^ 1];
                double out[NKERN];
                if (seg < 2) eval_half<bit(KID), true, 3>(X, Y, Xo, Yo, out, preS, preO, dir ? 2 : 1);
                else eval_half<bit(KID), false, 3>(X, Y, Xo, Yo, out, preS, preO, dir ? 2 : 1);
                sh.hv[s][42 * (task & 3) + node] = out[KID];
            }
        }
        qp_bar_sync(QP_BAR_EV, NE);
#ifdef TKB_K2_STATS
        const long long tp2_ = qp_clk();
        const int ntask_ = sh.nfull[g][0] + sh.nfull[g][1] + sh.nerbl[g][0] + sh.nerbl[g][1];
#endif

        // ---- K: K(x,y,xi) - K(y,x,xi) per node ----
        for (int e = qp_tid() - NBK; e < QG * 42; e += NE) {
            const int s = g * QG + e / 42, node = e % 42;
            const QSlot& S = sh.slot[s];
            if (node < 21 * S.nint) {
                const double* h = sh.hv[s] + node;
                const double Kxy = combine_halves<KID>(h[0], h[42], S.xi);
                const double Kyx = combine_halves<KID>(h[84], h[126], S.xi);
                sh.fv[s][node] = Kxy - Kyx;
            }
        }
        qp_bar_arrive(QP_BAR_VAL0 + g, NRQ);
#ifdef TKB_K2_STATS
        if (etid == 0) {
            atomicAdd(&g_k2stats[4], (unsigned long long)(te1_ - te0_));              // evaluators wait for requests
            atomicAdd(&g_k2stats[5], (unsigned long long)(qp_clk() - te1_));         // N+C, E, K
            atomicAdd(&g_k2stats[6], 1ULL);
            atomicAdd(&g_k2stats[16], (unsigned long long)(tp1_ - te1_));             // node pass
            atomicAdd(&g_k2stats[17], (unsigned long long)(tp2_ - tp1_));             // evaluation
            atomicAdd(&g_k2stats[18], (unsigned long long)(qp_clk() - tp2_));        // K(x,y) - K(y,x)
            atomicAdd(&g_k2stats[19], (unsigned long long)ntask_);
        }
#endif
    }
#ifdef TKB_K2_STATS
    if (etid == 0) {
        const long long now_ = qp_clk();
        atomicAdd(&g_k2stats[7], (unsigned long long)(now_ - tk0_));                              // block lifetime
        atomicAdd(&g_k2stats[15], (unsigned long long)(sh.t_exh ? now_ - sh.t_exh : 0));           // of which after the queue ran dry
    }
#endif
}

}  // namespace tkb
