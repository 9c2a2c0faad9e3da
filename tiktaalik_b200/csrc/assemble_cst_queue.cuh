// assemble_cst_queue.cuh -- K2, the QAGS constants, as a barrier-free work queue (round 2 experiment, TKB_K2_MODE=queue;
// the role-split pipeline of assemble_cst_pipe.cuh stays the default: it measured faster, see the end of this comment).
// Included by assemble.cu after qags_step.cuh (QSlot / QHead / qags_first / qags_step).
//
// Reference path: kernels_nlo.f90:60-78 ... 925-944 (the eight *_cst functions) -> integration.f90:51-64
// (adaptive_integrate) -> :231-904 (qags, qk21, qsort, qextr).
//
// The role-split pipeline (assemble_cst_pipe.cuh) still ran every group-round as three block-wide phases
// (node arguments + support classes, evaluation of the live half kernels, K(x,y) - K(y,x)) separated by barriers:
// ncu showed a third of all warp-cycles parked at a barrier (evaluators that ran out of 32-task chunks, the two
// latency-bound single passes in which every thread handles one node, the hand-overs).  Here nothing waits for a
// phase of ANOTHER group:
//   * the block owns NG groups of QG instances; warp g < NG is the BOOKKEEPER of group g (one lane pair per
//     instance, QUADPACK's scalar code, as in the pipeline);
//   * every other warp is a WORKER.  The work of a group-round is cut into 32-wide chunks of three kinds -- NODE
//     (node arguments, support classes, compaction), EVAL (live half kernels, four warp-uniform segments), K
//     (K(x,y) - K(y,x)) -- handed out through one ticket word per group: (epoch << 12 | next chunk), taken with
//     atomicAdd; a descriptor per epoch says what kind the chunks are and how many there are.  The worker that
//     finishes the last chunk of a phase publishes the next phase (NODE -> EVAL -> K) or, after K, wakes the
//     group's bookkeeper through a named barrier (bar.arrive / bar.sync, 64 threads); the bookkeeper publishes the
//     next NODE phase when its requests are posted.  A worker without a ticket in one group tries the next group,
//     so NODE / K chunks of one group run under the EVAL chunks of the others and no warp idles while any group
//     has work.
// The arithmetic and control flow of every instance are unchanged (same node arguments, same half-kernel values,
// same qk21 sums, same QUADPACK comparisons), so `result`, `last` and `ifail` stay bit-identical to the pipeline
// and to the oracle.
// Measured (C3, 256 xi, one B200; profiles/README.md round 2): K2 163 ms (2 groups of 16) ... 171 ms (3 x 12) ... 194 ms
// (4 x 10) against 145 ms for the pipeline in the same run.  The workers are 71-76 % busy in chunks, but (a) taking a
// ticket and signalling completion costs ~1.8 k cycles per 6-8 k-cycle chunk, and (b) the bookkeeping step, a serial
// chain of ~1 850 dependent instructions, slows from 24 k to 27-48 k cycles per round when every scheduler is kept
// busy, and every group waits for its bookkeeper once per round.  The SM is issue-limited during evaluation either
// way (IPC ~0.8), so keeping more warps busy lengthens every chunk instead of raising the throughput.
#pragma once

namespace tkb {

#ifndef TKB_QQ_SLEEP
#define TKB_QQ_SLEEP 100
#endif
enum { QQ_NODE = 0, QQ_EVAL = 1, QQ_K = 2 };

struct QDesc {
    unsigned et;            // (epoch << 12) | number of chunks: one word, so a reader never sees a torn pair
    int phase;
    int n[4];               // EVAL: tasks per segment (full dir 1, full dir 0, erbl dir 1, erbl dir 0)
};

template <int QG, int NG>
struct QQueueShared {
    QSlot slot[NG * QG];
    QHead head[NG * QG];
    double ny[NG * QG][42][2];
    double hv[NG * QG][168];
    double fv[NG * QG][42];
    double pre[NG * QG][2][PRE_N];      // per instance and half: one-argument transcendental values (eval_half, FIX)
    unsigned short live[NG][2][QG * 84];// per group and direction: full tasks from the front, erbl-only from the back
    unsigned ticket[NG];                // (epoch << 12) | next chunk of the group's current phase
    int done[NG];                       // chunks of the current phase that are finished
    QDesc desc[NG][2];                  // by epoch parity
    int nfull[NG][2], nerbl[NG][2];
    unsigned seq[NG];                   // order in which the groups' current rounds were posted (workers serve the oldest first)
    unsigned seqctr;
    int dead;                           // groups whose bookkeeper found no instance left
};

// one lane: open the next epoch of a group with `total` chunks of kind `phase`
template <int QG, int NG>
__device__ __forceinline__ void qq_publish(QQueueShared<QG, NG>& sh, int g, int phase, int total, int n0 = 0, int n1 = 0,
                                           int n2 = 0, int n3 = 0)
{
    const unsigned e = ((*(volatile unsigned*)&sh.ticket[g] >> 12) + 1u) & 0xfffffu;     // only a publisher changes the epoch
    QDesc& d = sh.desc[g][e & 1u];
    d.phase = phase;
    d.n[0] = n0; d.n[1] = n1; d.n[2] = n2; d.n[3] = n3;
    *(volatile int*)&sh.done[g] = 0;
    __threadfence_block();
    *(volatile unsigned*)&d.et = (e << 12) | (unsigned)total;
    __threadfence_block();
    atomicExch(&sh.ticket[g], e << 12);
}

template <int ID>
__device__ __forceinline__ void qq_bar_sync64()
{
    asm volatile("bar.sync %0, 64;" ::"n"(ID) : "memory");
}
template <int ID>
__device__ __forceinline__ void qq_bar_arrive64()
{
    __threadfence_block();
    asm volatile("bar.arrive %0, 64;" ::"n"(ID) : "memory");
}
__device__ __forceinline__ void qq_wait_values(int g)           // barrier ids 1 .. NG (immediates: see assemble_cst_pipe.cuh)
{
    switch (g) {
    case 0: qq_bar_sync64<1>(); break;
    case 1: qq_bar_sync64<2>(); break;
    case 2: qq_bar_sync64<3>(); break;
    case 3: qq_bar_sync64<4>(); break;
    case 4: qq_bar_sync64<5>(); break;
    default: qq_bar_sync64<6>(); break;
    }
}
__device__ __forceinline__ void qq_post_values(int g)
{
    switch (g) {
    case 0: qq_bar_arrive64<1>(); break;
    case 1: qq_bar_arrive64<2>(); break;
    case 2: qq_bar_arrive64<3>(); break;
    case 3: qq_bar_arrive64<4>(); break;
    case 4: qq_bar_arrive64<5>(); break;
    default: qq_bar_arrive64<6>(); break;
    }
}

template <int KID, int QG, int NG, int NT>
__global__ void __launch_bounds__(NT, 1)
assemble_cst_queue_kernel(int c, int n, int nxi, const double* __restrict__ xi_arr, int grid_type,
                          QLists* __restrict__ lists, unsigned long long* __restrict__ counter,
                          double* __restrict__ cst_res, int* __restrict__ cst_last, int* __restrict__ cst_ifail)
{
    static_assert(NG >= 1 && NG <= 6 && 2 * QG <= 32, "one bookkeeping warp per group, one lane pair per instance");
    static_assert(QG * 168 <= 65535 && (QG * 168 + 128) / 32 < 4096, "task ids are 16 bit, chunk counts 12 bit");
    constexpr int NI = NG * QG;
    constexpr int NCH = (QG * 42 + 31) / 32;          // NODE / K chunks of a group-round
    constexpr bool HAS_ERBL = (KID != K_NSN);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QQueueShared<QG, NG>& sh = *reinterpret_cast<QQueueShared<QG, NG>*>(smem_raw);
    log_tab_init();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    if (tid < NI) sh.slot[tid].nint = 0;
    if (tid < NG) {
        sh.ticket[tid] = 0u;
        sh.done[tid] = 0;
        sh.desc[tid][0].et = 0u;                      // epoch 0: no chunks
        sh.desc[tid][1].et = 0xfffff000u;             // an epoch no ticket carries yet
        sh.nfull[tid][0] = sh.nfull[tid][1] = 0;
        sh.nerbl[tid][0] = sh.nerbl[tid][1] = 0;
        sh.seq[tid] = 0u;
    }
    if (tid == 0) { sh.dead = 0; sh.seqctr = 0u; }
    __syncthreads();

    if (warp < NG) {
        // ======================= bookkeeper of group g: lane pair p owns instance p =======================
        const int g = warp, part = lane & 1;
        const bool mine = (lane >> 1) < QG;
        const int myslot = g * QG + (mine ? (lane >> 1) : 0);
        QSlot& S = sh.slot[myslot];
        QLists* myL = lists + (size_t)blockIdx.x * NI + myslot;
        bool exhausted = false;
        for (int round = 0;; round++) {
#ifdef TKB_K2_STATS
            const long long tb0_ = clock64();
#endif
            if (round >= 1) qq_wait_values(g);                  // the first round only pulls instances
#ifdef TKB_K2_STATS
            const long long tb1_ = clock64();
#endif
            const int nint0 = mine ? S.nint : 0;
            double r_ = 0, e_ = 0, ra_ = 0, rc_ = 0;
            if (nint0 == 2) qk21_reduce(sh.fv[myslot] + 21 * part, part ? S.a2 : S.a1, part ? S.b2 : S.b1, r_, e_, ra_, rc_);
            const double area2 = __shfl_down_sync(0xffffffffu, r_, 1);
            const double error2 = __shfl_down_sync(0xffffffffu, e_, 1);
            const double defab2 = __shfl_down_sync(0xffffffffu, rc_, 1);
#ifdef TKB_K2_STATS
            if (lane == 0) atomicAdd(&g_k2stats[12], (unsigned long long)(clock64() - tb1_));    // the two qk21 sums (warp-wide)
#endif
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
                            S.nint = 1;
                        }
                    }
                }
            }
            __syncwarp();
            const unsigned act = __ballot_sync(0xffffffffu, mine && part == 0 && S.nint != 0);
#ifdef TKB_K2_STATS
            if (lane == 0) {
                atomicAdd(&g_k2stats[0], (unsigned long long)(tb1_ - tb0_));          // bookkeeper waits for values
                atomicAdd(&g_k2stats[1], (unsigned long long)(clock64() - tb1_));     // bookkeeping
                atomicAdd(&g_k2stats[3], 1ULL);
            }
#endif
            if (act == 0u) {
                if (lane == 0) { __threadfence_block(); atomicAdd(&sh.dead, 1); }
                break;
            }
            if (lane == 0) {
                sh.nfull[g][0] = sh.nfull[g][1] = 0;
                sh.nerbl[g][0] = sh.nerbl[g][1] = 0;
                sh.seq[g] = atomicAdd(&sh.seqctr, 1u) & 0x0fffffffu;
                qq_publish(sh, g, QQ_NODE, NCH);
            }
            __syncwarp();
        }
        return;
    }

    // ======================= workers =======================
#ifdef TKB_K2_STATS
    long long st_idle_ = 0, st_node_ = 0, st_eval_ = 0, st_k_ = 0, st_nchunk_ = 0;
    const long long tk0_ = clock64();
#endif
    for (;;) {
        // lanes 0 .. NG-1 probe one group each; the oldest posted round with a free ticket is served first, so that a
        // group's phases finish (and its bookkeeper starts) as early as possible while the tails of its phases are
        // filled with the next group's chunks
        unsigned tk = 0u, et = 0u, key = 0xffffffffu;
        if (lane < NG) {
            tk = *(volatile unsigned*)&sh.ticket[lane];
            et = *(volatile unsigned*)&sh.desc[lane][(tk >> 12) & 1u].et;
            if ((et >> 12) == (tk >> 12) && (tk & 0xfffu) < (et & 0xfffu))
                key = (*(volatile unsigned*)&sh.seq[lane] << 3) | (unsigned)lane;
        }
        key = __reduce_min_sync(0xffffffffu, key);
        if (key == 0xffffffffu) {
            int dead = 0;
            if (lane == 0) dead = *(volatile int*)&sh.dead;
            dead = __shfl_sync(0xffffffffu, dead, 0);
            if (dead == NG) break;
#ifdef TKB_K2_STATS
            const long long ti_ = clock64();
#endif
            __nanosleep(TKB_QQ_SLEEP);
#ifdef TKB_K2_STATS
            st_idle_ += clock64() - ti_;
#endif
            continue;
        }
        const int g = (int)(key & 7u);
        int ok = 0;
        if (lane == 0) {
            tk = atomicAdd(&sh.ticket[g], 1u);
            __threadfence_block();
            et = *(volatile unsigned*)&sh.desc[g][(tk >> 12) & 1u].et;
            ok = ((et >> 12) == (tk >> 12) && (tk & 0xfffu) < (et & 0xfffu)) ? 1 : 0;
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) continue;                                   // another warp took the last ticket in between
        tk = __shfl_sync(0xffffffffu, tk, 0);
        et = __shfl_sync(0xffffffffu, et, 0);
        const int chunk = (int)(tk & 0xfffu), nchunks = (int)(et & 0xfffu);
        const QDesc& D = sh.desc[g][(tk >> 12) & 1u];
        const int phase = D.phase;
#ifdef TKB_K2_STATS
        const long long tc0_ = clock64();
#endif
        if (phase == QQ_NODE) {
            // ---- node arguments, support classes, compaction (per direction, so that a warp of the EVAL phase
            // evaluates one direction; warp-aggregated: one shared-memory atomic per warp and list) ----
            const int e = 32 * chunk + lane;
            const bool inr = e < QG * 42;
            const int s = g * QG + (inr ? e / 42 : 0), node = inr ? e % 42 : 0;
            const QSlot& S = sh.slot[s];
            const int nint = S.nint;
            const bool act = inr && node < 21 * nint;
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
                    if (cls[k] == 0) sh.hv[s][4 * node + k] = 0.0;
                }
                // one-argument transcendental values of an instance that starts in this round, from the per-row table
                if (qp_pre_need<KID>() != 0u && node == 0 && nint == 1) {
                    const double* rowtab = d_pre_tab + (size_t)(S.inst / 3ULL) * PRE_ROW;
#pragma unroll
                    for (int v = 0; v < 2 * PRE_N; v++)
                        if ((qp_pre_need<KID>() >> (v % PRE_N)) & 1u) sh.pre[s][v / PRE_N][v % PRE_N] = rowtab[v];
                }
            }
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const unsigned f0 = __ballot_sync(0xffffffffu, cls[2 * d] == 2), f1 = __ballot_sync(0xffffffffu, cls[2 * d + 1] == 2);
                const unsigned r0 = HAS_ERBL ? __ballot_sync(0xffffffffu, cls[2 * d] == 1) : 0u;
                const unsigned r1 = HAS_ERBL ? __ballot_sync(0xffffffffu, cls[2 * d + 1] == 1) : 0u;
                const int nf = __popc(f0) + __popc(f1), ne = __popc(r0) + __popc(r1);
                int bf = 0, be = 0;
                if (lane == 0) {
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
        } else if (phase == QQ_EVAL) {
            // ---- live half kernels: four segments (direction x {full, erbl-only}), each padded to whole warps so
            // that a warp runs one variant; most expensive segment first (K(y,x) needs more transcendental values
            // than K(x,y)), so that the cheap chunks fill the end of the phase ----
            const int n0 = D.n[0], n1 = D.n[1], n2 = D.n[2], n3 = D.n[3];
            const int p0 = (n0 + 31) & ~31, p1 = p0 + ((n1 + 31) & ~31), p2 = p1 + ((n2 + 31) & ~31);
            const int idx = chunk * 32 + lane;
            const int seg = (idx < p0) ? 0 : ((idx < p1) ? 1 : ((idx < p2) ? 2 : 3));
            const int j = idx - ((seg == 0) ? 0 : ((seg == 1) ? p0 : ((seg == 2) ? p1 : p2)));
            const int cnt = (seg == 0) ? n0 : ((seg == 1) ? n1 : ((seg == 2) ? n2 : n3));
            if (j < cnt) {
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
                sh.hv[s][task] = out[KID];
            }
        } else {
            // ---- K(x,y,xi) - K(y,x,xi) per node ----
            const int e = 32 * chunk + lane;
            if (e < QG * 42) {
                const int s = g * QG + e / 42, node = e % 42;
                const QSlot& S = sh.slot[s];
                if (node < 21 * S.nint) {
                    const double* h = sh.hv[s] + 4 * node;
                    const double Kxy = combine_halves<KID>(h[0], h[1], S.xi);
                    const double Kyx = combine_halves<KID>(h[2], h[3], S.xi);
                    sh.fv[s][node] = Kxy - Kyx;
                }
            }
        }
#ifdef TKB_K2_STATS
        {
            const long long dt_ = clock64() - tc0_;
            if (phase == QQ_NODE) st_node_ += dt_; else if (phase == QQ_EVAL) st_eval_ += dt_; else st_k_ += dt_;
            st_nchunk_++;
        }
#endif
        // ---- chunk done; the warp that finishes the phase opens the next one ----
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = (atomicAdd(&sh.done[g], 1) == nchunks - 1) ? 1 : 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (phase == QQ_K) {
                qq_post_values(g);                                   // whole warp: bar.arrive on the bookkeeper's barrier
            } else if (lane == 0) {
                __threadfence_block();
                if (phase == QQ_NODE) {
                    const int n0 = *(volatile int*)&sh.nfull[g][1], n1 = *(volatile int*)&sh.nfull[g][0];
                    const int n2 = HAS_ERBL ? *(volatile int*)&sh.nerbl[g][1] : 0, n3 = HAS_ERBL ? *(volatile int*)&sh.nerbl[g][0] : 0;
                    const int p3 = ((n0 + 31) & ~31) + ((n1 + 31) & ~31) + ((n2 + 31) & ~31) + ((n3 + 31) & ~31);
                    if (p3 > 0) qq_publish(sh, g, QQ_EVAL, p3 / 32, n0, n1, n2, n3);
                    else qq_publish(sh, g, QQ_K, NCH);
                } else {
                    qq_publish(sh, g, QQ_K, NCH);
                }
            }
            __syncwarp();
        }
    }
#ifdef TKB_K2_STATS
    if (lane == 0) {
        atomicAdd(&g_k2stats[4], (unsigned long long)st_idle_);
        atomicAdd(&g_k2stats[16], (unsigned long long)st_node_);
        atomicAdd(&g_k2stats[17], (unsigned long long)st_eval_);
        atomicAdd(&g_k2stats[18], (unsigned long long)st_k_);
        atomicAdd(&g_k2stats[19], (unsigned long long)st_nchunk_);
        atomicAdd(&g_k2stats[7], (unsigned long long)(clock64() - tk0_));
        atomicAdd(&g_k2stats[6], 1ULL);
    }
#endif
}

}  // namespace tkb
