// assemble_cst_replay.cuh -- K2 with shared integrand evaluations between constants.
// Included by assemble.cu after assemble_cst_batched.cuh (whose round structure and helpers it reuses).
//
// The seven plus-prescription constants of one (xi, i, sub-interval) integrate different kernels over the
// same interval, and QUADPACK's bisection is driven by the same endpoint singularity, so their bisection
// paths start out identical: measured with the oracle, the common prefix with KV1_NSp covers 84 % of the
// integrand evaluations of KV1_NSm and 70 % of KV1_NS_nfl; with KV1_GG it covers 72 % (KA1_GG), 57 %
// (KV1_GG_nfl) and 61 % (KA1_GG_nfl).  Evaluating a second kernel at a node that is already being
// evaluated is nearly free (shared logarithms and dilogarithms, tkb_kernels.cuh).  So:
//   * a LEADER launch runs the QUADPACK instances of one constant exactly as before, but evaluates its
//     companions' kernels at every node as well and logs, per instance and step, the requested intervals and
//     the companions' 2 x 21 integrand values;
//   * a FOLLOWER launch runs the instances of a companion constant: while its own request equals the
//     leader's logged request of the same step it takes the integrand values from the log (bookkeeping
//     only, several steps per round); at the first difference it leaves the log and evaluates by itself.
// Every instance still executes its own QUADPACK control flow on its own kernel's values at its own nodes:
// results, `last` and `ifail` are bit-identical to the unshared evaluation.
#pragma once

namespace tkb {

constexpr int QR_MAXIT = 6;              // log-replay steps per round and instance

template <int QBv, int NKER>
struct QReplayShared {
    QSlot slot[QBv];
    QHead head[QBv];
    double ny[QBv][42][2];
    double hv[QBv][NKER][168];
    double fv[QBv][NKER][42];
    double pre[QBv][2][PRE_N];          // per instance and half: one-argument transcendental values (eval_half, FIX)
    unsigned short live[2][QBv * 84];   // per direction: full tasks from the front, erbl-only from the back
    int nfull[2], nerbl[2];
    int ev[QBv];                         // instance needs an evaluation in this round
    int have[QBv];                       // fv[slot][0] holds the values of the current request
    int insync[QBv];                     // follower: still on the leader's path
};

template <int KID, int C1, int C2, int C3>
struct KernelSet {
    static constexpr int NC = (C1 >= 0) + (C2 >= 0) + (C3 >= 0);
    static constexpr int NKER = 1 + NC;
    static constexpr unsigned MASK = (1u << KID) | (C1 >= 0 ? (1u << (C1 >= 0 ? C1 : 0)) : 0u) |
                                     (C2 >= 0 ? (1u << (C2 >= 0 ? C2 : 0)) : 0u) | (C3 >= 0 ? (1u << (C3 >= 0 ? C3 : 0)) : 0u);
    __device__ static constexpr int kid(int k) { return k == 0 ? KID : (k == 1 ? C1 : (k == 2 ? C2 : C3)); }
    // an ERBL-only piece exists for every QAGS kernel except KV1_qq_half_nfl
    static constexpr bool HAS_ERBL = (KID != K_NSN) || (C1 >= 0 && C1 != K_NSN) || (C2 >= 0 && C2 != K_NSN) ||
                                     (C3 >= 0 && C3 != K_NSN);
};

// c: constant index of KID (result slot).  Leader (NC > 0): log/nlog are written.  Follower (FOLLOW): log/nlog
// of its leader are read, `ci` is its companion index there and `rec` the leader's record length.
template <int KID, int C1, int C2, int C3, bool FOLLOW, int QBv, int QTv, int BPS>
__global__ void __launch_bounds__(QTv, BPS)
assemble_cst_replay_kernel(int c, int n, int nxi, const double* __restrict__ xi_arr, int grid_type,
                           QLists* __restrict__ lists, unsigned long long* __restrict__ counter,
                           double* __restrict__ cst_res, int* __restrict__ cst_last, int* __restrict__ cst_ifail,
                           double* __restrict__ log, int* __restrict__ nlog, int L, int rec, int ci)
{
    using KS = KernelSet<KID, C1, C2, C3>;
    constexpr int NKER = KS::NKER;
    constexpr bool LEADER = (KS::NC > 0);
    static_assert(!(LEADER && FOLLOW), "a launch is either a leader or a follower");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QReplayShared<QBv, NKER>& sh = *reinterpret_cast<QReplayShared<QBv, NKER>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long total = (unsigned long long)nxi * n * 3ULL;
    bool exhausted = false;
    if (tid < QBv) { sh.slot[tid].nint = 0; sh.ev[tid] = 0; sh.have[tid] = 0; sh.insync[tid] = 0; }
    __syncthreads();

    for (;;) {
        // ---------------- S: bookkeeping (and log replay for followers) ----------------
        {
            constexpr int NW = QTv / 32, SLW = (QBv + NW - 1) / NW;
            const int w = tid >> 5;
            const int cand = (lane >> 1) * NW + w;
            const bool mine = (lane < 2 * SLW) && (cand < QBv);
            const int myslot = mine ? cand : 0;
            const int part = lane & 1;
            QSlot& S = sh.slot[myslot];
            QLists* myL = lists + (size_t)blockIdx.x * QBv + myslot;
            for (int it = 0; it < (FOLLOW ? QR_MAXIT : 1); it++) {
                // (a) consume the values of the current request
                const int nint0 = (mine && sh.have[myslot]) ? S.nint : 0;
                double r_ = 0, e_ = 0, ra_ = 0, rc_ = 0;
#ifdef TKB_K2_STATS
                long long ts0_ = clock64();
#endif
                if (nint0 == 2) qk21_reduce(sh.fv[myslot][0] + 21 * part, part ? S.a2 : S.a1, part ? S.b2 : S.b1, r_, e_, ra_, rc_);
#ifdef TKB_K2_STATS
                if (nint0 == 2 && part == 0) { atomicAdd(&g_k2stats[12], (unsigned long long)(clock64() - ts0_)); ts0_ = clock64(); }
#endif
                const double area2 = __shfl_down_sync(0xffffffffu, r_, 1);
                const double error2 = __shfl_down_sync(0xffffffffu, e_, 1);
                const double defab2 = __shfl_down_sync(0xffffffffu, rc_, 1);
                if (mine && part == 0 && (nint0 != 0 || S.nint == 0)) {
                    int state = QS_DONE;
                    if (nint0 == 1) {
                        state = qags_first(S, sh.head[myslot], *myL, sh.fv[myslot][0]);
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
#ifdef TKB_K2_STATS
                        atomicAdd(&g_k2stats[13], (unsigned long long)(clock64() - ts0_)); atomicAdd(&g_k2stats[14], 1ULL);
#endif
                    }
                    sh.have[myslot] = 0;
                    if (state == QS_DONE) {
                        if (nint0 != 0) {
                            const unsigned long long row = S.inst / 3ULL;
                            const size_t idx = ((size_t)row * 7 + (size_t)c) * 3 + (size_t)(S.inst % 3ULL);
                            cst_res[idx] = S.result;
                            cst_last[idx] = S.last;
                            cst_ifail[idx] = S.ifail;
                            if (LEADER) nlog[S.inst] = (S.last < L) ? S.last : L;
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
                                sh.insync[myslot] = FOLLOW ? 1 : 0;
                            }
                        }
                    }
                }
                __syncwarp();
                // (b) follower: take the next values from the leader's log while the requests coincide
                bool again = false;
                if (FOLLOW) {
                    int hit = 0;
                    const double* recp = nullptr;
                    if (mine && S.nint != 0 && !sh.have[myslot] && sh.insync[myslot]) {
                        const int step = (S.nint == 1) ? 0 : S.last - 1;
                        if (step < L && step < nlog[S.inst]) {
                            recp = log + ((size_t)S.inst * L + step) * rec;
                            hit = (recp[0] == S.a1 && recp[1] == S.b1 && recp[2] == S.a2 && recp[3] == S.b2) ? 1 : 0;
                        }
                        if (!hit && part == 0) sh.insync[myslot] = 0;
                    }
                    __syncwarp();
                    if (hit) {
                        const double* src = recp + 4 + 42 * ci + 21 * part;
                        double* dst = sh.fv[myslot][0] + 21 * part;
                        if (S.nint == 2 || part == 0)
                            for (int q = 0; q < 21; q++) dst[q] = src[q];
                    }
                    __syncwarp();
                    if (hit && part == 0) sh.have[myslot] = 1;
                    again = __any_sync(0xffffffffu, hit != 0);
                    __syncwarp();
                }
                if (!again) break;
            }
            if (mine && part == 0) sh.ev[myslot] = (S.nint != 0 && !sh.have[myslot]) ? 1 : 0;
        }
        if (tid == 0) { sh.nfull[0] = sh.nfull[1] = 0; sh.nerbl[0] = sh.nerbl[1] = 0; }
        __syncthreads();
        const int any = __syncthreads_or(tid < QBv && sh.slot[tid < QBv ? tid : 0].nint != 0);
        if (!any) break;
        const int anyev = __syncthreads_or(tid < QBv && sh.ev[tid < QBv ? tid : 0] != 0);
        if (!anyev) continue;

        // ---------------- N + C: node arguments, support classes, compaction ----------------
        for (int e = tid; e < QBv * 42; e += QTv) {
            const int s = e / 42, node = e - 42 * s;
            const QSlot& S = sh.slot[s];
            if (!sh.ev[s] || node >= 21 * S.nint) continue;
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
            cls[0] = task_class(xx1, yy1, safe);
            cls[1] = task_class(xx2, yy2, safe);
            cls[2] = task_class(yy1, xx1, safe);
            cls[3] = task_class(yy2, xx2, safe);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!KS::HAS_ERBL && cls[k] == 1) cls[k] = 0;
                if (cls[k] == 0) {
#pragma unroll
                    for (int m = 0; m < NKER; m++) sh.hv[s][m][4 * node + k] = 0.0;
                }
            }
            // tasks are listed per direction so that a warp evaluates one direction (one set of pre-computed values)
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const int nf = (cls[2 * d] == 2) + (cls[2 * d + 1] == 2), ne = (cls[2 * d] == 1) + (cls[2 * d + 1] == 1);
                int of = nf ? atomicAdd(&sh.nfull[d], nf) : 0;
                int oe = ne ? atomicAdd(&sh.nerbl[d], ne) : 0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int k = 2 * d + h;
                    const unsigned short id = (unsigned short)(s * 168 + 4 * node + k);
                    if (cls[k] == 2) sh.live[d][of++] = id;
                    if (cls[k] == 1) sh.live[d][QBv * 84 - 1 - (oe++)] = id;
                }
            }
        }
        // one-argument transcendental values of the instances whose first request is evaluated in this round
        for (int e = tid; e < QBv * 2 * PRE_N; e += QTv) {
            const int s = e / (2 * PRE_N), v = e % (2 * PRE_N);
            const QSlot& S = sh.slot[s];
            if (sh.ev[s] && (S.nint == 1 || FOLLOW))
                sh.pre[s][v / PRE_N][v % PRE_N] = pre_value(v % PRE_N, (v / PRE_N) ? S.xx2 : S.xx1);
        }
        __syncthreads();

        // ---------------- E: evaluate the live half kernels (all kernels of the set at once).  Four segments
        // (direction x {full, erbl-only}), each padded to whole warps so that a warp runs one variant ----------------
        {
            const int n0 = sh.nfull[0], n1 = sh.nfull[1];
            const int n2 = KS::HAS_ERBL ? sh.nerbl[0] : 0, n3 = KS::HAS_ERBL ? sh.nerbl[1] : 0;
            const int p0 = (n0 + 31) & ~31, p1 = p0 + ((n1 + 31) & ~31), p2 = p1 + ((n2 + 31) & ~31), p3 = p2 + ((n3 + 31) & ~31);
            for (int idx = tid; idx < p3; idx += QTv) {
                const int seg = (idx < p0) ? 0 : ((idx < p1) ? 1 : ((idx < p2) ? 2 : 3));
                const int j = idx - ((seg == 0) ? 0 : ((seg == 1) ? p0 : ((seg == 2) ? p1 : p2)));
                const int cnt = (seg == 0) ? n0 : ((seg == 1) ? n1 : ((seg == 2) ? n2 : n3));
                if (j >= cnt) continue;
                const int dir = seg & 1;
                const int e = (seg < 2) ? sh.live[dir][j] : sh.live[dir][QBv * 84 - 1 - j];
                const int s = e / 168, task = e - 168 * s;
                const QSlot& S = sh.slot[s];
                const int node = task >> 2, half = task & 1;
                const double yy1 = sh.ny[s][node][0], yy2 = sh.ny[s][node][1];
                const double U1 = dir ? yy1 : S.xx1, U2 = dir ? yy2 : S.xx2;
                const double V1 = dir ? S.xx1 : yy1, V2 = dir ? S.xx2 : yy2;
                const double X = half ? U2 : U1, Y = half ? V2 : V1, Xo = half ? U1 : U2, Yo = half ? V1 : V2;
                const double* preS = sh.pre[s][half];
                const double* preO = sh.pre[s][half ^ 1];
                double out[NKERN];
                if (seg == 0) eval_half<KS::MASK, true, 1>(X, Y, Xo, Yo, out, preS, preO);
                else if (seg == 1) eval_half<KS::MASK, true, 2>(X, Y, Xo, Yo, out, preS, preO);
                else if (seg == 2) eval_half<KS::MASK, false, 1>(X, Y, Xo, Yo, out, preS, preO);
                else eval_half<KS::MASK, false, 2>(X, Y, Xo, Yo, out, preS, preO);
#pragma unroll
                for (int m = 0; m < NKER; m++) sh.hv[s][m][task] = out[KS::kid(m)];
            }
        }
        __syncthreads();

        // ---------------- K: K(x,y,xi) - K(y,x,xi) per node and kernel; leaders write the log ----------------
        for (int e = tid; e < QBv * 42; e += QTv) {
            const int s = e / 42, node = e - 42 * s;
            const QSlot& S = sh.slot[s];
            if (!sh.ev[s] || node >= 21 * S.nint) continue;
            double* recp = nullptr;
            if (LEADER) {
                const int step = (S.nint == 1) ? 0 : S.last - 1;
                if (step < L) {
                    recp = log + ((size_t)S.inst * L + step) * rec;
                    if (node == 0) { recp[0] = S.a1; recp[1] = S.b1; recp[2] = S.a2; recp[3] = S.b2; }
                }
            }
#pragma unroll
            for (int m = 0; m < NKER; m++) {
                const double* h = sh.hv[s][m] + 4 * node;
                const int kid = KS::kid(m);
                const double Kxy = combine_halves_rt(kid, h[0], h[1], S.xi);
                const double Kyx = combine_halves_rt(kid, h[2], h[3], S.xi);
                const double v = Kxy - Kyx;
                sh.fv[s][m][node] = v;
                if (LEADER && m > 0 && recp) recp[4 + 42 * (m - 1) + node] = v;
            }
            if (node == 0) sh.have[s] = 1;
        }
        __syncthreads();
    }
}

}  // namespace tkb
