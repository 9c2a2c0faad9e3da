/* tko_matevo.c -- oracle restatement of kernel-matrix assembly, alpha_s and RK4 stepping.
 *
 * TEST INFRASTRUCTURE ONLY (see tko_fp.h).
 * Follows
 *   f90src/evolution/convolution.f90:26-97   (pixel_conv = reg + pls + cst)
 *   f90src/evolution/matevo.f90:64-96        (make_kernels, make_matrices)
 *   f90src/evolution/matevo.f90:124-278      (kernel / matrix getters)
 *   f90src/evolution/matevo.f90:283-480      (make_kernels_NS_0/NS_1/SG_0/SG_1)
 *   f90src/evolution/matevo.f90:485-696      (make_evomat_*, rk4_NS/SG, initialize_2D)
 *   f90src/qcd/alpha_qcd.f90:22-127          (alpha_s running, get_neff)
 * mode 0 keeps the reference's unfactored loop nest (one pixel_conv per (i,j), both
 * integrals run even for zero_func slots, OpenMP over ix like matevo.f90:288).
 * mode 1 evaluates each kernel once per (xi,i,node) and replays qk21's accumulation
 * per column j: identical per-term values and identical summation order, so it is
 * bit-identical to mode 0 (tests/test_oracle_assembly.py checks this) but ~nx times faster.
 * All arrays are Fortran-ordered (first index fastest), as the reference's are.
 */
#include <stdlib.h>
#include <string.h>
#include "tko.h"
#include "tko_fp.h"

#define PI 3.141592653589793
#define SQ(a) ((a) * (a))

/* ------------------------------------------------------------------ alpha_s */

static const double cMass2 = 1.29 * 1.29, bMass2 = 4.18 * 4.18, tMass2 = 172.44 * 172.44;
static const double alphaMZ = 0.118, MZ2 = 91.1876 * 91.1876;

/* alpha_qcd.f90:97-110 with norder = 1 */
static double get_beta_qcd(double alpha, int nfl)
{
    double fn = (double)(float)nfl;
    double beta0 = (11.0 * TKO_CA - 4.0 * fn * TKO_TF) / (12.0 * PI);
    double beta1 = (17.0 * SQ(TKO_CA) - fn * TKO_TF * (10.0 * TKO_CA + 6.0 * TKO_CF)) / (24.0 * SQ(PI));
    double beta = beta0;
    beta = beta + alpha * beta1;
    beta = -beta * SQ(alpha);
    return beta;
}
/* alpha_qcd.f90:52-72 */
static double evolve_alpha_fixed_nfl(double alpha0, double Q20, double Q2, int nfl, int nsteps)
{
    double h = tko_log(Q2 / Q20) / (double)(float)nsteps;
    double alpha = alpha0;
    for (int i = 1; i <= nsteps; i++) {
        double k1 = get_beta_qcd(alpha, nfl);
        double k2 = get_beta_qcd(alpha + h * k1 / 2.0, nfl);
        double k3 = get_beta_qcd(alpha + h * k2 / 2.0, nfl);
        double k4 = get_beta_qcd(alpha + h * k3, nfl);
        alpha = alpha + h / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
    }
    return alpha;
}
/* alpha_qcd.f90:74-92 */
static double alpha_mt2(void) { return evolve_alpha_fixed_nfl(alphaMZ, MZ2, tMass2, 5, 17); }
static double alpha_mb2(void) { return evolve_alpha_fixed_nfl(alphaMZ, MZ2, bMass2, 5, 17); }
static double alpha_mc2(void) { return evolve_alpha_fixed_nfl(alpha_mb2(), bMass2, cMass2, 4, 17); }
/* alpha_qcd.f90:22-50 */
double tko_get_alpha_qcd(double Q2)
{
    if (Q2 == MZ2) return alphaMZ;
    if (Q2 == tMass2) return alpha_mt2();
    if (Q2 == bMass2) return alpha_mb2();
    if (Q2 == cMass2) return alpha_mc2();
    if (Q2 > tMass2) return evolve_alpha_fixed_nfl(alpha_mt2(), tMass2, Q2, 6, 17);
    if (Q2 > bMass2) return evolve_alpha_fixed_nfl(alpha_mb2(), bMass2, Q2, 5, 17);
    if (Q2 > cMass2) return evolve_alpha_fixed_nfl(alpha_mc2(), cMass2, Q2, 4, 17);
    return evolve_alpha_fixed_nfl(alpha_mc2(), cMass2, Q2, 3, 17);
}
/* alpha_qcd.f90:115-127 */
int tko_get_neff(double Q2)
{
    int nfl = 3;
    if (Q2 >= tMass2) nfl = 6;
    else if (Q2 >= bMass2) nfl = 5;
    else if (Q2 >= cMass2) nfl = 4;
    return nfl;
}

/* -------------------------------------------------------------- convolution */

static double zero_k3(double x, double y, double xi) { (void)x; (void)y; (void)xi; return 0.0; }
static double zero_k2(double x, double xi) { (void)x; (void)xi; return 0.0; }

typedef struct { tko_k3 kernel; double x, xi; int n_pixels, j, grid_type; } conv_ctx;
/* convolution.f90:50-57 */
static double reg_integrand(double y, void *v)
{
    conv_ctx *c = (conv_ctx *)v;
    double fy = tko_interpixel(c->n_pixels, c->j, y, c->xi, c->grid_type);
    return c->kernel(c->x, y, c->xi) * fy;
}
/* convolution.f90:72-80 */
static double pls_integrand(double y, void *v)
{
    conv_ctx *c = (conv_ctx *)v;
    double fy = tko_interpixel(c->n_pixels, c->j, y, c->xi, c->grid_type);
    double fx = tko_interpixel(c->n_pixels, c->j, c->x, c->xi, c->grid_type);
    return c->kernel(c->x, y, c->xi) * (fy - fx);
}
/* convolution.f90:26-97 */
double tko_pixel_conv(tko_k3 Kreg, tko_k3 Kpls, tko_k2 Kcst, double xi, int N, int i, int j, int grid_type)
{
    if (!Kreg) Kreg = zero_k3;
    if (!Kpls) Kpls = zero_k3;
    if (!Kcst) Kcst = zero_k2;
    double x = tko_grid_node(i, N, xi, grid_type);
    double shift = 0.0;
    conv_ctx c = { Kreg, x, xi, N, j, grid_type };
    shift = shift + tko_integrate(reg_integrand, &c, x, xi, N, grid_type);
    c.kernel = Kpls;
    shift = shift + tko_integrate(pls_integrand, &c, x, xi, N, grid_type);
    double cst = 0.0;
    if (i == j) {
        double fx = tko_interpixel(N, j, x, xi, grid_type);
        cst = Kcst(x, xi) * fx;
    }
    shift = shift + cst;
    return shift;
}

/* ------------------------------------------------- the 24 pixel_conv "slots" */

typedef struct { tko_k3 reg, pls; tko_k2 cst; } slot_def;
enum {
    S_NS0, S_NS1P_0, S_NS1P_N, S_NS1M_0, S_NS1M_N,
    S_A0_QG, S_A0_GQ, S_A0_GG, S_A0_GGN,
    S_VMA0_QG, S_VMA0_GQ, S_VMA0_GG,
    S_V1_QQ, S_V1_QG, S_V1_GQ, S_V1_GQN, S_V1_GG, S_V1_GGN,
    S_A1_QQ, S_A1_QG, S_A1_GQ, S_A1_GQN, S_A1_GG, S_A1_GGN, N_SLOTS
};
/* matevo.f90:293, 306-311, 360-363, 382-384, 427-432, 453-458 */
static const slot_def slots[N_SLOTS] = {
    { 0, tko_K0_qq_pls, tko_K0_qq_cst },
    { 0, tko_KV1_NSp_pls, tko_KV1_NSp_cst },
    { 0, tko_KV1_NSp_pls_nfl, tko_KV1_NSp_cst_nfl },
    { 0, tko_KV1_NSm_pls, tko_KV1_NSm_cst },
    { 0, tko_KV1_NSm_pls_nfl, tko_KV1_NSm_cst_nfl },
    { tko_KA0_qG_reg, 0, 0 },
    { tko_KA0_Gq_reg, 0, 0 },
    { 0, tko_KA0_GG_pls, tko_KA0_GG_cst },
    { 0, 0, tko_KA0_GG_nfl },
    { tko_KVmA0_qG_reg, 0, 0 },
    { tko_KVmA0_Gq_reg, 0, 0 },
    { tko_KVmA0_GG_reg, 0, 0 },
    { tko_KV1_qq_reg, 0, 0 },
    { tko_KV1_qG_reg, 0, 0 },
    { tko_KV1_Gq_reg, 0, 0 },
    { tko_KV1_Gq_reg_nfl, 0, 0 },
    { 0, tko_KV1_GG_pls, tko_KV1_GG_cst },
    { 0, tko_KV1_GG_pls_nfl, tko_KV1_GG_cst_nfl },
    { tko_KA1_qq_reg, 0, 0 },
    { tko_KA1_qG_reg, 0, 0 },
    { tko_KA1_Gq_reg, 0, 0 },
    { tko_KA1_Gq_reg_nfl, 0, 0 },
    { 0, tko_KA1_GG_pls, tko_KA1_GG_cst },
    { 0, tko_KA1_GG_pls_nfl, tko_KA1_GG_cst_nfl },
};

/* ------------------------------------------------------------- module state */
/* matevo.f90:27-46 */
static int nx_cache = 0, nxi_cache = 0, nQ2_cache = 0, l_nlo_cache = 0;
static double *xi_cache = 0;
static double *K_NS_0[3], *KV_NS_1p[3], *KV_NS_1m[3];      /* (n,n,nxi) per nfl=3,4,5 */
static double *KV_SG_0[3], *KA_SG_0[3], *KV_SG_1[3], *KA_SG_1[3];   /* (2n,2n,nxi)    */
static double *M_NS_pls = 0, *M_NS_min = 0, *MV_SG = 0, *MA_SG = 0;

static void free_kernels(void)
{
    for (int f = 0; f < 3; f++) {
        free(K_NS_0[f]); free(KV_NS_1p[f]); free(KV_NS_1m[f]);
        free(KV_SG_0[f]); free(KA_SG_0[f]); free(KV_SG_1[f]); free(KA_SG_1[f]);
        K_NS_0[f] = KV_NS_1p[f] = KV_NS_1m[f] = 0;
        KV_SG_0[f] = KA_SG_0[f] = KV_SG_1[f] = KA_SG_1[f] = 0;
    }
}
static void free_matrices(void)
{
    free(M_NS_pls); free(M_NS_min); free(MV_SG); free(MA_SG);
    M_NS_pls = M_NS_min = MV_SG = MA_SG = 0;
    nQ2_cache = 0;
}
void tko_free_all(void)
{
    free_kernels(); free_matrices(); free(xi_cache); xi_cache = 0;
    nx_cache = nxi_cache = 0;
}
int tko_get_nx(void) { return nx_cache; }
int tko_get_nxi(void) { return nxi_cache; }
int tko_get_nQ2(void) { return nQ2_cache; }

/* ----------------------------------------------------- factored slot assembly */

/* One (xi, i): all 24 slot rows P[s][j], j=1..n, bit-identical to pixel_conv(i,j). */
static void assemble_row_factored(int n, double xi, int i, int grid_type, double *rows /* [N_SLOTS][n] */)
{
    static const double xgk[12] = { 0.0,
        9.956571630258081e-01, 9.739065285171717e-01, 9.301574913557082e-01, 8.650633666889845e-01,
        7.808177265864169e-01, 6.794095682990244e-01, 5.627571346686047e-01, 4.333953941292472e-01,
        2.943928627014602e-01, 1.488743389816312e-01, 0.000000000000000 };
    static const double wgk[12] = { 0.0,
        1.169463886737187e-02, 3.255816230796473e-02, 5.475589657435200e-02, 7.503967481091995e-02,
        9.312545458369761e-02, 1.093871588022976e-01, 1.234919762620659e-01, 1.347092173114733e-01,
        1.427759385770601e-01, 1.477391049013385e-01, 1.494455540029169e-01 };
    double x = tko_grid_node(i, n, xi, grid_type);
    double aa[8];
    aa[1] = tko_pull_back(-1.0, xi, n, grid_type);
    aa[2] = tko_pull_back(-fabs(x), xi, n, grid_type);
    aa[3] = tko_pull_back(-fabs(xi), xi, n, grid_type);
    aa[4] = tko_pull_back(0.0, xi, n, grid_type);
    aa[5] = tko_pull_back(fabs(xi), xi, n, grid_type);
    aa[6] = tko_pull_back(fabs(x), xi, n, grid_type);
    aa[7] = tko_pull_back(1.0, xi, n, grid_type);
    if (aa[3] < aa[2]) { double z = aa[2]; aa[2] = aa[3]; aa[3] = z; }
    if (aa[6] < aa[5]) { double z = aa[5]; aa[5] = aa[6]; aa[6] = z; }

    /* node table in qk21 evaluation order: per interval [centre, then (c-absc, c+absc) for
     * jtw = 2,4,..,10, then for jtwm1 = 1,3,..,9]; slot index 0 = centre, 1+2*(t-1) = minus, +1 = plus
     * where t runs over xgk index 1..10 */
    double yk[6][21], Jk[6][21], hl[6];
    for (int r = 0; r < 6; r++) {
        double a = aa[r + 1], b = aa[r + 2];
        double centr = 5.0e-01 * (a + b);
        double hlgth = 5.0e-01 * (b - a);
        hl[r] = hlgth;
        for (int t = 0; t <= 10; t++) {
            if (t == 0) {
                yk[r][0] = tko_push_forward(centr, xi, n, grid_type);
                Jk[r][0] = tko_push_jacob(centr, xi, n, grid_type);
            } else {
                double absc = hlgth * xgk[t];
                double e1 = centr - absc, e2 = centr + absc;
                yk[r][2 * t - 1] = tko_push_forward(e1, xi, n, grid_type);
                Jk[r][2 * t - 1] = tko_push_jacob(e1, xi, n, grid_type);
                yk[r][2 * t] = tko_push_forward(e2, xi, n, grid_type);
                Jk[r][2 * t] = tko_push_jacob(e2, xi, n, grid_type);
            }
        }
    }
    double *fy = (double *)malloc(sizeof(double) * 126 * n);   /* f_j(y_k) */
    double *fx = (double *)malloc(sizeof(double) * n);         /* f_j(x_i) */
    for (int r = 0; r < 6; r++)
        for (int k = 0; k < 21; k++)
            for (int j = 1; j <= n; j++)
                fy[((size_t)r * 21 + k) * n + (j - 1)] = tko_interpixel(n, j, yk[r][k], xi, grid_type);
    for (int j = 1; j <= n; j++) fx[j - 1] = tko_interpixel(n, j, x, xi, grid_type);

    double Kv[6][21];
    for (int s = 0; s < N_SLOTS; s++) {
        double *row = rows + (size_t)s * n;
        for (int j = 0; j < n; j++) row[j] = 0.0;
        for (int pass = 0; pass < 2; pass++) {
            tko_k3 fn = pass == 0 ? slots[s].reg : slots[s].pls;
            if (!fn) continue;                  /* zero_func slot: contributes exactly 0 */
            for (int r = 0; r < 6; r++)
                for (int k = 0; k < 21; k++) Kv[r][k] = fn(x, yk[r][k], xi);
            for (int j = 0; j < n; j++) {
                double integral = 0.0;
                for (int r = 0; r < 6; r++) {
#define FV(k) ((pass == 0 ? Kv[r][k] * fy[((size_t)r * 21 + (k)) * n + j]                       \
                          : Kv[r][k] * (fy[((size_t)r * 21 + (k)) * n + j] - fx[j])) * Jk[r][k])
                    double resk = wgk[11] * FV(0);
                    for (int jj = 1; jj <= 5; jj++) {
                        int t = 2 * jj;
                        double fsum = FV(2 * t - 1) + FV(2 * t);
                        resk = resk + wgk[t] * fsum;
                    }
                    for (int jj = 1; jj <= 5; jj++) {
                        int t = 2 * jj - 1;
                        double fsum = FV(2 * t - 1) + FV(2 * t);
                        resk = resk + wgk[t] * fsum;
                    }
#undef FV
                    integral = integral + resk * hl[r];
                }
                row[j] = row[j] + integral;
            }
        }
        if (slots[s].cst) row[i - 1] = row[i - 1] + slots[s].cst(x, xi) * fx[i - 1];
    }
    free(fy);
    free(fx);
}

/* ------------------------------------------------------------- make_kernels */

#define IDX3(i, j, k, n) ((size_t)(i) + (size_t)(n) * ((size_t)(j) + (size_t)(n) * (size_t)(k)))

int tko_make_kernels(int nx, int nxi, const double *xi_array, int grid_type, int mode)
{
    if (nx < 6 || nxi < 1) return 1;
    /* matevo.f90:64-80: deallocate_all, initialize_lagrange_weights, initialize_2D */
    free_kernels();
    free_matrices();
    tko_initialize_lagrange_weights(nx, 6);
    nx_cache = nx;
    nxi_cache = nxi;
    free(xi_cache);
    xi_cache = (double *)malloc(sizeof(double) * nxi);
    memcpy(xi_cache, xi_array, sizeof(double) * nxi);
    /* matevo.f90:688-689: only the first and last xi are clamped */
    if (xi_cache[0] < 1e-6) xi_cache[0] = 1e-6;
    if (xi_cache[nxi - 1] > 0.999999) xi_cache[nxi - 1] = 0.999999;

    const int n = nx;
    size_t nn = (size_t)n * n * nxi;
    double *P[N_SLOTS];
    for (int s = 0; s < N_SLOTS; s++) P[s] = (double *)malloc(sizeof(double) * nn);

    if (mode == 0) {
        /* reference-shaped: OMP over ix, then iy, iz, one pixel_conv per slot (matevo.f90:288-297) */
#pragma omp parallel for schedule(dynamic, 1)
        for (int ix = 1; ix <= n; ix++)
            for (int iy = 1; iy <= n; iy++)
                for (int iz = 1; iz <= nxi; iz++)
                    for (int s = 0; s < N_SLOTS; s++)
                        P[s][IDX3(ix - 1, iy - 1, iz - 1, n)] =
                            tko_pixel_conv(slots[s].reg, slots[s].pls, slots[s].cst,
                                           xi_cache[iz - 1], n, ix, iy, grid_type);
    } else {
#pragma omp parallel
        {
            double *rows = (double *)malloc(sizeof(double) * N_SLOTS * n);
#pragma omp for schedule(dynamic, 1) collapse(2)
            for (int iz = 1; iz <= nxi; iz++)
                for (int ix = 1; ix <= n; ix++) {
                    assemble_row_factored(n, xi_cache[iz - 1], ix, grid_type, rows);
                    for (int s = 0; s < N_SLOTS; s++)
                        for (int iy = 1; iy <= n; iy++)
                            P[s][IDX3(ix - 1, iy - 1, iz - 1, n)] = rows[(size_t)s * n + (iy - 1)];
                }
            free(rows);
        }
    }

    /* composition into nfl slices */
    size_t n2 = (size_t)2 * n;
    size_t nn2 = n2 * n2 * nxi;
    for (int f = 0; f < 3; f++) {
        double fn = (double)(float)(f + 3);
        K_NS_0[f] = (double *)malloc(sizeof(double) * nn);
        KV_NS_1p[f] = (double *)malloc(sizeof(double) * nn);
        KV_NS_1m[f] = (double *)malloc(sizeof(double) * nn);
        KA_SG_0[f] = (double *)calloc(nn2, sizeof(double));
        KV_SG_0[f] = (double *)calloc(nn2, sizeof(double));
        KA_SG_1[f] = (double *)calloc(nn2, sizeof(double));
        KV_SG_1[f] = (double *)calloc(nn2, sizeof(double));
        for (size_t e = 0; e < nn; e++) {
            K_NS_0[f][e] = P[S_NS0][e];                                       /* matevo.f90:293 */
            KV_NS_1p[f][e] = P[S_NS1P_0][e] + P[S_NS1P_N][e] * fn;            /* :331-334 */
            KV_NS_1m[f][e] = P[S_NS1M_0][e] + P[S_NS1M_N][e] * fn;
        }
        for (int iz = 0; iz < nxi; iz++)
            for (int iy = 0; iy < n; iy++)
                for (int ix = 0; ix < n; ix++) {
                    size_t e = IDX3(ix, iy, iz, n);
                    size_t qq = IDX3(ix, iy, iz, n2), qg = IDX3(ix, iy + n, iz, n2);
                    size_t gq = IDX3(ix + n, iy, iz, n2), gg = IDX3(ix + n, iy + n, iz, n2);
                    /* matevo.f90:370-376 */
                    KA_SG_0[f][qq] = K_NS_0[f][e];
                    KA_SG_0[f][qg] = fn * P[S_A0_QG][e];
                    KA_SG_0[f][gq] = P[S_A0_GQ][e];
                    KA_SG_0[f][gg] = P[S_A0_GG][e] + fn * P[S_A0_GGN][e];
                    /* matevo.f90:389-396 */
                    KV_SG_0[f][qq] = 0.0 + KA_SG_0[f][qq];
                    KV_SG_0[f][qg] = fn * P[S_VMA0_QG][e] + KA_SG_0[f][qg];
                    KV_SG_0[f][gq] = P[S_VMA0_GQ][e] + KA_SG_0[f][gq];
                    KV_SG_0[f][gg] = P[S_VMA0_GG][e] + KA_SG_0[f][gg];
                    /* matevo.f90:437-446 */
                    KV_SG_1[f][qq] = fn * P[S_V1_QQ][e] + KV_NS_1p[f][e];
                    KV_SG_1[f][qg] = fn * P[S_V1_QG][e];
                    KV_SG_1[f][gq] = P[S_V1_GQ][e] + fn * P[S_V1_GQN][e];
                    KV_SG_1[f][gg] = P[S_V1_GG][e] + fn * P[S_V1_GGN][e];
                    /* matevo.f90:463-472 */
                    KA_SG_1[f][qq] = fn * P[S_A1_QQ][e] + KV_NS_1m[f][e];
                    KA_SG_1[f][qg] = fn * P[S_A1_QG][e];
                    KA_SG_1[f][gq] = P[S_A1_GQ][e] + fn * P[S_A1_GQN][e];
                    KA_SG_1[f][gg] = P[S_A1_GG][e] + fn * P[S_A1_GGN][e];
                }
    }
    for (int s = 0; s < N_SLOTS; s++) free(P[s]);
    return 0;
}

int tko_raw_kernel(int which, int nfl, double *K)
{
    if (nfl < 3 || nfl > 5 || !K_NS_0[0]) return 1;
    int f = nfl - 3;
    size_t n = nx_cache, nn = n * n * nxi_cache, nn2 = 4 * nn;
    double *src[7] = { K_NS_0[f], KV_NS_1p[f], KV_NS_1m[f], KV_SG_0[f], KA_SG_0[f], KV_SG_1[f], KA_SG_1[f] };
    if (which < 0 || which > 6) return 1;
    memcpy(K, src[which], sizeof(double) * (which < 3 ? nn : nn2));
    return 0;
}

/* ------------------------------------------------------------ RK4 stepping */

/* C = A B, all m x m column-major; sequential k summation */
static void matmul(int m, const double *A, const double *B, double *C)
{
    for (int j = 0; j < m; j++)
        for (int i = 0; i < m; i++) {
            double s = 0.0;
            for (int k = 0; k < m; k++) s = s + A[(size_t)i + (size_t)m * k] * B[(size_t)k + (size_t)m * j];
            C[(size_t)i + (size_t)m * j] = s;
        }
}

/* matevo.f90:600-673 (rk4_NS and rk4_SG are the same code at m = nx and m = 2 nx) */
static void rk4(int m, double Q2i, double Q2f, const double *K0, const double *K1, double *K)
{
    size_t mm = (size_t)m * m;
    double *W1 = (double *)malloc(sizeof(double) * mm * 8);
    double *W2 = W1 + mm, *W3 = W2 + mm, *W4 = W3 + mm;
    double *M_ini = W4 + mm, *M_mid = M_ini + mm, *M_end = M_mid + mm, *T = M_end + mm;
    double h = tko_log(Q2f / Q2i);
    double Q2m = sqrt(Q2i * Q2f);
    double a_ini = tko_get_alpha_qcd(Q2i) / (2.0 * PI);
    double a_mid = tko_get_alpha_qcd(Q2m) / (2.0 * PI);
    double a_end = tko_get_alpha_qcd(Q2f) / (2.0 * PI);
    for (size_t e = 0; e < mm; e++) {
        M_ini[e] = a_ini * K0[e];
        M_mid[e] = a_mid * K0[e];
        M_end[e] = a_end * K0[e];
    }
    if (l_nlo_cache)
        for (size_t e = 0; e < mm; e++) {
            M_ini[e] = M_ini[e] + SQ(a_ini) * K1[e];
            M_mid[e] = M_mid[e] + SQ(a_mid) * K1[e];
            M_end[e] = M_end[e] + SQ(a_end) * K1[e];
        }
    for (size_t e = 0; e < mm; e++) W1[e] = M_ini[e];
    matmul(m, M_mid, W1, T);
    for (size_t e = 0; e < mm; e++) W2[e] = M_mid[e] + 0.5 * h * T[e];
    matmul(m, M_mid, W2, T);
    for (size_t e = 0; e < mm; e++) W3[e] = M_mid[e] + 0.5 * h * T[e];
    matmul(m, M_end, W3, T);
    for (size_t e = 0; e < mm; e++) W4[e] = M_end[e] + h * T[e];
    for (size_t e = 0; e < mm; e++) K[e] = h / 6.0 * (W1[e] + 2.0 * W2[e] + 2.0 * W3[e] + W4[e]);
    free(W1);
}

/* matevo.f90:485-595: M(:,:,xi,1) = I; M(:,:,xi,iQ2) = matmul(I + rk4(...), M(:,:,xi,iQ2-1)) */
static void make_evomat(int m, int nxi, int nQ2, const double *Q2, double *const K0[3], double *const K1[3], double *M)
{
    size_t mm = (size_t)m * m;
    memset(M, 0, sizeof(double) * mm * nxi * nQ2);
    for (int ixi = 0; ixi < nxi; ixi++)
        for (int ix = 0; ix < m; ix++) M[(size_t)ix + (size_t)m * ix + mm * ixi] = 1.0;
    for (int iQ2 = 1; iQ2 < nQ2; iQ2++) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int ixi = 0; ixi < nxi; ixi++) {
            int nfl = tko_get_neff(Q2[iQ2 - 1]);
            if (nfl > 5) nfl = 5;         /* out of bounds in the reference (nfl_max = 5); clamp */
            double *S = (double *)malloc(sizeof(double) * mm);
            rk4(m, Q2[iQ2 - 1], Q2[iQ2], K0[nfl - 3] + mm * ixi, K1[nfl - 3] + mm * ixi, S);
            for (int ix = 0; ix < m; ix++) S[(size_t)ix + (size_t)m * ix] = 1.0 + S[(size_t)ix + (size_t)m * ix];
            matmul(m, S, M + mm * ((size_t)ixi + (size_t)nxi * (iQ2 - 1)), M + mm * ((size_t)ixi + (size_t)nxi * iQ2));
            free(S);
        }
    }
}

int tko_make_matrices(int nQ2, const double *Q2_array, int l_nlo)
{
    if (!K_NS_0[0] || nQ2 < 1) return 1;
    free_matrices();
    l_nlo_cache = l_nlo ? 1 : 0;
    nQ2_cache = nQ2;
    int n = nx_cache, nxi = nxi_cache;
    size_t nn = (size_t)n * n * nxi * nQ2;
    M_NS_pls = (double *)malloc(sizeof(double) * nn);
    M_NS_min = (double *)malloc(sizeof(double) * nn);
    MV_SG = (double *)malloc(sizeof(double) * nn * 4);
    MA_SG = (double *)malloc(sizeof(double) * nn * 4);
    make_evomat(n, nxi, nQ2, Q2_array, K_NS_0, KV_NS_1p, M_NS_pls);
    make_evomat(n, nxi, nQ2, Q2_array, K_NS_0, KV_NS_1m, M_NS_min);
    make_evomat(2 * n, nxi, nQ2, Q2_array, KV_SG_0, KV_SG_1, MV_SG);
    make_evomat(2 * n, nxi, nQ2, Q2_array, KA_SG_0, KA_SG_1, MA_SG);
    return 0;
}

/* matevo.f90:246-278 */
int tko_evomat(int type, int nstype, double *M)
{
    if (!M_NS_pls) return 1;
    size_t nn = (size_t)nx_cache * nx_cache * nxi_cache * nQ2_cache;
    const double *src = 0;
    switch (type) {
    case 0: src = (nstype == 1) ? M_NS_pls : (nstype == -1 ? M_NS_min : 0); break;
    case 1: src = MV_SG; nn *= 4; break;
    case 2: src = (nstype == 1) ? M_NS_min : (nstype == -1 ? M_NS_pls : 0); break;
    case 3: src = MA_SG; nn *= 4; break;
    }
    if (!src) return 1;
    memcpy(M, src, sizeof(double) * nn);
    return 0;
}

/* matevo.f90:124-241 */
int tko_kernel(int type, double Q2, int nfl, int l_nlo, int i_ns_type, double *K)
{
    if (!K_NS_0[0] || nfl < 3 || nfl > 5) return 1;
    int f = nfl - 3, n = nx_cache, nxi = nxi_cache;
    size_t n2 = (size_t)2 * n;
    double al2pi = tko_get_alpha_qcd(Q2) / (2.0 * PI);
    int isA = type >= 4, blk = type & 3;          /* 0 QQ, 1 QG, 2 GQ, 3 GG */
    double *const *SG0 = isA ? KA_SG_0 : KV_SG_0;
    double *const *SG1 = isA ? KA_SG_1 : KV_SG_1;
    int ro = (blk == 2 || blk == 3) ? n : 0, co = (blk == 1 || blk == 3) ? n : 0;
    for (int iz = 0; iz < nxi; iz++)
        for (int iy = 0; iy < n; iy++)
            for (int ix = 0; ix < n; ix++) {
                size_t e = IDX3(ix, iy, iz, n), e2 = IDX3(ix + ro, iy + co, iz, n2);
                double v;
                if (blk == 0) {
                    v = al2pi * K_NS_0[f][e];
                    if (l_nlo) {
                        int t = isA ? -i_ns_type : i_ns_type;     /* A swaps plus and minus */
                        if (i_ns_type == 0) v = v + SQ(al2pi) * SG1[f][e2];
                        else if (t == 1) v = v + SQ(al2pi) * KV_NS_1p[f][e];
                        else if (t == -1) v = v + SQ(al2pi) * KV_NS_1m[f][e];
                    }
                } else {
                    v = al2pi * SG0[f][e2];
                    if (l_nlo) v = v + SQ(al2pi) * SG1[f][e2];
                }
                K[e] = v;
            }
    return 0;
}
