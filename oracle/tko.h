/* tko.h -- public declarations of the CPU oracle (libtko.so).
 *
 * TEST INFRASTRUCTURE ONLY (see tko_fp.h).  The oracle is a plain-C restatement of
 * the reference algorithm of quantom-collab/tiktaalik for the hot path "assemble
 * and apply GPD evolution matrices".  Every function cites the reference file:line
 * it follows (paths relative to the reference root).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
 * path and it cannot be built here (no Fortran compiler),
 * so this restatement is pinned only by reference-independent known-answer tests
 * (mpmath dilogarithm, SciPy QUADPACK, block identities, polynomiality, sum rules, the
 * forward limit xi -> 0 against the published LO/NLO DGLAP splitting functions, and at
 * xi != 0 the conformal-moment mixing: LO singlet, NLO non-singlet (Mueller 1994)); see
 * tests/test_oracle_known_answers.py, tests/splitting_functions.py and DESIGN.md.
 */
#ifndef TKO_H
#define TKO_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (f90src/qcd/constants.f90:16-39) ---- */
#define TKO_CF (4.0 / 3.0)
#define TKO_CA 3.0
#define TKO_TF 0.5
extern const double tko_pi, tko_zeta2, tko_zeta3;

typedef double (*tko_k3)(double x, double y, double xi);   /* K(x,y,xi) */
typedef double (*tko_k2)(double x, double xi);             /* Kcst(x,xi) */
typedef double (*tko_f1)(double y, void *ctx);             /* integrand with closure */

/* ---- specfun.f90 ---- */
double tko_dilog(double z);

/* ---- kernels_common.f90 ---- */
double tko_theta_step(double X);
double tko_rho_step(double X, double Y);
double tko_erbl_step(double X);
double tko_abslog(double x);
double tko_log2(double x);
double tko_logprod(double x1, double x2);

/* ---- kernels_lo.f90 ---- */
double tko_K0_qq_pls(double x, double y, double xi);
double tko_K0_qq_cst(double x, double xi);
double tko_KA0_qG_reg(double x, double y, double xi);
double tko_KVmA0_qG_reg(double x, double y, double xi);
double tko_KA0_Gq_reg(double x, double y, double xi);
double tko_KVmA0_Gq_reg(double x, double y, double xi);
double tko_KA0_GG_pls(double x, double y, double xi);
double tko_KA0_GG_cst(double x, double xi);
double tko_KA0_GG_nfl(double x, double xi);
double tko_KVmA0_GG_reg(double x, double y, double xi);

/* ---- kernels_nlo.f90 ---- */
double tko_KV1_NSp_pls(double x, double y, double xi);
double tko_KV1_NSm_pls(double x, double y, double xi);
double tko_KV1_NSp_pls_nfl(double x, double y, double xi);
double tko_KV1_NSm_pls_nfl(double x, double y, double xi);
double tko_KV1_qq_reg(double x, double y, double xi);
double tko_KA1_qq_reg(double x, double y, double xi);
double tko_KV1_qG_reg(double x, double y, double xi);
double tko_KA1_qG_reg(double x, double y, double xi);
double tko_KV1_Gq_reg(double x, double y, double xi);
double tko_KA1_Gq_reg(double x, double y, double xi);
double tko_KV1_Gq_reg_nfl(double x, double y, double xi);
double tko_KA1_Gq_reg_nfl(double x, double y, double xi);
double tko_KV1_GG_pls(double x, double y, double xi);
double tko_KA1_GG_pls(double x, double y, double xi);
double tko_KV1_GG_pls_nfl(double x, double y, double xi);
double tko_KA1_GG_pls_nfl(double x, double y, double xi);
double tko_KV1_NSp_cst(double x, double xi);
double tko_KV1_NSm_cst(double x, double xi);
double tko_KV1_NSp_cst_nfl(double x, double xi);
double tko_KV1_NSm_cst_nfl(double x, double xi);
double tko_KV1_GG_cst(double x, double xi);
double tko_KA1_GG_cst(double x, double xi);
double tko_KV1_GG_cst_nfl(double x, double xi);
double tko_KA1_GG_cst_nfl(double x, double xi);
/* the bare plus-prescription integral of one of the 7 distinct constants
 * (which: 0 NSp, 1 NSm, 2 NS_nfl, 3 VGG, 4 AGG, 5 VGG_nfl, 6 AGG_nfl) together with
 * the QAGS bookkeeping of its three sub-integrals (for path-exactness tests). */
double tko_cst_integral(int which, double x, double xi, int last[3], int ifail[3]);
/* two-variable kernel by index 0..22 in the order of tko_kernel_names */
double tko_kernel_by_index(int idx, double x, double y, double xi);
extern const char *const tko_kernel_names[23];

/* ---- gridspace.f90 / pixelation.f90 ---- */
double tko_push_forward(double eta, double xi, int nx, int grid_type);
double tko_pull_back(double x, double xi, int nx, int grid_type);
double tko_push_jacob(double eta, double xi, int nx, int grid_type);
void   tko_initialize_lagrange_weights(int nx, int norder);
double tko_interpixel(int n_pixels, int i, double x, double xi, int grid_type);
void   tko_pixelspace(int nx, double xi, int grid_type, double *x);   /* f90wrap/matrices.f90:21-30 */
void tko_set_variant(int eta_recip_f32, int x0_recip_f32);             /* -Ofast float32-division variants, SURVEY 8c */
double tko_grid_node(int i, int nx, double xi, int grid_type);       /* convolution.f90:45-46 */

/* test hooks for the primitives of tko_fp.h */
double tko_log_x(double x);
double tko_pow_x(double x, double y);
double tko_log_cr_x(double x);   /* the grid maps' logarithm (correctly rounded in practice) */

/* ---- integration.f90 ---- */
void tko_qk21(tko_f1 f, void *ctx, double a, double b,
              double *result, double *abserr, double *resabs, double *resasc);
void tko_qags(tko_f1 f, void *ctx, double a, double b, double epsabs, double epsrel,
              double *result, double *abserr, int *neval, int *ifail, int *last);
double tko_integrate(tko_f1 f, void *ctx, double x, double xi, int n_pixels, int grid_type);
double tko_adaptive_integrate(tko_f1 f, void *ctx, double x, double xi, int last[3], int ifail[3]);

/* ---- convolution.f90 ---- */
double tko_pixel_conv(tko_k3 Kreg, tko_k3 Kpls, tko_k2 Kcst,
                      double xi, int N, int i, int j, int grid_type);

/* ---- alpha_qcd.f90 ---- */
double tko_get_alpha_qcd(double Q2);
int    tko_get_neff(double Q2);

/* ---- matevo.f90: one global context, like the reference's module state ---- */
/* mode 0: unfactored, reference-shaped loop nest (pixel_conv per (i,j)); mode 1: factored
 * (kernel evaluated once per (xi,i,node), same per-term values and summation order). */
int  tko_make_kernels(int nx, int nxi, const double *xi_array, int grid_type, int mode);
int  tko_make_matrices(int nQ2, const double *Q2_array, int l_nlo);
int  tko_get_nx(void);
int  tko_get_nxi(void);
int  tko_get_nQ2(void);
/* getters write Fortran-ordered arrays (first index fastest) */
int  tko_evomat(int type /*0 VNS 1 VSG 2 ANS 3 ASG*/, int nstype, double *M);
int  tko_kernel(int type /*0 VQQ 1 VQG 2 VGQ 3 VGG 4 AQQ 5 AQG 6 AGQ 7 AGG*/,
                double Q2, int nfl, int l_nlo, int i_ns_type, double *K);
/* raw nfl-slice arrays: which 0 K_NS_0, 1 KV_NS_1p, 2 KV_NS_1m (n,n,nxi); 3 KV_SG_0, 4 KA_SG_0,
 * 5 KV_SG_1, 6 KA_SG_1 (2n,2n,nxi) */
int  tko_raw_kernel(int which, int nfl, double *K);
void tko_free_all(void);

/* ---- model/gk.f90 ---- */
double tko_gk(int which /*0 Hu 1 Hd 2 Hs 3 Hg 4 Hu_val 5 Hd_val 6 Hu_sea 7 Hd_sea*/,
              double x, double xi, double t);
void tko_gk_grid(int which, int nx, int nxi, int nt, const double *x, const double *xi,
                 const double *t, double *f /* (nx,nxi,nt) Fortran order */);

#ifdef __cplusplus
}
#endif
#endif
