/* tiktaalik_b200.h -- C ABI of libtiktaalik_b200.so, the B200 (sm_100a) drop-in for the
 * f2py/Fortran layer under tiktaalik.matrices / tiktaalik.evolution.
 *
 * Every entry point replaces one subroutine of the reference's f2py module `dummy`
 * (tiktaalik/f90wrap/matrices.f90, tiktaalik/f90wrap/qcd.f90; paths relative to the
 * reference root).  Plain pointers and sizes only: host arrays for the small inputs
 * (xi, Q2), CUDA device pointers for kernels and matrices, the caller's CUDA stream as
 * void*.  All entry points return 0 on success and non-zero on failure;
 * tkb_last_error() then holds a message (the reference has no error channel below
 * Python: matrices.py:40-41,68,198).
 *
 * State model: like the reference's Fortran module state (matevo.f90:27-46) there is one
 * implicit context per process, bound to the CUDA device that is current at the first
 * tkb_make_kernels call.  Not thread-safe, like the reference.
 *
 * Device layouts (index semantics are the API, strides are free):
 *   evolution matrices  M[xi][Q2][i][j]  row-major m x m tiles, m = nx (NS) or 2 nx (SG);
 *                       logical element (i, j, xi, Q2) of the reference's M(i,j,xi,Q2)
 *   kernel slices       K[xi][i][j]
 *   apply               g[j][xi][s] (the caller's (m, nxi, S) C-contiguous array),
 *                       out[xi][Q2][i][s]
 */
#ifndef TIKTAALIK_B200_H
#define TIKTAALIK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* matrix / kernel type codes */
enum { TKB_VNS = 0, TKB_VSG = 1, TKB_ANS = 2, TKB_ASG = 3 };
enum { TKB_VQQ = 0, TKB_VQG = 1, TKB_VGQ = 2, TKB_VGG = 3, TKB_AQQ = 4, TKB_AQG = 5, TKB_AGQ = 6, TKB_AGG = 7 };

const char* tkb_last_error(void);
int tkb_version(void);

/* replaces make_kernels_wrap (f90wrap/matrices.f90:35-41 -> matevo.f90:64-80, 283-480).
 * Assembles all LO and NLO kernel matrices for nx pixels and nxi skewness values.  Like
 * initialize_2D (matevo.f90:688-689) it clamps xi[0] >= 1e-6 and xi[nxi-1] <= 0.999999.
 * Frees any evolution matrices. */
int tkb_make_kernels(int nx, int nxi, const double* xi_host, int grid_type, void* stream);

/* same, for one rank's slice of a xi list that was clamped globally before sharding:
 * clamp_first / clamp_last say whether this slice holds the global first / last xi. */
int tkb_make_kernels_shard(int nx, int nxi, const double* xi_host, int grid_type,
                           int clamp_first, int clamp_last, void* stream);

/* replaces make_matrices_wrap (f90wrap/matrices.f90:43-53 -> matevo.f90:82-96, 485-673). */
int tkb_make_matrices(int nq2, const double* q2_host, int nlo, void* stream);

/* replace get_nx_wrap / get_nxi_wrap / get_nQ2_wrap (f90wrap/matrices.f90:58-71) */
int tkb_get_nx(void);
int tkb_get_nxi(void);
int tkb_get_nq2(void);
int tkb_get_lnlo(void);

/* replace evomatrix_{vns,vsg,ans,asg}_wrap (f90wrap/matrices.f90:76-110 -> matevo.f90:246-278).
 * tkb_evomatrix returns a zero-copy view (valid until the next make_kernels/make_matrices; the memory itself is
 * only rewritten by the next tkb_make_matrices, so an asynchronous read that is already enqueued may overlap a
 * tkb_make_kernels of the same shapes):
 * shape = (m, m, nxi, nq2) and element strides for the logical indices (i, j, xi, Q2).
 * tkb_evomatrix_copy writes a private copy in the same [xi][Q2][i][j] layout (f2py's
 * intent(out) semantics: the caller may modify it). */
int tkb_evomatrix(int type, int nstype, const double** dev_ptr, int64_t shape[4], int64_t strides[4]);
int tkb_evomatrix_copy(int type, int nstype, double* dev_out, void* stream);

/* replace evokernel_{v,a}{qq,qg,gq,gg}_wrap (f90wrap/matrices.f90:115-193 -> matevo.f90:124-241):
 * dev_out[xi][i][j] = a K0 (+ a^2 K1), a = alpha_s(q2)/(2 pi); nstype in {1,-1,0} for the QQ types. */
int tkb_evokernel(int type, double q2, int nfl, int nlo, int nstype, double* dev_out, void* stream);

/* replaces pixelspace_wrap (f90wrap/matrices.f90:21-30) */
int tkb_pixelspace(int nx, double xi, int grid_type, double* host_out);

/* replaces alpha_wrap (f90wrap/qcd.f90:19-29 -> alpha_qcd.f90:22-50) */
int tkb_alpha(int n, const double* q2_host, double* out_host);

/* replaces np.einsum('ijkl,jk...->ik...l', M, g) (evolution.py:259-348,507-509):
 * out[xi][Q2][i][s] = sum_j M[xi][Q2][i][j] g[j][xi][s]; ldg = row stride of g in elements
 * (nxi*S for a contiguous (m, nxi, S) array). */
int tkb_apply(int m, int nxi, int nq2, int64_t S, const double* M_dev, const double* g_dev, int64_t ldg,
              double* out_dev, void* stream);

/* the same for a block of S samples taken out of a larger (m, nxi, S_total) array without copying it
 * (the inference loop of BASELINE config 4: 10^5 samples do not fit HBM as one output, so they are applied
 * S at a time): g[j][xi][s] is read at g_dev[j*ldg + xi*xig + s], i.e. ldg = nxi*S_total, xig = S_total and
 * g_dev points at the block's first sample.  out is contiguous [xi][Q2][i][S] as above. */
int tkb_apply_strided(int m, int nxi, int nq2, int64_t S, const double* M_dev, const double* g_dev, int64_t ldg,
                      int64_t xig, double* out_dev, void* stream);

/* replaces evolve_fixed_flavor (tiktaalik/evolution.py:495-514) = make_evo_basis (:435-462) + three applications
 * (:507-509) + make_flv_basis (:465-492) + zero padding (:532-540) for one threshold segment at a fixed number of
 * flavours nfl, fused: one decoupling kernel, three applications, one recoupling kernel.
 *   gpd_dev  [F][nx][nxi][S]       flavour basis, row 0 gluon, rows 1.. quarks (u, d, s, ...); missing rows count as zero
 *   M_*_dev  [xi][Q2][m][m]        the segment's singlet (m = 2 nx), plus-type and minus-type non-singlet matrices
 *   out_dev  [Fout][nx][nxi][ldq][S]  receives the nq2 - skip last Q2 points of the segment in Q2 rows q0, q0 + 1, ...
 *                                  (skip = 1 drops the segment's first point, which repeats its input); flavour rows
 *                                  > nfl are zero.  The sample index is the fastest one (coalesced on both sides);
 *                                  the Python layer returns the permuted view (flavour, x, xi, sample, Q2)
 *   tpls_host, tpls_inv_host       the nfl x nfl matrix (u+, d+, ...) -> (Sigma, T3, T8, ...) of evolution.py:517-529 and its
 *                                  inverse (np.linalg.inv in the reference), row-major, host memory
 *   work_dev                       tkb_evolve_workspace(...) doubles of device scratch */
int64_t tkb_evolve_workspace(int nx, int nxi, int nq2, int64_t S, int nfl);
int tkb_evolve_fixed_flavor(int nx, int nxi, int nq2, int64_t S, int nfl, int F, int Fout, const double* tpls_host,
                            const double* tpls_inv_host, const double* M_sg_dev, const double* M_nsp_dev,
                            const double* M_nsm_dev, const double* gpd_dev, double* out_dev, int64_t ldq, int q0, int skip,
                            double* work_dev, void* stream);

/* replaces hu_wrap ... hd_sea_wrap (f90wrap/model.f90:16-158 -> model/gk.f90:75-241): the GK model GPD
 * out[ix][ixi][it] = H_which(x[ix], xi[ixi], t[it]); which = 0 Hu, 1 Hd, 2 Hs, 3 Hg, 4 Hu_val, 5 Hd_val,
 * 6 Hu_sea, 7 Hd_sea.  Input generator of the examples, not part of the timed path. */
int tkb_gk(int which, int nx, int nxi, int nt, const double* x_dev, const double* xi_dev, const double* t_dev,
           double* out_dev, void* stream);

/* -Ofast ambiguity of the reference build: gfortran may have compiled the float32 divisions by the
 * pixel count -- real(2*i-1)/real(n_pixels), convolution.f90:45,67,92, and real(2*iStart-1)/real(Nx_cache),
 * pixelation.f90:85 -- as multiplications by a rounded reciprocal, which moves 1/6 of the nodes by one float32
 * ulp and kernel entries by up to 2e-4.  0 = IEEE division (canonical, the default), 1 = reciprocal multiply;
 * applies to the next tkb_make_kernels.  The oracle has the same switches (oracle/tko_grid.c). */
int tkb_set_variant(int eta_recip_f32, int x0_recip_f32);

/* ---- multi-GPU: one process per GPU, xi-sharded.  Every loop of the reference is independent
 * across xi (matevo.f90:291,325,359,381,426,452,514,550,584), so ranks assemble and step their own block with
 * tkb_make_kernels_shard / tkb_make_matrices and exchange nothing; the one optional collective is the final
 * all-gather of the matrices along xi.  NCCL is bound at run time (dlopen), not at link time.
 * tkb_shard_unique_id: rank 0 creates the 128-byte NCCL id, the host program distributes it (MPI, torch.distributed
 * store, a file); tkb_shard_init: every rank joins (binds the communicator to its current CUDA device);
 * tkb_allgather_matrices: dev_out[sum nxi_per_rank][nq2][m][m] receives every rank's [xi][Q2][i][j] slab in rank
 * order (type / nstype as tkb_evomatrix; nxi_per_rank[rank] must equal this rank's xi count), enqueued on `stream`. */
int tkb_shard_unique_id(char id_out[128]);
int tkb_shard_init(const char id_in[128], int rank, int world);
int tkb_allgather_matrices(int type, int nstype, const int* nxi_per_rank, double* dev_out, void* stream);
/* the same for a xi list sharded by STRIDE (rank r holds xi[r], xi[r + world], ...; every rank the same number of
 * them): this rank's current matrices, local xi k = 0 .. nxi - 1, go to dev_out[(k_first + k) * world + rank][nq2][m][m],
 * so the gathered array is in the order of the global list.  k_first > 0 gathers a later chunk of the rank's xi
 * (a sweep processed in chunks, each chunk's gather running under the next chunk's assembly on another stream). */
int tkb_allgather_matrices_strided(int type, int nstype, int k_first, double* dev_out, void* stream);
int tkb_shard_finalize(void);

/* the exchange by the copy engines instead (no NCCL, no SMs: it runs UNDER the QAGS kernels, which own every SM for the
 * whole assembly): every rank pushes its per-xi slabs into the receive buffers of all ranks with one strided
 * device-to-device copy per peer over NVLink.  tkb_p2p_alloc: this rank's receive buffer of `slot` (0..3, one per
 * matrix family), a cudaMalloc allocation of `bytes` = nxi_total * nq2 * m * m * 8; tkb_p2p_handle: its 64-byte CUDA IPC
 * handle, which the host program hands to every rank (world * 64 bytes, rank order); tkb_p2p_open: maps the peers'
 * buffers; tkb_p2p_push_matrices: this rank's current matrices, local xi k -> [(k_first + k) * world + rank] of every
 * rank's buffer (the order of the global xi list for a strided split), enqueued on `stream`; the data is complete on a
 * rank when every rank's push has finished (synchronise the ranks before reading); tkb_p2p_close unmaps and frees. */
int tkb_p2p_alloc(int slot, int64_t bytes, double** dev_ptr);
int tkb_p2p_handle(int slot, char handle_out[64]);
int tkb_p2p_open(int slot, int rank, int world, const char* handles_all);
int tkb_p2p_push_matrices(int slot, int type, int nstype, int k_first, void* stream);
int tkb_p2p_close(void);

/* releases every device buffer of the implicit context */
int tkb_free(void);

/* ---- diagnostics used by the parity tests and the benchmark ---- */
/* device pointers to the QAGS bookkeeping of the last tkb_make_kernels:
 * arrays [xi][i][c][sub], c = 0 NSp, 1 NSm, 2 NS_nfl, 3 VGG, 4 AGG, 5 VGG_nfl, 6 AGG_nfl; sub = 0..2 */
int tkb_qags_info(const double** res_dev, const int** last_dev, const int** ifail_dev);
/* composed nfl slice, row-major per xi: which = 0 K_NS_0, 1 KV_NS_1p, 2 KV_NS_1m (n x n),
 * 3 KV_SG_0, 4 KA_SG_0, 5 KV_SG_1, 6 KA_SG_1 (2n x 2n) */
int tkb_raw_kernel(int which, int nfl, const double** dev_ptr, void* stream);
/* all 22 two-variable kernels at npts points: out[p][22] */
int tkb_eval_kernels(int npts, const double* x_dev, const double* y_dev, const double* xi_dev, double* out_dev,
                     void* stream);
/* the QAGS integrand K(x,y,xi) - K(y,x,xi) of constant c (index as in tkb_qags_info; kernels_nlo.f90:71-77 and its
 * siblings) at npts points, evaluated straight from the reference formulas (all four half kernels, nothing skipped):
 * the unit of ALGORITHMIC work of the QAGS kernels' roofline, and a parity hook */
int tkb_eval_integrand(int c, int npts, const double* x_dev, const double* y_dev, const double* xi_dev, double* out_dev,
                       void* stream);
/* the task skipping of the assembly kernels, exposed for the tests: for constant c (index as in tkb_qags_info) and
 * the four half-kernel tasks of a node (K(x,y) halves 0/1, K(y,x) halves 0/1): support class (2 full, 1 erbl-only,
 * 0 skipped as an exact zero), value of the full evaluation, value of the erbl-only evaluation; arrays [npts][4].
 * c = 7: all 22 kernels at once as the fixed-quadrature kernel evaluates them; full = number of values that are not
 * an exact zero */
int tkb_task_classes(int c, int npts, const double* x_dev, const double* y_dev, const double* xi_dev, int* cls_dev,
                     double* full_dev, double* erbl_dev, void* stream);
/* which = 0 log(a) (kernels), 1 pow(a,b), 2 dilog(a), 3 log(a) of the grid maps (correctly rounded in practice), with the library's own device arithmetic */
int tkb_eval_scalar(int which, int npts, const double* a_dev, const double* b_dev, double* out_dev, void* stream);
/* device time in ms of the stages of the last tkb_make_kernels / tkb_make_matrices:
 * ms[0] fixed quadrature, ms[1] QAGS constants, ms[2] RK4 stepping (incl. composition) */
int tkb_stage_ms(double ms[3]);
/* measured DFMA throughput of the current device (8 independent FMA chains per thread, no memory
 * traffic): the FP64 roofline denominator, which MEASURED_PEAKS.json does not carry */
int tkb_fp64_peak(double* tflops);
/* number of kernel launches issued by the library since load */
long long tkb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
