// kernels.h -- host-side launchers of the kernels in this directory (internal; the public surface is
// include/mgcmt_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <string>

#include "common.cuh"

namespace mgcmt {

// sets the thread's mgcmt_last_error() text and returns `code` (api.cu)
int set_error(int code, const std::string &msg);

enum { OP_JACOBI = 0, OP_RESIDUAL = 1, OP_APPLY = 2, OP_RAYLEIGH = 3 };

// number of kernels this library has launched since it was loaded (reported by bench.py as gpu_launches)
extern long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }

// stencil.cu
cudaError_t launch_jacobi_sweep(const LevelDev &L, double shift, double omega, const double *v_in,
                                const double *f, double *v_out, const double *halo_top,
                                const double *halo_bot, cudaStream_t s);
cudaError_t launch_residual(const LevelDev &L, double shift, const double *v, const double *f, double *r,
                            const double *halo_top, const double *halo_bot, cudaStream_t s);
cudaError_t launch_apply(const LevelDev &L, double shift, const double *x, double *y, const double *halo_top,
                         const double *halo_bot, cudaStream_t s);

cudaError_t launch_rayleigh_partials(const LevelDev &L, const double *x, double *partials, const double *halo_top,
                                     const double *halo_bot, cudaStream_t s);
int march_grid_blocks(const LevelDev &L);

// transfer.cu
cudaError_t launch_galerkin_tridiag(int n_fine, const double *lo, const double *di, const double *up,
                                    double *clo, double *cdi, double *cup, cudaStream_t s);
cudaError_t launch_restrict(const LevelDev &Lf, bool coarsen_rows, const double *fine, double *coarse,
                            cudaStream_t s);
cudaError_t launch_residual_restrict(const LevelDev &Lf, bool coarsen_rows, double shift, const double *v,
                                     const double *f, double *rc, cudaStream_t s);
cudaError_t launch_prolong(const LevelDev &Lf, bool coarsen_rows, bool accumulate, const double *coarse,
                           double *fine, cudaStream_t s);

// fused.cu: one pass = nu (0..4) Jacobi sweeps fused with the neighbouring transfer (2-D levels)
enum { FUSED_SMOOTH = 0,     // v_out = J^nu(v_in)
       FUSED_DOWN = 1,       // v_out = J^nu(v_in), r_coarse = R (f - A v_out)
       FUSED_DOWN_ZERO = 2,  // same with v_in == 0 (not read)
       FUSED_UP = 3,         // v_out = J^nu(v_in + P e_coarse)
       FUSED_UP_RQ = 4 };    // FUSED_UP (Jacobi, nu = 4, 5-point level) + per-warp partials of w^T A_s w, w^T w of v_out
#ifndef MGCMT_FUSED_C5
#define MGCMT_FUSED_C5 4     // columns per lane, 5-point (finest) level
#endif
#ifndef MGCMT_FUSED_C9
#define MGCMT_FUSED_C9 2     // columns per lane, 9-point (Galerkin) levels
#endif
extern int g_fused_c5, g_fused_c9, g_fused_skew_cols;
int fused_rq_slots(const LevelDev &L, int gs = 0);
// fused_uni.cu: the same legs for constant-coefficient 5-point levels (LevelDev::uni), half the fp64 instructions
extern int g_fused_uni, g_uni_minctas, g_uni_wfreg, g_uni_bulk;
bool uni5_available(const LevelDev &L);
cudaError_t launch_uni5_leg(const LevelDev &L, int gs, int mode, int nu, double shift, double omega, const double *v_in,
                            const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s,
                            int *slots_out = nullptr);
int uni5_rq_slots(const LevelDev &L, int gs);
void uni5_coefficients(double c, double d, double shift, double omega, double *out7);
// fused_uni9.cu: the same for the Galerkin (9-point) levels with constant interior coefficients (LevelDev::uni == 2)
extern int g_fused_uni9, g_uni9_lag, g_uni9_min_cols;
bool uni9_available(const LevelDev &L);
cudaError_t launch_uni9_leg(const LevelDev &L, int gs, int mode, int nu, double shift, double omega, const double *v_in,
                            const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s);
// chunk height of a streaming leg so that gx * chunks CTAs fill whole waves of `slots` resident CTAs
int leg_rows_per_chunk(int nrows, int gx, int slots, int nstage, int max_rpc);
int num_sms();
extern int g_leg_min_rpc;
cudaError_t launch_fused_leg(const LevelDev &L, int mode, int nu, double shift, double omega,
                             const double *v_in, const double *f, double *v_out, const double *e_coarse,
                             double *r_coarse, cudaStream_t s);

// tile.cu: shared-memory tile version of the fused legs (mid-size levels; same modes, nu 0..4 at run
// time), and the single-CTA kernel that runs all levels first..coarsest of a V-cycle in one launch
cudaError_t launch_tile_leg(const LevelDev &L, int mode, int nu, double shift, double omega, const double *v_in,
                            const double *f, double *v_out, const double *e_coarse, double *r_coarse,
                            cudaStream_t s);
// Gauss-Seidel / SOR legs on shared-memory tiles: up to 4 colour sweeps + transfer in one launch (small levels)
cudaError_t launch_tile_gs_leg(const LevelDev &L, int mode, int sweeps, double shift, double omega, const double *v_in,
                               const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s);
constexpr int kTailMaxLevels = 12;
constexpr size_t kTailMaxSmem = 216 * 1024;
size_t tail_smem_bytes(const LevelDev *levels, int nlev);
cudaError_t launch_tail(const LevelDev *levels, int nlev, bool coarsen_rows, const double *inv, double shift,
                        double omega, const double *f_first, double *v_first, cudaStream_t s);

cudaError_t launch_fused_gs_leg(const LevelDev &L, int mode, int sweeps, double shift, double omega,
                                const double *v_in, const double *f, double *v_out, const double *e_coarse,
                                double *r_coarse, cudaStream_t s);

// gs.cu
cudaError_t launch_rbgs(const LevelDev &L, double shift, double omega, int nu, double *v, const double *f,
                        cudaStream_t s);
cudaError_t launch_gs_lex(const LevelDev &L, double shift, double omega, int nu, double *v, const double *f,
                          double *scratch, cudaStream_t s);

// coarse.cu
cudaError_t launch_build_dense(const LevelDev &L, double shift, double *aug, cudaStream_t s);
cudaError_t launch_gauss_jordan(int n, double *aug, int *status, double *mult, cudaStream_t s);
cudaError_t launch_extract_inverse(int n, const double *aug, double *inv, cudaStream_t s);
cudaError_t launch_gemv(int n, const double *inv, const double *x, double *y, cudaStream_t s);
// the same inverse through a banded LU with partial pivoting (coarsest levels of thousands of unknowns)
size_t band_workspace_doubles(int n, int kl, int ku);
bool band_inverse_fits(const LevelDev &L);
cudaError_t launch_band_inverse2d(const LevelDev &L, double shift, double *inv, int *status, double *work, cudaStream_t s);

// reduce.cu
constexpr int kReduceBlocks = 592;  // 148 SMs x 4
cudaError_t launch_dot(long long n, const double *x, const double *y, double *partials, double *out,
                       cudaStream_t s);
// out[m] = <x0 + m*stride, y>, m < M <= 16; partials: M * kReduceBlocks doubles of scratch
cudaError_t launch_multidot(long long n, int M, const double *x0, long long stride, const double *y,
                            double *partials, double *out, cudaStream_t s);
// out[m] = ordered sum of partials[m*B .. m*B+B)
cudaError_t launch_finish(int M, int B, const double *partials, double *out, cudaStream_t s);
// fused modified Gram-Schmidt passes (MGCMTProcessor.py:44-50), see reduce.cu
cudaError_t launch_mgs_scale_dots(long long n, int m, double *wi, const double *sumsq, const double *wj0,
                                  long long stride, double *partials, double *out, cudaStream_t s);
cudaError_t launch_mgs_update(long long n, int m, const double *qi, const double *dots, double *wj0, long long stride,
                              double *partials, double *out_sumsq, cudaStream_t s);
// Gram-matrix (Cholesky-QR) orthonormalisation pieces: packed upper Gram matrix (k(k+1)/2 dots), the inverse of its
// Cholesky factor, and Q = W R^-1 in place; k <= 6
cudaError_t launch_gram(long long n, int k, const double *w0, long long stride, double *partials, double *out, cudaStream_t s);
cudaError_t launch_chol_inverse(int k, const double *g, double *rinv, int *status, cudaStream_t s);
cudaError_t launch_cholqr_apply(long long n, int k, double *w0, long long stride, const double *rinv, cudaStream_t s);
cudaError_t launch_scale_by_inv_norm(long long n, double *x, const double *sumsq, cudaStream_t s);
cudaError_t launch_axpy_dev(long long n, const double *alpha, const double *denom, double sign,
                            const double *x, double *y, cudaStream_t s);
cudaError_t launch_rq_unshift(double *out2, double shift, cudaStream_t s);
cudaError_t launch_axpby(long long n, double a, const double *x, double b, const double *y, double *out, cudaStream_t s);
cudaError_t launch_scale_to(long long n, const double *x, const double *sumsq, double *y, cudaStream_t s);

// rq.cu: device-resident Rayleigh-quotient minimisation (MGCMTSolver.rqmin); scal: 32 device doubles, partials: 8 * kReduceBlocks
cudaError_t launch_rq_pencil(long long n, const double *x, const double *p, const double *Ax, const double *Ap,
                             const double *Mx, const double *Mp, double *partials, double *scal, cudaStream_t s);
cudaError_t launch_rq_update(long long n, bool mass, double *x, const double *p, double *Ax, const double *Ap, double *Mx,
                             const double *Mp, double *partials, double *scal, cudaStream_t s);
cudaError_t launch_rq_sums(long long n, const double *x, const double *Ax, const double *Mx, double *partials, double *scal,
                           cudaStream_t s);
cudaError_t launch_rq_grad(long long n, bool mass, const double *Ax, const double *Mx, double *g, double *partials, double *scal,
                           cudaStream_t s);
cudaError_t launch_rq_dir(long long n, bool first, const double *scal, const double *g, double *p, cudaStream_t s);

// band.cu: general banded complex128 operators (1-D multiband Hamiltonians); vectors are complex interleaved
constexpr int kBandMaxCoarse = 512;  // largest coarsest level the dense complex solve takes
constexpr int kBandMaxDiags = 96;
extern int g_band_gs_scan, g_band_gs_split;
cudaError_t launch_band_apply(const BandDev &L, double shift, const double *x, double *y, cudaStream_t s);
cudaError_t launch_band_jacobi(const BandDev &L, double shift, double omega, const double *vin, const double *f,
                               double *vout, cudaStream_t s);
cudaError_t launch_band_residual_restrict(const BandDev &L, double shift, const double *v, const double *f, double *rc,
                                          cudaStream_t s);
cudaError_t launch_band_prolong_correct(int n_fine, const double *ec, double *v, cudaStream_t s);
cudaError_t launch_band_galerkin(const BandDev &F, int nc, int ndiag_c, const int *offs_c, const int *lut, double *vals_c,
                                 cudaStream_t s);
cudaError_t launch_band_lower_solve(const BandDev &L, double shift, double wl, double cf, double cd, double cu,
                                    double oscale, const double *vin, const double *f, double *y, const double *g,
                                    double *vout, double *rhs_scratch, cudaStream_t s);
cudaError_t launch_band_inverse(const BandDev &L, double shift, double *aug, int *status, cudaStream_t s);
cudaError_t launch_band_gemv(int n, const double *aug, const double *f, double *y, cudaStream_t s);

// staging.cu: upload of pageable host memory through page-locked chunks filled by several host threads (returns when every
// chunk is enqueued on `stream`); page-locked sources go straight to cudaMemcpyAsync
extern int g_stage_threads, g_stage_chunk_kib;
bool host_pointer_is_pinned(const void *p);
cudaError_t staged_upload(void *d_dst, const void *h_src, size_t bytes, cudaStream_t stream);

}  // namespace mgcmt
