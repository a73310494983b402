/*
 * mgcmt_b200.h -- C ABI of the B200-native multigrid V-cycle path (libmgcmt_b200.so).
 *
 * The reference (AndyMN/MultigridCMT) has no FFI layer: its boundary is the Python class surface
 * MGCMTStencilMaker / MGCMTSolver / MGCMTProcessor.  The host-side mirror of those classes lives in
 * multigridcmt_b200/*.py and binds the entry points below with ctypes (see INTEGRATION.md for the
 * stub a reference maintainer would add).  Each entry point cites the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - plain C types only; every `double*` named `d_*` is a DEVICE pointer (fp64), every `h_*` is
 *     a HOST pointer; `stream` is a cudaStream_t passed as void*.
 *   - no ownership transfer: callers own every buffer they pass in; the hierarchy owns its
 *     coefficient arrays, coarse-level work vectors and cached coarse inverses.
 *   - every function returns MGCMT_OK (0) or an error code; mgcmt_last_error() gives the message
 *     of the last failure on the calling thread.  Nothing here synchronises the host except where
 *     stated (scalar results are written to device memory, not returned).
 *   - grids: a level is an nrows x ncols array of doubles, row-major (index i*ncols + j), exactly the
 *     reference's 2-D ordering (MGCMTStencilMaker.py:24, kronsum => i*N + j).  A 1-D problem is a
 *     grid with nrows == 1.  ncols (and nrows when > 1) are powers of two.
 *   - operators are kept in separable form  A_l = Ma_l (x) Kb_l + Ka_l (x) Mb_l  with tridiagonal
 *     1-D factors (Ka/Ma act on the row index, Kb/Mb on the column index, M_0 = I).  Galerkin
 *     coarsening R A P of the reference (MGCMTSolver.py:318) maps each factor to R T P exactly, so
 *     every level costs O(n) coefficient storage instead of a sparse matrix.
 *   - the shift of the shift method is re-applied as  -shift*I  on every level and A is coarsened
 *     unshifted (MGCMTSolver.py:287-288,318-320).
 */
#ifndef MGCMT_B200_H
#define MGCMT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MGCMT_ABI_VERSION 1

enum {
  MGCMT_OK = 0,
  MGCMT_ERR_ARG = 1,   /* bad argument (size not a power of two, null pointer, bad level ...) */
  MGCMT_ERR_CUDA = 2,  /* a CUDA runtime call or kernel launch failed */
  MGCMT_ERR_STATE = 3, /* hierarchy not usable for this call */
  MGCMT_ERR_NUMERIC = 4 /* singular coarsest operator */
};

/* smoother selectors for mgcmt_smooth / mgcmt_vcycle */
enum {
  MGCMT_SMOOTH_WJACOBI = 0, /* MGCMTSolver.wjacobi, MGCMTSolver.py:182-208 */
  MGCMT_SMOOTH_RBGS = 1,    /* red-black (four-colour) Gauss-Seidel/SOR: the working replacement of the
                               reference's dead gseidelrb, MGCMTSolver.py:248-279 */
  MGCMT_SMOOTH_GSLEX = 2    /* lexicographic Gauss-Seidel / SOR, MGCMTSolver.py:210-246 (incl. quirk Q6) */
};

typedef struct mgcmt_hier mgcmt_hier_t; /* opaque grid hierarchy */

int mgcmt_abi_version(void);
const char *mgcmt_last_error(void);

/* ---- instrumentation for bench.py ------------------------------------------------------------------
 * mgcmt_launch_count: kernels launched by this library since load (bench.py's gpu_launches).
 * mgcmt_profile_enable(1): bracket every finest-level smoother launch with CUDA events on its stream;
 * mgcmt_profile_read: synchronise, return the summed event time (ms) and the number of bracketed
 * intervals since the last read, and reset. */
long long mgcmt_launch_count(void);
int mgcmt_profile_enable(int on);
int mgcmt_profile_read(double *ms_total, long long *intervals);
/* the same per kind of finest-level launch (arrays of MGCMT_PROF_KINDS entries): the legs of a V-cycle move different
 * amounts of data (zero-start down leg 18 B, down / up leg 26 B per unknown), so a roofline fraction is per kind */
#define MGCMT_PROF_SWEEP 0      /* a single smoother sweep / un-fused smoother call */
#define MGCMT_PROF_DOWN_ZERO 1  /* fused down leg, zero initial guess (v not read) */
#define MGCMT_PROF_DOWN 2       /* fused down leg */
#define MGCMT_PROF_UP 3         /* fused up leg */
#define MGCMT_PROF_UP_RQ 4      /* fused up leg that also leaves the Rayleigh sums (mgcmt_vcycle_rq) */
#define MGCMT_PROF_KINDS 5
int mgcmt_profile_read_kinds(double *ms_by_kind, long long *intervals_by_kind);

/* ---- hierarchy -------------------------------------------------------------------------------
 * Builds all levels from the finest grid (nrows x ncols) down to the level whose column count is
 * `lowest_level` (the reference's lowest_level, MGCMTSolver.py:305).  The finest operator is
 *   A_0 = I (x) Kb + Ka (x) I,  Ka = tridiag(h_row_lo, h_row_di, h_row_up) (nrows entries each; lo[0]
 *   and up[nrows-1] ignored),  Kb likewise with ncols entries.  For a 1-D problem pass nrows = 1,
 *   h_row_* = {0}.  Coarse factors are the Galerkin products with the reference's transfer operators
 *   (MGCMTStencilMaker.py:27-78: coarse j <-> fine 2j+1, R = [1/4 1/2 1/4] with a truncated last row),
 *   computed on the device.  coarsen_rows = 0 keeps the row count on all levels (1-D problems).
 * Replaces: MGCMTStencilMaker.laplacian/restriction/interpolation + the `R*A*P` at
 * MGCMTSolver.py:310-311,318 that the reference rebuilds on every call. */
int mgcmt_hier_create(mgcmt_hier_t **out, int nrows, int ncols, int coarsen_rows,
                      const double *h_row_lo, const double *h_row_di, const double *h_row_up,
                      const double *h_col_lo, const double *h_col_di, const double *h_col_up,
                      int lowest_level, void *stream);
/* As mgcmt_hier_create, but levels finer than first_work_level get no work vectors (only their operators):
 * the replicated coarse part of a slab-decomposed solver, driven through mgcmt_vcycle_from. */
int mgcmt_hier_create2(mgcmt_hier_t **out, int nrows, int ncols, int coarsen_rows,
                       const double *h_row_lo, const double *h_row_di, const double *h_row_up,
                       const double *h_col_lo, const double *h_col_di, const double *h_col_up,
                       int lowest_level, int first_work_level, void *stream);
/* Row-slab piece of an nrows_glob x ncols grid for multi-GPU runs: levels 0..nlevels-1
 * hold the rows [row_begin >> l, (row_begin + nrows_own) >> l) of level l plus `halo` (even; >= 6 for the Jacobi legs, >= 10 for the red-black ones) rows above
 * and below, as one dense (own + 2 halo) x ncols array; halo rows outside the global grid stay zero.  The
 * operator factors are the full global ones.  Valid calls on such a handle: the single-level operators,
 * mgcmt_fused_leg (streaming implementation; restriction from the last slab level writes into / prolongation
 * reads from a FULL coarse array, i.e. the replicated coarse grid) and mgcmt_slab_rayleigh.  Keeping the
 * halo rows current between legs (NCCL send/recv, multigridcmt_b200/slab.py) is the caller's job. */
int mgcmt_hier_create_slab(mgcmt_hier_t **out, int nrows_glob, int ncols, int row_begin, int nrows_own,
                           int nlevels, int halo, const double *h_row_lo, const double *h_row_di,
                           const double *h_row_up, const double *h_col_lo, const double *h_col_di,
                           const double *h_col_up, void *stream);
/* the hierarchy's own v / f / scratch vectors of a level (NULL where not allocated) */
int mgcmt_hier_level_buffers(mgcmt_hier_t *h, int level, double **d_v, double **d_f, double **d_tmp);
int mgcmt_hier_destroy(mgcmt_hier_t *h);
int mgcmt_hier_num_levels(const mgcmt_hier_t *h);
/* grid size of a level (level 0 = finest) */
int mgcmt_hier_level_shape(const mgcmt_hier_t *h, int level, int *nrows, int *ncols);
/* copies the 12 tridiagonal factor arrays of a level to the host, for tests:
 * order ka_lo,ka_di,ka_up,ma_lo,ma_di,ma_up (nrows each) then kb_*,mb_* (ncols each). Synchronises. */
int mgcmt_hier_level_coefs(const mgcmt_hier_t *h, int level, double *h_rowcoef6, double *h_colcoef6);

/* ---- single-level operators (all on `stream`, device pointers, level sizes) --------------------- */
/* y = (A_l - shift I) x            -- the `shifted_matrix * v` of MGCMTSolver.py:315 */
int mgcmt_apply(mgcmt_hier_t *h, int level, double shift, const double *d_x, double *d_y, void *stream);
/* y = M_l x with M_l = Ma_l (x) Mb_l, the Galerkin-coarsened mass matrix R..R I P..P of the RQ multigrid
 * (`M_coarse = R*M*P`, MGCMTSolver.py:79,111); the identity on the finest level */
int mgcmt_apply_mass(mgcmt_hier_t *h, int level, const double *d_x, double *d_y, void *stream);
/* r = f - (A_l - shift I) v        -- MGCMTSolver.py:315 (inner bracket) */
int mgcmt_residual(mgcmt_hier_t *h, int level, double shift, const double *d_v, const double *d_f,
                   double *d_r, void *stream);
/* nu sweeps of the chosen smoother on (A_l - shift I) v = f, in place in d_v.
 * d_tmp: scratch of the level's size (may be NULL: the hierarchy's own scratch is used).
 * omega: Jacobi weight (reference default 2/3) or SOR factor (reference default 1).
 * Replaces MGCMTSolver.wjacobi/gseidel/sor, MGCMTSolver.py:182-246. */
int mgcmt_smooth(mgcmt_hier_t *h, int level, int smoother, double shift, double omega, int nu,
                 double *d_v, const double *d_f, double *d_tmp, void *stream);
/* coarse = R fine   (full weighting, MGCMTStencilMaker.py:57-78) level -> level+1 */
int mgcmt_restrict(mgcmt_hier_t *h, int level, const double *d_fine, double *d_coarse, void *stream);
/* r_coarse = R (f - (A_l - shift I) v), fused         -- MGCMTSolver.py:315 */
int mgcmt_residual_restrict(mgcmt_hier_t *h, int level, double shift, const double *d_v,
                            const double *d_f, double *d_rcoarse, void *stream);
/* fine = P coarse   (linear interpolation, MGCMTStencilMaker.py:27-54) level+1 -> level */
int mgcmt_prolong(mgcmt_hier_t *h, int level, const double *d_coarse, double *d_fine, void *stream);
/* v += P e, fused                                     -- MGCMTSolver.py:323-324 */
int mgcmt_prolong_correct(mgcmt_hier_t *h, int level, const double *d_ecoarse, double *d_v, void *stream);
/* v = (A_L - shift I)^-1 f on the coarsest level L     -- spsolve at MGCMTSolver.py:305-308.
 * The dense inverse is computed on the device once per shift (pivoted Gauss-Jordan) and cached. */
int mgcmt_coarse_solve(mgcmt_hier_t *h, double shift, const double *d_f, double *d_v, void *stream);

/* ---- whole cycle ---------------------------------------------------------------------------------
 * One V-cycle of MGCMTSolver.vcycle (MGCMTSolver.py:281-329) on the finest level: d_v in/out, d_f in.
 * nu1/nu2 apply to the finest level only; all coarser levels run 4/4 (MGCMTSolver.py:320, quirk Q4).
 * If the hierarchy has a single level this is the exact solve (quirk Q7).
 * v0_is_zero != 0: the caller guarantees the initial guess is zero (the shift-method drivers always
 * start from w0 = 0, e.g. 2DPotGS.py:94); d_v is then not read. */
int mgcmt_vcycle(mgcmt_hier_t *h, double shift, int nu1, int nu2, int smoother, double omega,
                 double *d_v, const double *d_f, int v0_is_zero, void *stream);

/* The loop body of the shift-method drivers for HOST vectors (2DPotGS.py:93-95: one zero-start V-cycle per eigenvector,
 * numpy in, numpy out): v_host[c] = Vcycle(0, f_host[c]; shifts[c]) for c < k, all vectors of the finest level's length.
 * The k cycles are independent, so their PCIe copies are pipelined against each other: the upload of vector c+1 and the
 * download of vector c-1 run on two internal copy streams beside cycle c on `stream` (three rotating device slots owned
 * by the hierarchy).  Host buffers may be pageable or page-locked; page-locked ones move at full PCIe rate in both
 * directions at once.  Returns when every v_host[c] is complete (the call synchronises, like the reference's). */
int mgcmt_vcycle_host_block(mgcmt_hier_t *h, int k, const double *shifts, int nu1, int nu2, int smoother, double omega,
                            const double *const *f_host, double *const *v_host, void *stream);

/* mgcmt_vcycle followed by the Rayleigh quotient of the result: d_out2[0] = w^T A_0 w, d_out2[1] = w^T w (the
 * `w/||w||` + `v^T H v` the drivers do after every cycle, 2DPotGS.py:96,103).  On 2-D Jacobi cycles with nu2 = 4 the
 * sums are taken inside the finest up leg (an extra pipeline stage evaluates A w on the final iterate), so no extra
 * pass over w is made; otherwise it is mgcmt_vcycle + mgcmt_rayleigh. */
int mgcmt_vcycle_rq(mgcmt_hier_t *h, double shift, int nu1, int nu2, int smoother, double omega, double *d_v,
                    const double *d_f, int v0_is_zero, double *d_out2, void *stream);
/* The V-cycle restricted to levels level..coarsest with a zero initial guess and 4/4 sweeps -- what every
 * coarse level of MGCMTSolver.vcycle runs (MGCMTSolver.py:316-320).  d_v, d_f: vectors of that level. */
int mgcmt_vcycle_from(mgcmt_hier_t *h, int level, double shift, int smoother, double omega, double *d_v,
                      const double *d_f, void *stream);
/* slab piece: d_out2[0] = x^T A x, d_out2[1] = x^T x summed over the OWNED rows of this rank only (the caller
 * all-reduces); d_x is the slab array including its halo rows */
int mgcmt_slab_rayleigh(mgcmt_hier_t *h, int level, const double *d_x, double *d_out2, void *stream);
/* slab piece, finest level: the up leg (mgcmt_fused_leg mode 3, nu = 4; gs != 0: four red-black sweeps) with the Rayleigh
 * sums of its result taken in the same pass: d_out2[0] = w^T A w, d_out2[1] = w^T w over the OWNED rows of this rank (the caller all-reduces).
 * Needs no exchange of the result's halo rows: the leg already has the exact neighbouring rows in its pipeline. */
int mgcmt_slab_up_rq(mgcmt_hier_t *h, int gs, double shift, double omega, const double *d_vin, const double *d_f, double *d_vout,
                     const double *d_ecoarse, double *d_out2, void *stream);

/* One fused pass over a 2-D level (what mgcmt_vcycle is made of when the smoother is weighted Jacobi):
 * nu (0..4) Jacobi sweeps fused with the neighbouring grid transfer, out of place (d_vin != d_vout).
 *   mode 0: d_vout = J^nu(d_vin)                                       MGCMTSolver.py:313 / :326
 *   mode 1: d_vout = J^nu(d_vin),  d_rcoarse = R (f - A_s d_vout)      MGCMTSolver.py:313-315
 *   mode 2: as mode 1 with d_vin == 0 (d_vin is not read; the zero start of every coarse level, :316)
 *   mode 3: d_vout = J^nu(d_vin + P d_ecoarse)                         MGCMTSolver.py:323-326
 * For nu == 0, modes 1/2 leave d_vout untouched (the residual is taken of d_vin).
 * mode | 32: the nu sweeps are four-colour (red-black on the 5-point level) Gauss-Seidel/SOR sweeps instead of Jacobi
 *   sweeps (omega = SOR factor; 1..4 sweeps per pass on the 5-point level, 1..2 on the 9-point levels).
 * mode | 16 selects the shared-memory tile implementation (used for mid-size levels) instead of the
 * register-streaming one; both compute the same thing (mode | 16 | 32: Gauss-Seidel sweeps on tiles, 1..4 per pass). */
int mgcmt_fused_leg(mgcmt_hier_t *h, int level, int mode, int nu, double shift, double omega,
                    const double *d_vin, const double *d_f, double *d_vout, const double *d_ecoarse,
                    double *d_rcoarse, void *stream);
/* runtime switches, for tests and A/B timing: "fused" (1/0), "fused_min_cols" (smallest level width
 * that uses the fused legs), "tile_max_cols" (levels at most this wide use the tile legs),
 * "tail_max_cols" (levels at most this wide are collapsed into the single-CTA tail kernel; 0 = off),
 * "fused_c5" / "fused_c9" (columns per lane of the general legs), "leg_min_rpc" (smallest chunk height),
 * "fused_uni" (1/0: constant-coefficient 5-point legs, fused_uni.cu) with its variants "uni_wfreg" (1/0), "uni_minctas"
 * (0|2|3), "uni_bulk" (1/0: row ring fed by cp.async.bulk + mbarrier); "fused_uni9" (0 | 1 all | 2 auto: 9-point legs of
 * fused_uni9.cu on levels >= "uni9_min_cols" wide), "uni9_lag" (1|2); "fused_skew_cols", "tile_gs_max_cols" (alternative
 * Gauss-Seidel leg designs, 0 = off); "coarse_banded" (0 | 1 | 2 auto: coarsest inverse by banded LU);
 * "band_gs_scan" (1/0: scan form of the in-chunk recurrence of the banded Gauss-Seidel sweep), "band_gs_split" (1/0:
 * old-value part of that sweep in its own launch).  DESIGN.md sections 3 / 3c say what each one measured. */
int mgcmt_set_option(const char *name, int value);
/* Host-only views of two decisions the legs make, so that they can be tested without a GPU:
 * the coefficients of fused_uni.cu for a constant 5-point stencil (off-diagonal c, unshifted diagonal d):
 * h_out7 = {1 - om_hi, -om_hi, -om_lo, -beta, beta / c, c / beta, d + 4c - shift} with om_hi + om_lo = beta (d - shift) / c to
 * double-double accuracy ("Exact arithmetic"); and the chunk height chosen for a streaming leg. */
int mgcmt_debug_uni_coefficients(double c, double d, double shift, double omega, double *h_out7);
int mgcmt_debug_leg_rows_per_chunk(int nrows, int gx, int slots, int nstage, int max_rpc);
/* Test hook: one kernel that fills the shared memory of every SM with NaN bit patterns (shared memory is not cleared
 * between kernels).  The streaming legs read ring rows in front of a chunk before anything has filled them; whatever
 * they compute from those must never reach a result (tests/test_gpu_parity.py: Rayleigh sums after a poisoning). */
int mgcmt_debug_poison_shared_memory(void *stream);
/* the phase table of one native slab-block cycle (csrc/slab_block.cu: build_phases) for nlev_slab distributed levels:
 * h_kinds[i] = 0 down leg, 1 / 2 first / second pass of a two-pass (red-black, 9-point) down leg, 3 replicated coarse
 * part, 4 up leg, 5 / 6 first / second pass of a two-pass up leg, 7 separate Rayleigh pass; h_levels[i] its level.
 * Every phase is one halo exchange followed by one leg per vector.  Returns the number of phases (-1: bad arguments or
 * capacity too small).  Host only. */
int mgcmt_debug_slab_phases(int nlev_slab, int smoother, int with_rq_stage, int *h_kinds, int *h_levels, int capacity);

/* ---- reductions / vector post-processing (MGCMTProcessor.py, Rayleigh quotients in the drivers) --
 * Deterministic: fixed two-stage tree, independent of launch timing.  Results go to DEVICE memory. */
int mgcmt_dot(long long n, const double *d_x, const double *d_y, double *d_out, void *stream);
/* d_out[0] = x^T (A_0 x), d_out[1] = x^T x   (one fused pass; e.g. 2DPotGS.py:103) */
int mgcmt_rayleigh(mgcmt_hier_t *h, int level, const double *d_x, double *d_out2, void *stream);
/* x *= 1/||x||_2          -- `w / np.linalg.norm(w)`, e.g. 2DPotGS.py:96; MGCMTProcessor.normalize */
int mgcmt_normalize(long long n, double *d_x, void *stream);
/* the two halves of modified == 2 for vectors `stride` doubles apart (slab-decomposed blocks: the packed upper Gram
 * matrix d_out[k(k+1)/2] of the owned rows is all-reduced between the two calls) */
int mgcmt_gram(long long n, int k, const double *d_V, long long stride, double *d_out, void *stream);
int mgcmt_cholqr_apply(long long n, int k, double *d_V, long long stride, const double *d_gram, void *stream);
/* Breakdown flag of the Gram-matrix orthonormalisation (modified == 2, mgcmt_cholqr_apply): *h_flag = 1 if, since the
 * last call, a Gram matrix was not positive definite (nearly dependent columns -- e.g. two shifts converging on the
 * same eigenvector); the affected block is then NOT orthonormal and must be redone column by column (modified = 1).
 * Synchronises `stream`, resets the flag. */
int mgcmt_ortho_status(int *h_flag, void *stream);
/* x /= sqrt(*d_sumsq) with the sum of squares read from device memory (e.g. after an all-reduce of per-rank
 * partial sums in the slab-decomposed path) */
int mgcmt_scale_inv_norm(long long n, double *d_x, const double *d_sumsq, void *stream);
/* y = alpha*x + y with alpha read from device memory, scaled by `sign` */
int mgcmt_axpy_dev(long long n, const double *d_alpha, double sign, const double *d_x, double *d_y,
                   void *stream);
/* Eigen-residual with the Rayleigh quotient kept on the device: d_r = A_0 x - (d_rq2[0] / d_rq2[1]) x and
 * d_out_sumsq[0] = ||d_r||^2  (d_rq2 as written by mgcmt_rayleigh / mgcmt_vcycle_rq).  The reference never forms it
 * (no convergence criterion, `2DPotGS.py:91-105` runs a fixed count); it is the convergence measure
 * ||H v - rho v|| of ShiftMethod.solve and the right-hand side of its correction form. */
int mgcmt_eigen_residual(mgcmt_hier_t *h, int level, const double *d_x, const double *d_rq2, double *d_r,
                         double *d_out_sumsq, void *stream);
/* `nu` steps of the Rayleigh-quotient minimisation `rqmin(A, v0, M, nu)` (MGCMTSolver.py:17-57) for the level operator
 * A_level and the level mass matrix M_level = Ma (x) Mb (the identity on the finest level; use_mass = 0 treats it as the
 * identity without reading it), entirely on the device: d_x is updated in place, d_rq2[0] = x^T A x, d_rq2[1] = x^T M x
 * of the result (rho = their ratio).  d_work: 7 vectors of the level's size (4 when use_mass = 0) + 32 + 8 * 592 doubles.
 * No host synchronisation; the 2 x 2 generalised eigenproblem of every step (scipy.linalg.eig in the reference, :48) is
 * solved in closed form on the device. */
int mgcmt_rqmin(mgcmt_hier_t *h, int level, int use_mass, double *d_x, int nu, double *d_work, long long work_doubles,
                double *d_rq2, void *stream);
long long mgcmt_rqmin_work_doubles(mgcmt_hier_t *h, int level, int use_mass);
/* out = a*x + b*y with host scalars (the vector updates of rqmin, MGCMTSolver.py:34-36,51,54) */
int mgcmt_axpby(long long n, double a, const double *d_x, double b, const double *d_y, double *d_out,
                void *stream);
/* Gram-Schmidt of k vectors of length n stored one after another (vector-major: d_V + c*n is
 * column c).  modified != 0: MGS exactly as MGCMTProcessor.gramschmidt (MGCMTProcessor.py:44-50);
 * modified == 0: classical GS followed by normalisation (MGCMTProcessor.py:35-42). In place.
 * modified == 2: Gram-matrix (Cholesky-QR) form -- Q = W R^-1 with W^T W = R^T R, the same Q in exact arithmetic
 *   (QR with positive diagonal is unique) in 3k instead of ~k^2+3k vector passes; rounding differs from MGS by
 *   O(cond(W)^2 eps), so it is meant for the nearly orthonormal blocks of the eigen-iteration (k <= 6). */
int mgcmt_gramschmidt(long long n, int k, double *d_V, int modified, void *stream);

/* ---- banded operators: general complex128 matrices kept by diagonals (1-D multiband Hamiltonians) -------------
 * ThesisProblem.py:26-104 hands MGCMTSolver.vcycle the complex sparse matrix PotWellSolver.makeMatrix builds
 * (PotWellSolver.py:54-233): 4 (or 6) coupled bands, every block tridiagonal -- a dozen diagonals.  It is treated as
 * a 1-D problem of n = bands * gridpoints unknowns, exactly as the reference does (transfer operators included).
 * Vectors and diagonals are complex interleaved (re, im): n complex numbers = 2n doubles, 16-byte aligned.
 * Storage: d_vals[(k*n + i)] (complex) = A[i, i + h_offsets[k]], zero where the column falls outside the matrix;
 * offsets strictly ascending and containing 0.  Coarse operators are the Galerkin products R A P
 * (MGCMTSolver.py:318), formed on the device; n must be lowest_level * 2^L, lowest_level <= 512.
 * Smoothers: MGCMT_SMOOTH_WJACOBI, MGCMT_SMOOTH_GSLEX (omega == 1: gseidel; otherwise sor with quirk Q6). */
typedef struct mgcmt_band mgcmt_band_t;
int mgcmt_band_create(int n, int ndiag, const int *h_offsets, const double *d_vals, int lowest_level, void *stream,
                      mgcmt_band_t **out);
int mgcmt_band_destroy(mgcmt_band_t *h);
int mgcmt_band_num_levels(const mgcmt_band_t *h, int *out);
int mgcmt_band_level_shape(const mgcmt_band_t *h, int level, int *n, int *ndiag);
/* host copies of a level's offsets (ndiag ints) and diagonals (2*ndiag*n doubles); either may be NULL */
int mgcmt_band_level_diags(const mgcmt_band_t *h, int level, int *h_offsets, double *h_vals);
/* y = (A_l - shift I) x                                   -- `shifted_matrix * v`, MGCMTSolver.py:315 */
int mgcmt_band_apply(mgcmt_band_t *h, int level, double shift, const double *d_x, double *d_y, void *stream);
/* nu sweeps of wjacobi / gseidel / sor on (A_l - shift I) v = f, in place   -- MGCMTSolver.py:182-246 */
int mgcmt_band_smooth(mgcmt_band_t *h, int level, int smoother, int nu, double shift, double omega, double *d_v,
                      const double *d_f, void *stream);
/* rc = R (f - (A_l - shift I) v)                           -- MGCMTSolver.py:315 */
int mgcmt_band_residual_restrict(mgcmt_band_t *h, int level, double shift, const double *d_v, const double *d_f,
                                 double *d_rc, void *stream);
/* v += P ec                                                -- MGCMTSolver.py:323-324 */
int mgcmt_band_prolong_correct(mgcmt_band_t *h, int level, const double *d_ec, double *d_v, void *stream);
/* v = (A_coarsest - shift I)^-1 f                          -- MGCMTSolver.py:305-308 */
int mgcmt_band_coarse_solve(mgcmt_band_t *h, double shift, const double *d_f, double *d_v, void *stream);
/* one V-cycle on (A_0 - shift I) v = f, v holds the start vector on entry  -- MGCMTSolver.py:281-329 */
int mgcmt_band_vcycle(mgcmt_band_t *h, double shift, int nu1, int nu2, int smoother, double omega, double *d_v,
                      const double *d_f, void *stream);

/* ---- slab block: the multi-GPU step driven natively ------------------------------------------------------------
 * No counterpart in the reference (it has no parallelism).  multigridcmt_b200/slab.py defines the row-slab
 * decomposition and drives it from Python over torch.distributed; these entry points issue
 * the same kernels and exchanges from C++ with NCCL called directly, because at 8 GPUs the Python issue time of a step
 * is as long as the step.  NCCL is looked up at run time in the library the process already loaded (PyTorch's):
 * mgcmt_nccl_load(path or NULL); rank 0 makes an id (128 bytes) that the caller broadcasts; every rank creates its
 * communicator from it.  One rank per GPU; all ranks must make the same calls in the same order. */
typedef struct mgcmt_slabblock mgcmt_slabblock_t;
int mgcmt_nccl_load(const char *path);
int mgcmt_nccl_unique_id(void *out128);
int mgcmt_nccl_comm_create(const void *id128, int world, int rank, void **out_comm);
int mgcmt_nccl_comm_destroy(void *comm);
/* k vectors of an n x n grid, rows [rank n/world, (rank+1) n/world) + 10 halo rows on both sides on this rank;
 * levels 0..nlev_slab-1 stay decomposed, coarser ones are replicated after an all-gather.  Coefficient arrays as in
 * mgcmt_hier_create (host, length n).  comm may be NULL when world == 1.  comm2: a second communicator over the same
 * ranks (or NULL): with it the vectors run as two halves half a phase apart, one half's halo exchange overlapping the
 * other half's kernels. */
int mgcmt_slabblock_create(void *comm, void *comm2, int world, int rank, int n, int nlev_slab, int lowest_level, int k,
                           const double *h_row_lo, const double *h_row_di, const double *h_row_up,
                           const double *h_col_lo, const double *h_col_di, const double *h_col_up, double omega,
                           void *stream, mgcmt_slabblock_t **out);
int mgcmt_slabblock_destroy(mgcmt_slabblock_t *b);
/* k V(4,4) weighted-Jacobi cycles in lock-step, zero start: (H - h_shifts[c]) w = f_c; h_f0[c] / h_v0[c] are device
 * pointers to finest-level slab arrays ((n/world + 20) x n doubles; owned rows of f in, owned rows of w out, halo rows
 * are scratch).  d_lam != NULL: d_lam[2c], d_lam[2c+1] = w_c^T H w_c, w_c^T w_c, summed over all ranks.
 * (the drivers' loop body, 2DPotGS.py:93-103, for all k vectors) */
int mgcmt_slabblock_cycle(mgcmt_slabblock_t *b, const double *h_shifts, double *const *h_f0, double *const *h_v0,
                          double *d_lam, void *stream);
/* instrumentation: with profiling on, the stages of every following cycle (lock-step form, comm2 == NULL) are bracketed
 * by CUDA events on the block's ordering stream; _read returns, for the last cycle, the milliseconds of each
 * communication stage and of the compute stage that follows it (stage order: down levels 0.., replicated coarse part,
 * up levels ..0, Rayleigh sums if with_lam); arrays of at least 4*nlev_slab + 2 doubles */
int mgcmt_slabblock_profile(mgcmt_slabblock_t *b, int on);
/* smoother of the block's cycles: MGCMT_SMOOTH_WJACOBI (default, omega 2/3) or MGCMT_SMOOTH_RBGS (red-black / four-colour
 * Gauss-Seidel, the smoother of BASELINE config 3; on 9-point slab levels a leg is two passes of two sweeps with a halo
 * exchange of the intermediate iterate in between: the profile then has 4*nlev_slab stages) */
int mgcmt_slabblock_set_smoother(mgcmt_slabblock_t *b, int smoother, double omega);
int mgcmt_slabblock_profile_read(mgcmt_slabblock_t *b, int with_lam, double *h_ms_comm, double *h_ms_comp, int *nstage);
/* orthonormalise the k slab vectors d_block + c*stride (Gram-matrix form of MGCMTProcessor.gramschmidt: local packed
 * Gram matrix of the owned rows, one all-reduce, Q = W R^-1 locally) */
int mgcmt_slabblock_gram(mgcmt_slabblock_t *b, double *d_block, long long stride, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MGCMT_B200_H */
