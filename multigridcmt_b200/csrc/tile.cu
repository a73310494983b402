// tile.cu -- shared-memory kernels for the levels that are too small to feed the streaming kernels.
//
//  * tile_leg_kernel: one V-cycle leg (nu Jacobi sweeps fused with restriction of the residual, or with
//    prolongation + correction) on a T x T tile per CTA, overlapped tiling: the CTA stages the tile plus
//    a 6-point halo of v and f in shared memory, sweeps there (a barrier per sweep) and writes only
//    its own T x T (and (T/2)^2 coarse) points.  Used for mid-size levels, where a whole leg is a few
//    microseconds of work and what matters is having thousands of threads in flight, not HBM traffic.
//  * tail_kernel: the whole bottom of the V-cycle -- every level from `first` down to the coarsest,
//    including the exact coarsest solve, and back up -- in ONE launch by ONE CTA, all level vectors
//    resident in shared memory (north_star: "coarse levels ... collapsed into a single persistent-CTA
//    kernel").  Un-collapsed, these levels cost ~11 launches each at launch latency.
//
// Arithmetic: identical operators to stencil.cu / transfer.cu / coarse.cu (weighted Jacobi
// MGCMTSolver.py:182-208, residual + full weighting :315, interpolation + correction :323-324, exact
// coarsest solve :305-308 through the cached dense inverse).
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

namespace {

constexpr int kTileHalo = 6;  // >= nu + 2 for nu <= 4, even

struct RowC {
  double ka_lo, ka_di, ka_up, inside;  // inside: 1.0 for rows of the grid, else 0 (operator row zeroed)
  double ma_lo, ma_di, ma_up, pad;
};

// (A_s x)(r,c) from a shared-memory array with pitch P; coefficient rows rc (this row), column factors in
// registers.  FIVE: Ma = Mb = I.
template <bool FIVE>
__device__ __forceinline__ double apply_point(const double *__restrict__ a, int P, int r, int c, const RowC &rc,
                                              double kbl, double kbd, double kbu, double mbl, double mbd,
                                              double mbu, double shift) {
  const double *p = a + r * P + c;
  if (FIVE) {
    const double t = kbl * p[-1] + kbd * p[0] + kbu * p[1];
    return rc.ka_lo * p[-P] + (t + rc.ka_di * p[0]) + rc.ka_up * p[P] - shift * p[0];
  }
  const double tm = kbl * p[-P - 1] + kbd * p[-P] + kbu * p[-P + 1];
  const double t0 = kbl * p[-1] + kbd * p[0] + kbu * p[1];
  const double tp = kbl * p[P - 1] + kbd * p[P] + kbu * p[P + 1];
  const double sm = mbl * p[-P - 1] + mbd * p[-P] + mbu * p[-P + 1];
  const double s0 = mbl * p[-1] + mbd * p[0] + mbu * p[1];
  const double sp = mbl * p[P - 1] + mbd * p[P] + mbu * p[P + 1];
  return (rc.ma_lo * tm + rc.ka_lo * sm) + (rc.ma_di * t0 + rc.ka_di * s0) + (rc.ma_up * tp + rc.ka_up * sp) -
         shift * p[0];
}

}  // namespace

// mode: FUSED_SMOOTH / FUSED_DOWN / FUSED_DOWN_ZERO / FUSED_UP (kernels.h)
template <bool FIVE, int T>
__global__ void __launch_bounds__((T + 2 * kTileHalo) * (T == 32 ? 5 : 9))
tile_leg_kernel(LevelDev L, int mode, int nu, double shift, double omega, const double *__restrict__ v_in,
                const double *__restrict__ f, double *__restrict__ v_out, const double *__restrict__ e_coarse,
                double *__restrict__ r_coarse) {
  constexpr int H = kTileHalo;
  constexpr int E = T + 2 * H;   // extended tile edge
  constexpr int P = E + 1;       // pitch
  constexpr int RPP = (T == 32 ? 5 : 9);  // tile rows handled per pass (threads = E * RPP)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *A = reinterpret_cast<double *>(smem_raw);
  double *B = A + E * P;
  double *F = B + E * P;
  RowC *rows = reinterpret_cast<RowC *>(F + E * P);

  const int tid = threadIdx.x;
  const int c = tid % E;        // this thread's tile column (fixed)
  const int rbase = tid / E;    // first tile row
  const int i0 = blockIdx.y * T - H, j0 = blockIdx.x * T - H;  // grid coords of tile (0,0)
  const int gj = j0 + c;
  const bool cin = (gj >= 0 && gj < L.ncols);
  const int gjc = min(max(gj, 0), L.ncols - 1);
  const int ncc = L.ncols / 2, nrc = L.nrows / 2;

  // row table (zero operator rows outside the grid: Dirichlet zeros reproduce themselves)
  for (int r = tid; r < E; r += blockDim.x) {
    const int gi = i0 + r;
    const bool rin = (gi >= 0 && gi < L.nrows);
    const int g = L.row0 + min(max(gi, 0), L.nrows - 1);
    RowC rc;
    rc.ka_lo = rin ? L.ka_lo[g] : 0.0; rc.ka_di = rin ? L.ka_di[g] : 0.0; rc.ka_up = rin ? L.ka_up[g] : 0.0;
    if (FIVE) { rc.ma_lo = 0.0; rc.ma_di = rin ? 1.0 : 0.0; rc.ma_up = 0.0; }
    else { rc.ma_lo = rin ? L.ma_lo[g] : 0.0; rc.ma_di = rin ? L.ma_di[g] : 0.0; rc.ma_up = rin ? L.ma_up[g] : 0.0; }
    rc.inside = rin ? 1.0 : 0.0;
    rc.pad = 0.0;
    rows[r] = rc;
  }
  // column factors of this thread (zero outside the grid)
  const double kbl = cin ? L.kb_lo[gjc] : 0.0, kbd = cin ? L.kb_di[gjc] : 0.0, kbu = cin ? L.kb_up[gjc] : 0.0;
  const double mbl = (!FIVE && cin) ? L.mb_lo[gjc] : 0.0, mbd = FIVE ? 1.0 : (cin ? L.mb_di[gjc] : 0.0),
               mbu = (!FIVE && cin) ? L.mb_up[gjc] : 0.0;
  const double kbd_g = L.kb_di[gjc], mbd_g = FIVE ? 1.0 : L.mb_di[gjc];
  // omega / diag for the tile's reference row (its middle); other rows (first / last grid rows of a
  // Galerkin level) take the division
  const int gref = L.row0 + min(max(i0 + E / 2, 0), L.nrows - 1);
  const double kad_ref = L.ka_di[gref], mad_ref = FIVE ? 1.0 : L.ma_di[gref];
  const double wref = cin ? omega / ((mad_ref * kbd_g + kad_ref * mbd_g) - shift) : 0.0;

  // ---- stage the tile: A = v (+ P e), F = f ---------------------------------------------------------
  for (int r = rbase; r < E; r += RPP) {
    const int gi = i0 + r;
    const bool in = cin && gi >= 0 && gi < L.nrows;
    double x = 0.0, ff = 0.0;
    if (in) {
      ff = f[(size_t)gi * L.ncols + gj];
      if (mode != FUSED_DOWN_ZERO) x = v_in[(size_t)gi * L.ncols + gj];
      if (mode == FUSED_UP) {
        // (P e)(gi,gj): odd index -> injection of coarse (i-1)/2; even -> mean of coarse i/2-1 and i/2
        const int I = gi >> 1, J = gj >> 1;
        auto E_ = [&](int ii, int jj) -> double {
          return (ii >= 0 && ii < nrc && jj >= 0 && jj < ncc) ? e_coarse[(size_t)ii * ncc + jj] : 0.0;
        };
        double pe;
        if (gi & 1) {
          pe = (gj & 1) ? E_(I, J) : 0.5 * (E_(I, J - 1) + E_(I, J));
        } else {
          const double top = (gj & 1) ? E_(I - 1, J) : 0.5 * (E_(I - 1, J - 1) + E_(I - 1, J));
          const double bot = (gj & 1) ? E_(I, J) : 0.5 * (E_(I, J - 1) + E_(I, J));
          pe = 0.5 * (top + bot);
        }
        x += pe;
      }
    }
    A[r * P + c] = x;
    F[r * P + c] = ff;
  }
  __syncthreads();

  // ---- nu weighted-Jacobi sweeps, ping-pong A <-> B.  Border ring of the extended tile is never
  // updated: it is halo whose staleness moves inwards one point per sweep (H >= nu + 2 covers it).
  const bool cint = (c >= 1 && c <= E - 2);
  for (int s = 0; s < nu; ++s) {
    for (int r = rbase; r < E; r += RPP) {
      double o = A[r * P + c];
      if (cint && r >= 1 && r <= E - 2) {
        const RowC rc = rows[r];
        const double av = apply_point<FIVE>(A, P, r, c, rc, kbl, kbd, kbu, mbl, mbd, mbu, shift);
        double w = wref;
        if (rc.ka_di != kad_ref || rc.ma_di != mad_ref)
          w = (cin && rc.inside != 0.0) ? omega / ((rc.ma_di * kbd_g + rc.ka_di * mbd_g) - shift) : 0.0;
        o = o + w * (F[r * P + c] - av);
      }
      B[r * P + c] = o;
    }
    __syncthreads();
    double *t = A; A = B; B = t;
  }

  // ---- write the tile's own points ------------------------------------------------------------------
  if (!((mode == FUSED_DOWN || mode == FUSED_DOWN_ZERO) && nu == 0)) {
    if (c >= H && c < H + T && cin) {
      for (int r = rbase; r < E; r += RPP) {
        const int gi = i0 + r;
        if (r >= H && r < H + T && gi < L.nrows) v_out[(size_t)gi * L.ncols + gj] = A[r * P + c];
      }
    }
  }
  if (mode != FUSED_DOWN && mode != FUSED_DOWN_ZERO) return;

  // ---- residual into B, then full weighting ---------------------------------------------------------
  for (int r = rbase; r < E; r += RPP) {
    double res = 0.0;
    if (cint && r >= 1 && r <= E - 2) {
      const RowC rc = rows[r];
      if (cin && rc.inside != 0.0)
        res = F[r * P + c] - apply_point<FIVE>(A, P, r, c, rc, kbl, kbd, kbu, mbl, mbd, mbu, shift);
    }
    B[r * P + c] = res;
  }
  __syncthreads();
  for (int idx = tid; idx < (T / 2) * (T / 2); idx += blockDim.x) {
    const int ci = idx / (T / 2), cj = idx - ci * (T / 2);
    const int I = (blockIdx.y * T) / 2 + ci, J = (blockIdx.x * T) / 2 + cj;
    if (I < nrc && J < ncc) {
      const double *p = B + (H + 2 * ci) * P + (H + 2 * cj);
      const double a0 = 0.25 * p[0] + 0.5 * p[1] + 0.25 * p[2];
      const double a1 = 0.25 * p[P] + 0.5 * p[P + 1] + 0.25 * p[P + 2];
      const double a2 = 0.25 * p[2 * P] + 0.5 * p[2 * P + 1] + 0.25 * p[2 * P + 2];
      r_coarse[(size_t)I * ncc + J] = 0.25 * a0 + 0.5 * a1 + 0.25 * a2;
    }
  }
}

template <bool FIVE, int T>
static cudaError_t launch_tile_t(const LevelDev &L, int mode, int nu, double shift, double omega,
                                 const double *v_in, const double *f, double *v_out, const double *e_coarse,
                                 double *r_coarse, cudaStream_t s) {
  constexpr int E = T + 2 * kTileHalo;
  constexpr int threads = E * (T == 32 ? 5 : 9);
  const size_t smem = sizeof(double) * 3 * E * (E + 1) + sizeof(RowC) * E;
  auto kern = tile_leg_kernel<FIVE, T>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((L.ncols + T - 1) / T, (L.nrows + T - 1) / T);
  kern<<<grid, threads, smem, s>>>(L, mode, nu, shift, omega, v_in, f, v_out, e_coarse, r_coarse);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_tile_leg(const LevelDev &L, int mode, int nu, double shift, double omega, const double *v_in,
                            const double *f, double *v_out, const double *e_coarse, double *r_coarse,
                            cudaStream_t s) {
  if (nu < 0 || nu > 4 || L.nrows < 2) return cudaErrorInvalidValue;
  const bool small = (long long)L.nrows * L.ncols <= 256LL * 256;  // 16x16 tiles keep >= 64 CTAs busy
  if (L.five)
    return small ? launch_tile_t<true, 16>(L, mode, nu, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s)
                 : launch_tile_t<true, 32>(L, mode, nu, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  return small ? launch_tile_t<false, 16>(L, mode, nu, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s)
               : launch_tile_t<false, 32>(L, mode, nu, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
}

// -----------------------------------------------------------------------------------------------------
// Gauss-Seidel / SOR legs on a tile: `sweeps` colour sweeps (red-black on a 5-point level, the four colours
// (0,0) (1,1) (0,1) (1,0) of gs.cu on 9-point levels) IN PLACE in shared memory, a barrier per colour stage -- points of
// one colour only have neighbours of the other colours, so a stage has no read-write conflicts.  The halo covers all
// stages of a whole leg (4 sweeps x 4 colours + residual = 17 points), so a leg is ONE launch; the streaming
// colour-stage legs need two passes per leg on 9-point levels and ~20 us per launch on grids this small (one warp
// walking down its strip), which made the 512^2 .. 128^2 levels a quarter of an RB-GS cycle at 4096^2.
// -----------------------------------------------------------------------------------------------------
constexpr int kTileHaloGS = 18;  // >= 4 * sweeps + 2 for sweeps <= 4, even

template <bool FIVE, int T>
__global__ void __launch_bounds__((T + 2 * kTileHaloGS) * 9)
tile_gs_leg_kernel(LevelDev L, int mode, int sweeps, double shift, double omega, const double *__restrict__ v_in,
                   const double *__restrict__ f, double *__restrict__ v_out, const double *__restrict__ e_coarse,
                   double *__restrict__ r_coarse) {
  constexpr int H = kTileHaloGS;
  constexpr int E = T + 2 * H;
  constexpr int P = E + 1;
  constexpr int RPP = 9;
  constexpr int NCOL = FIVE ? 2 : 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *A = reinterpret_cast<double *>(smem_raw);
  double *F = A + E * P;
  RowC *rows = reinterpret_cast<RowC *>(F + E * P);

  const int tid = threadIdx.x;
  const int c = tid % E;
  const int rbase = tid / E;
  const int i0 = blockIdx.y * T - H, j0 = blockIdx.x * T - H;
  const int gj = j0 + c;
  const bool cin = (gj >= 0 && gj < L.ncols);
  const int gjc = min(max(gj, 0), L.ncols - 1);
  const int ncc = L.ncols / 2, nrc = L.nrows / 2;

  for (int r = tid; r < E; r += blockDim.x) {
    const int gi = i0 + r;
    const bool rin = (gi >= 0 && gi < L.nrows);
    const int g = L.row0 + min(max(gi, 0), L.nrows - 1);
    RowC rc;
    rc.ka_lo = rin ? L.ka_lo[g] : 0.0; rc.ka_di = rin ? L.ka_di[g] : 0.0; rc.ka_up = rin ? L.ka_up[g] : 0.0;
    if (FIVE) { rc.ma_lo = 0.0; rc.ma_di = rin ? 1.0 : 0.0; rc.ma_up = 0.0; }
    else { rc.ma_lo = rin ? L.ma_lo[g] : 0.0; rc.ma_di = rin ? L.ma_di[g] : 0.0; rc.ma_up = rin ? L.ma_up[g] : 0.0; }
    rc.inside = rin ? 1.0 : 0.0;
    rc.pad = 0.0;
    rows[r] = rc;
  }
  const double kbl = cin ? L.kb_lo[gjc] : 0.0, kbd = cin ? L.kb_di[gjc] : 0.0, kbu = cin ? L.kb_up[gjc] : 0.0;
  const double mbl = (!FIVE && cin) ? L.mb_lo[gjc] : 0.0, mbd = FIVE ? 1.0 : (cin ? L.mb_di[gjc] : 0.0),
               mbu = (!FIVE && cin) ? L.mb_up[gjc] : 0.0;
  const double kbd_g = L.kb_di[gjc], mbd_g = FIVE ? 1.0 : L.mb_di[gjc];
  const int gref = L.row0 + min(max(i0 + E / 2, 0), L.nrows - 1);
  const double kad_ref = L.ka_di[gref], mad_ref = FIVE ? 1.0 : L.ma_di[gref];
  const double wref = cin ? omega / ((mad_ref * kbd_g + kad_ref * mbd_g) - shift) : 0.0;

  // stage the tile: A = v (+ P e), F = f
  for (int r = rbase; r < E; r += RPP) {
    const int gi = i0 + r;
    const bool in = cin && gi >= 0 && gi < L.nrows;
    double x = 0.0, ff = 0.0;
    if (in) {
      ff = f[(size_t)gi * L.ncols + gj];
      if (mode != FUSED_DOWN_ZERO) x = v_in[(size_t)gi * L.ncols + gj];
      if (mode == FUSED_UP) {
        const int I = gi >> 1, J = gj >> 1;
        auto E_ = [&](int ii, int jj) -> double {
          return (ii >= 0 && ii < nrc && jj >= 0 && jj < ncc) ? e_coarse[(size_t)ii * ncc + jj] : 0.0;
        };
        double pe;
        if (gi & 1) {
          pe = (gj & 1) ? E_(I, J) : 0.5 * (E_(I, J - 1) + E_(I, J));
        } else {
          const double top = (gj & 1) ? E_(I - 1, J) : 0.5 * (E_(I - 1, J - 1) + E_(I - 1, J));
          const double bot = (gj & 1) ? E_(I, J) : 0.5 * (E_(I, J - 1) + E_(I, J));
          pe = 0.5 * (top + bot);
        }
        x += pe;
      }
    }
    A[r * P + c] = x;
    F[r * P + c] = ff;
  }
  __syncthreads();

  // colour stages, in place; the border ring of the extended tile is never updated (halo)
  const bool cint = (c >= 1 && c <= E - 2);
  const int pc = gj & 1;
  for (int st = 0; st < sweeps * NCOL; ++st) {
    const int colour = st % NCOL;
    for (int r = rbase; r < E; r += RPP) {
      const int pr = (i0 + r + L.row0) & 1;
      bool mine;
      if (FIVE) mine = ((pr + pc) & 1) == colour;
      else mine = colour == 0 ? (pr == 0 && pc == 0) : colour == 1 ? (pr == 1 && pc == 1) : colour == 2 ? (pr == 0 && pc == 1) : (pr == 1 && pc == 0);
      if (mine && cint && r >= 1 && r <= E - 2) {
        const RowC rc = rows[r];
        const double av = apply_point<FIVE>(A, P, r, c, rc, kbl, kbd, kbu, mbl, mbd, mbu, shift);
        double w = wref;
        if (rc.ka_di != kad_ref || rc.ma_di != mad_ref)
          w = (cin && rc.inside != 0.0) ? omega / ((rc.ma_di * kbd_g + rc.ka_di * mbd_g) - shift) : 0.0;
        A[r * P + c] += w * (F[r * P + c] - av);
      }
    }
    __syncthreads();
  }

  if (!((mode == FUSED_DOWN || mode == FUSED_DOWN_ZERO) && sweeps == 0)) {
    if (c >= H && c < H + T && cin) {
      for (int r = rbase; r < E; r += RPP) {
        const int gi = i0 + r;
        if (r >= H && r < H + T && gi < L.nrows) v_out[(size_t)gi * L.ncols + gj] = A[r * P + c];
      }
    }
  }
  if (mode != FUSED_DOWN && mode != FUSED_DOWN_ZERO) return;

  // residual in place into F (a point's residual needs only its own f), then full weighting
  for (int r = rbase; r < E; r += RPP) {
    double res = 0.0;
    if (cint && r >= 1 && r <= E - 2) {
      const RowC rc = rows[r];
      if (cin && rc.inside != 0.0)
        res = F[r * P + c] - apply_point<FIVE>(A, P, r, c, rc, kbl, kbd, kbu, mbl, mbd, mbu, shift);
    }
    F[r * P + c] = res;
  }
  __syncthreads();
  for (int idx = tid; idx < (T / 2) * (T / 2); idx += blockDim.x) {
    const int ci = idx / (T / 2), cj = idx - ci * (T / 2);
    const int I = (blockIdx.y * T) / 2 + ci, J = (blockIdx.x * T) / 2 + cj;
    if (I < nrc && J < ncc) {
      const double *p = F + (H + 2 * ci) * P + (H + 2 * cj);
      const double a0 = 0.25 * p[0] + 0.5 * p[1] + 0.25 * p[2];
      const double a1 = 0.25 * p[P] + 0.5 * p[P + 1] + 0.25 * p[P + 2];
      const double a2 = 0.25 * p[2 * P] + 0.5 * p[2 * P + 1] + 0.25 * p[2 * P + 2];
      r_coarse[(size_t)I * ncc + J] = 0.25 * a0 + 0.5 * a1 + 0.25 * a2;
    }
  }
}

template <bool FIVE, int T>
static cudaError_t launch_tile_gs_t(const LevelDev &L, int mode, int sweeps, double shift, double omega, const double *v_in,
                                    const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s) {
  constexpr int E = T + 2 * kTileHaloGS;
  constexpr int threads = E * 9;
  const size_t smem = sizeof(double) * 2 * E * (E + 1) + sizeof(RowC) * E;
  auto kern = tile_gs_leg_kernel<FIVE, T>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((L.ncols + T - 1) / T, (L.nrows + T - 1) / T);
  kern<<<grid, threads, smem, s>>>(L, mode, sweeps, shift, omega, v_in, f, v_out, e_coarse, r_coarse);
  count_launch();
  return cudaGetLastError();
}

// sweeps: 0..4 colour sweeps in one launch
cudaError_t launch_tile_gs_leg(const LevelDev &L, int mode, int sweeps, double shift, double omega, const double *v_in,
                               const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s) {
  if (sweeps < 0 || sweeps > 4 || L.nrows < 2) return cudaErrorInvalidValue;
  const bool small = (long long)L.nrows * L.ncols <= 256LL * 256;  // 32 x 32 tiles keep >= 64 CTAs busy
  if (L.five)
    return small ? launch_tile_gs_t<true, 32>(L, mode, sweeps, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s)
                 : launch_tile_gs_t<true, 64>(L, mode, sweeps, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  return small ? launch_tile_gs_t<false, 32>(L, mode, sweeps, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s)
               : launch_tile_gs_t<false, 64>(L, mode, sweeps, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
}

// =====================================================================================================
// tail kernel: levels first .. last (last = coarsest) of one hierarchy in one CTA.
// Shared-memory arena per level: V and TMP with a one-point zero halo (pitch n+2: no bounds checks, and
// the Dirichlet zeros / the truncated last restriction row come for free), F, W = omega/diag (the
// divisions are done once per launch, not once per sweep), and the 12 tridiagonal factor arrays.
// =====================================================================================================
struct TailLevel {
  LevelDev dev;
  int voff, toff, foff, woff, coff;  // offsets (doubles) into the arena: V, TMP (haloed), F, W, coefficients
};
struct TailArgs {
  int nlev;                 // number of levels handled (>= 2: at least one smoothing level + the coarsest)
  TailLevel lev[kTailMaxLevels];
  int arena_doubles;
  const double *inv;        // dense inverse of the coarsest shifted operator (n_c x n_c, row-major)
  double shift, omega;
  int coarsen_rows;         // 0: 1-D hierarchy (single-row levels, transfers act along the row only)
};

namespace {

// coefficient block of a level: [ka_lo ka_di ka_up ma_lo ma_di ma_up] x nrows, then the same for columns
struct TailCoef {
  const double *ka_lo, *ka_di, *ka_up, *ma_lo, *ma_di, *ma_up, *kb_lo, *kb_di, *kb_up, *mb_lo, *mb_di, *mb_up;
};
__device__ __forceinline__ TailCoef tail_coef(double *base, int nr, int nc) {
  TailCoef c;
  c.ka_lo = base; c.ka_di = base + nr; c.ka_up = base + 2 * nr;
  c.ma_lo = base + 3 * nr; c.ma_di = base + 4 * nr; c.ma_up = base + 5 * nr;
  double *b = base + 6 * nr;
  c.kb_lo = b; c.kb_di = b + nc; c.kb_up = b + 2 * nc;
  c.mb_lo = b + 3 * nc; c.mb_di = b + 4 * nc; c.mb_up = b + 5 * nc;
  return c;
}

// (A_s x)(i,j) on a haloed array (pitch P = nc + 2), general 9-point separable form (Ma = Mb = I on a
// 5-point level: the extra terms are exact zeros)
__device__ __forceinline__ double tail_apply(const double *x, int P, int i, int j, const TailCoef &c, double shift) {
  const double *p = x + (i + 1) * P + (j + 1);
  const double kbl = c.kb_lo[j], kbd = c.kb_di[j], kbu = c.kb_up[j];
  const double mbl = c.mb_lo[j], mbd = c.mb_di[j], mbu = c.mb_up[j];
  const double tm = kbl * p[-P - 1] + kbd * p[-P] + kbu * p[-P + 1], sm = mbl * p[-P - 1] + mbd * p[-P] + mbu * p[-P + 1];
  const double t0 = kbl * p[-1] + kbd * p[0] + kbu * p[1], s0 = mbl * p[-1] + mbd * p[0] + mbu * p[1];
  const double tp = kbl * p[P - 1] + kbd * p[P] + kbu * p[P + 1], sp = mbl * p[P - 1] + mbd * p[P] + mbu * p[P + 1];
  return (c.ma_lo[i] * tm + c.ka_lo[i] * sm) + (c.ma_di[i] * t0 + c.ka_di[i] * s0) + (c.ma_up[i] * tp + c.ka_up[i] * sp) -
         shift * p[0];
}

// dst = src + W (F - A_s src)  (one weighted-Jacobi sweep), or dst = F - A_s src when residual
template <bool RESIDUAL>
__device__ __forceinline__ void tail_pass(int nr, int nc, int lgc, const TailCoef &c, double shift, const double *src,
                                          const double *F, const double *W, double *dst) {
  const int P = nc + 2;
  for (int idx = threadIdx.x; idx < nr * nc; idx += blockDim.x) {
    const int i = idx >> lgc, j = idx & (nc - 1);
    const double av = tail_apply(src, P, i, j, c, shift);
    const double r = F[idx] - av;
    dst[(i + 1) * P + (j + 1)] = RESIDUAL ? r : (src[(i + 1) * P + (j + 1)] + W[idx] * r);
  }
}

}  // namespace

// f_first (global) -> v_first (global): the V-cycle restricted to levels first..last with a zero start
__global__ void __launch_bounds__(1024) tail_kernel(TailArgs a, const double *__restrict__ f_first,
                                                    double *__restrict__ v_first) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *arena = reinterpret_cast<double *>(smem_raw);
  const int nl = a.nlev;
  const int tid = threadIdx.x, nt = blockDim.x;
  // ---- setup: zero everything (halos, zero initial guesses), stage coefficients, F of the first level
  for (int i = tid; i < a.arena_doubles; i += nt) arena[i] = 0.0;
  __syncthreads();
  for (int l = 0; l < nl; ++l) {
    const LevelDev &L = a.lev[l].dev;
    double *cb = arena + a.lev[l].coff;
    const int nr = L.nrows, nc = L.ncols;
    for (int i = tid; i < nr; i += nt) {
      cb[i] = L.ka_lo[L.row0 + i]; cb[nr + i] = L.ka_di[L.row0 + i]; cb[2 * nr + i] = L.ka_up[L.row0 + i];
      cb[3 * nr + i] = L.five ? 0.0 : L.ma_lo[L.row0 + i];
      cb[4 * nr + i] = L.five ? 1.0 : L.ma_di[L.row0 + i];
      cb[5 * nr + i] = L.five ? 0.0 : L.ma_up[L.row0 + i];
    }
    double *cc = cb + 6 * nr;
    for (int j = tid; j < nc; j += nt) {
      cc[j] = L.kb_lo[j]; cc[nc + j] = L.kb_di[j]; cc[2 * nc + j] = L.kb_up[j];
      cc[3 * nc + j] = L.five ? 0.0 : L.mb_lo[j];
      cc[4 * nc + j] = L.five ? 1.0 : L.mb_di[j];
      cc[5 * nc + j] = L.five ? 0.0 : L.mb_up[j];
    }
  }
  {
    const int n0 = a.lev[0].dev.nrows * a.lev[0].dev.ncols;
    double *f0 = arena + a.lev[0].foff;
    for (int i = tid; i < n0; i += nt) f0[i] = f_first[i];
  }
  __syncthreads();
  for (int l = 0; l < nl - 1; ++l) {  // W = omega / diag, once
    const LevelDev &L = a.lev[l].dev;
    const int nr = L.nrows, nc = L.ncols, lgc = 31 - __clz(nc);
    const TailCoef c = tail_coef(arena + a.lev[l].coff, nr, nc);
    double *W = arena + a.lev[l].woff;
    for (int idx = tid; idx < nr * nc; idx += nt) {
      const int i = idx >> lgc, j = idx & (nc - 1);
      W[idx] = a.omega / ((c.ma_di[i] * c.kb_di[j] + c.ka_di[i] * c.mb_di[j]) - a.shift);
    }
  }
  __syncthreads();

  // ---- down: 4 sweeps (V -> T -> V -> T -> V), residual into T, full weighting into the next F --------
  for (int l = 0; l < nl - 1; ++l) {
    const LevelDev &L = a.lev[l].dev;
    const int nr = L.nrows, nc = L.ncols, lgc = 31 - __clz(nc), P = nc + 2;
    const TailCoef c = tail_coef(arena + a.lev[l].coff, nr, nc);
    double *V = arena + a.lev[l].voff, *T = arena + a.lev[l].toff;
    const double *F = arena + a.lev[l].foff, *W = arena + a.lev[l].woff;
    for (int s = 0; s < 2; ++s) {
      tail_pass<false>(nr, nc, lgc, c, a.shift, V, F, W, T);
      __syncthreads();
      tail_pass<false>(nr, nc, lgc, c, a.shift, T, F, W, V);
      __syncthreads();
    }
    tail_pass<true>(nr, nc, lgc, c, a.shift, V, F, W, T);
    __syncthreads();
    const int nrc = a.coarsen_rows ? nr >> 1 : nr, ncc = nc >> 1, lgcc = lgc - 1;
    double *Fc = arena + a.lev[l + 1].foff;
    for (int idx = tid; idx < nrc * ncc; idx += nt) {
      const int I = idx >> lgcc, J = idx & (ncc - 1);
      if (a.coarsen_rows) {
        const double *p = T + (2 * I + 1) * P + (2 * J + 1);  // haloed: fine (2I, 2J); row/col n are the zero halo
        const double a0 = 0.25 * p[0] + 0.5 * p[1] + 0.25 * p[2];
        const double a1 = 0.25 * p[P] + 0.5 * p[P + 1] + 0.25 * p[P + 2];
        const double a2 = 0.25 * p[2 * P] + 0.5 * p[2 * P + 1] + 0.25 * p[2 * P + 2];
        Fc[idx] = 0.25 * a0 + 0.5 * a1 + 0.25 * a2;
      } else {
        const double *p = T + (I + 1) * P + (2 * J + 1);      // 1-D: full weighting along the row only
        Fc[idx] = 0.25 * p[0] + 0.5 * p[1] + 0.25 * p[2];
      }
    }
    __syncthreads();
  }
  // ---- coarsest: v = inv f (one warp per row, same summation order as gemv_kernel in coarse.cu) --------
  {
    const LevelDev &L = a.lev[nl - 1].dev;
    const int n = L.nrows * L.ncols, nc = L.ncols, P = nc + 2, lgc = 31 - __clz(nc);
    const double *F = arena + a.lev[nl - 1].foff;
    double *V = arena + a.lev[nl - 1].voff;
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int r = warp; r < n; r += nw) {
      const double *row = a.inv + (size_t)r * n;
      double acc = 0.0;
      for (int cidx = lane; cidx < n; cidx += 32) acc += row[cidx] * F[cidx];
      acc = warp_sum(acc);
      if (lane == 0) V[((r >> lgc) + 1) * P + (r & (nc - 1)) + 1] = acc;
    }
  }
  __syncthreads();
  // ---- up: v += P e, then 4 sweeps ------------------------------------------------------------------------
  for (int l = nl - 2; l >= 0; --l) {
    const LevelDev &L = a.lev[l].dev;
    const int nr = L.nrows, nc = L.ncols, lgc = 31 - __clz(nc), P = nc + 2, Pc = (nc >> 1) + 2;
    const TailCoef c = tail_coef(arena + a.lev[l].coff, nr, nc);
    double *V = arena + a.lev[l].voff, *T = arena + a.lev[l].toff;
    const double *F = arena + a.lev[l].foff, *W = arena + a.lev[l].woff;
    const double *E = arena + a.lev[l + 1].voff;
    for (int idx = tid; idx < nr * nc; idx += nt) {
      const int i = idx >> lgc, j = idx & (nc - 1);
      const int I = a.coarsen_rows ? i >> 1 : i, J = j >> 1;
      const double *e = E + (I + 1) * Pc + (J + 1);  // e[-1], e[-Pc]: coarse neighbours (zero halo at the edges)
      double pe;
      if (!a.coarsen_rows) {
        pe = (j & 1) ? e[0] : 0.5 * (e[-1] + e[0]);   // 1-D: interpolation along the row only
      } else if (i & 1) {
        pe = (j & 1) ? e[0] : 0.5 * (e[-1] + e[0]);
      } else {
        const double top = (j & 1) ? e[-Pc] : 0.5 * (e[-Pc - 1] + e[-Pc]);
        const double bot = (j & 1) ? e[0] : 0.5 * (e[-1] + e[0]);
        pe = 0.5 * (top + bot);
      }
      V[(i + 1) * P + (j + 1)] += pe;
    }
    __syncthreads();
    for (int s = 0; s < 2; ++s) {
      tail_pass<false>(nr, nc, lgc, c, a.shift, V, F, W, T);
      __syncthreads();
      tail_pass<false>(nr, nc, lgc, c, a.shift, T, F, W, V);
      __syncthreads();
    }
  }
  {
    const LevelDev &L = a.lev[0].dev;
    const int nc = L.ncols, lgc = 31 - __clz(nc), P = nc + 2;
    const double *V = arena + a.lev[0].voff;
    for (int idx = tid; idx < L.nrows * nc; idx += nt) v_first[idx] = V[((idx >> lgc) + 1) * P + (idx & (nc - 1)) + 1];
  }
}

static size_t tail_layout(const LevelDev *levels, int nlev, TailArgs *a) {
  int off = 0;
  auto take = [&](int n) { const int o = off; off += (n + 1) & ~1; return o; };
  for (int l = 0; l < nlev; ++l) {
    const int nr = levels[l].nrows, nc = levels[l].ncols;
    const int hal = (nr + 2) * (nc + 2);
    const int vo = take(hal), to = (l + 1 < nlev) ? take(hal) : 0, fo = take(nr * nc);
    const int wo = (l + 1 < nlev) ? take(nr * nc) : 0, co = take(6 * nr + 6 * nc);
    if (a) {
      a->lev[l].dev = levels[l];
      a->lev[l].voff = vo; a->lev[l].toff = to; a->lev[l].foff = fo; a->lev[l].woff = wo; a->lev[l].coff = co;
    }
  }
  if (a) a->arena_doubles = off;
  return sizeof(double) * (size_t)off;
}

size_t tail_smem_bytes(const LevelDev *levels, int nlev) { return tail_layout(levels, nlev, nullptr); }

cudaError_t launch_tail(const LevelDev *levels, int nlev, bool coarsen_rows, const double *inv, double shift,
                        double omega, const double *f_first, double *v_first, cudaStream_t s) {
  if (nlev < 2 || nlev > kTailMaxLevels) return cudaErrorInvalidValue;
  TailArgs a;
  a.nlev = nlev;
  a.coarsen_rows = coarsen_rows ? 1 : 0;
  a.inv = inv;
  a.shift = shift;
  a.omega = omega;
  const size_t smem = tail_layout(levels, nlev, &a);
  if (smem > kTailMaxSmem) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailMaxSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int n0 = levels[0].nrows * levels[0].ncols;
  int threads = n0 >= 1024 ? 1024 : (n0 < 64 ? 64 : ((n0 + 31) / 32) * 32);
  tail_kernel<<<1, threads, smem, s>>>(a, f_first, v_first);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mgcmt
