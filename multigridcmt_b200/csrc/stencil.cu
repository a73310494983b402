// stencil.cu -- row-marching stencil kernels: weighted Jacobi, residual, operator apply.
//
// One thread owns two adjacent columns and marches down a chunk of rows with a three-row window in
// registers, so every grid value is read from HBM once per sweep (plus one halo row per chunk, which
// the neighbouring CTA has just pulled through L2).  Left/right neighbours come from warp shuffles;
// only the two edge lanes of a warp issue an extra (L1-resident) scalar load.
//
// Reference arithmetic being restated (paths relative to the reference root):
//   weighted Jacobi   MGCMTSolver.py:182-208   v <- (I - w D^-1 A) v + w D^-1 f
//   residual          MGCMTSolver.py:315       f - (A - shift I) v
// with A_l = Ma (x) Kb + Ka (x) Mb and the shift applied as -shift*I on every level
// (MGCMTSolver.py:287-288).
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

struct Row4 {
  double xl, x0, x1, xr;
};

// value pair of local row i (i may be -1 or nrows: halo row or Dirichlet zero), plus the two
// horizontal neighbours
__device__ __forceinline__ Row4 load_row(const LevelDev &L, const double *__restrict__ v,
                                         const double *__restrict__ halo_top,
                                         const double *__restrict__ halo_bot, int i, int j0, bool active,
                                         int lane) {
  const double *src;
  if (i < 0)
    src = halo_top;
  else if (i >= L.nrows)
    src = halo_bot;
  else
    src = v + (size_t)i * L.ncols;
  Row4 r;
  double2 x = make_double2(0.0, 0.0);
  const bool ok = (src != nullptr) && active;
  if (ok) x = *reinterpret_cast<const double2 *>(src + j0);
  r.x0 = x.x;
  r.x1 = x.y;
  r.xl = __shfl_up_sync(0xffffffffu, x.y, 1);
  r.xr = __shfl_down_sync(0xffffffffu, x.x, 1);
  if (lane == 0) r.xl = (ok && j0 > 0) ? src[j0 - 1] : 0.0;
  if (lane == 31) r.xr = (ok && j0 + 2 < L.ncols) ? src[j0 + 2] : 0.0;
  return r;
}

// row-direction ("horizontal") factor applied to a row: t = Kb row, s = Mb row at the two columns
struct HV {
  double t0, t1, s0, s1, x0, x1;
};

template <int OP, bool FIVE>
__global__ void __launch_bounds__(128)
stencil_march_kernel(LevelDev L, double shift, double omega, const double *__restrict__ v,
                     const double *__restrict__ f, double *__restrict__ out,
                     const double *__restrict__ halo_top, const double *__restrict__ halo_bot,
                     int rows_per_cta) {
  const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const bool active = j0 < L.ncols;
  const int lane = threadIdx.x & 31;
  const int i_begin = blockIdx.y * rows_per_cta;
  const int i_end = min(i_begin + rows_per_cta, L.nrows);
  const int jc = active ? j0 : 0;

  const double kbl0 = L.kb_lo[jc], kbd0 = L.kb_di[jc], kbu0 = L.kb_up[jc];
  const double kbl1 = L.kb_lo[jc + 1], kbd1 = L.kb_di[jc + 1], kbu1 = L.kb_up[jc + 1];
  double mbl0 = 0, mbd0 = 1, mbu0 = 0, mbl1 = 0, mbd1 = 1, mbu1 = 0;
  if (!FIVE) {
    mbl0 = L.mb_lo[jc]; mbd0 = L.mb_di[jc]; mbu0 = L.mb_up[jc];
    mbl1 = L.mb_lo[jc + 1]; mbd1 = L.mb_di[jc + 1]; mbu1 = L.mb_up[jc + 1];
  }

  auto horiz = [&](const Row4 &r) {
    HV h;
    h.x0 = r.x0;
    h.x1 = r.x1;
    h.t0 = kbl0 * r.xl + kbd0 * r.x0 + kbu0 * r.x1;
    h.t1 = kbl1 * r.x0 + kbd1 * r.x1 + kbu1 * r.xr;
    if (!FIVE) {
      h.s0 = mbl0 * r.xl + mbd0 * r.x0 + mbu0 * r.x1;
      h.s1 = mbl1 * r.x0 + mbd1 * r.x1 + mbu1 * r.xr;
    } else {
      h.s0 = r.x0;
      h.s1 = r.x1;
    }
    return h;
  };

  HV p, c, n;
  if (FIVE) {
    // 5-point: only the centre row needs its horizontal part
    Row4 rp = load_row(L, v, halo_top, halo_bot, i_begin - 1, j0, active, lane);
    p.x0 = rp.x0; p.x1 = rp.x1; p.s0 = rp.x0; p.s1 = rp.x1; p.t0 = p.t1 = 0;
  } else {
    p = horiz(load_row(L, v, halo_top, halo_bot, i_begin - 1, j0, active, lane));
  }
  c = horiz(load_row(L, v, halo_top, halo_bot, i_begin, j0, active, lane));

  double last_kad = 0, last_mad = 0, winv0 = 0, winv1 = 0;
  bool have_w = false;

  for (int i = i_begin; i < i_end; ++i) {
    Row4 rn = load_row(L, v, halo_top, halo_bot, i + 1, j0, active, lane);
    double2 ff = make_double2(0.0, 0.0);
    if (OP != OP_APPLY && active) ff = *reinterpret_cast<const double2 *>(f + (size_t)i * L.ncols + j0);
    if (FIVE) {
      n.x0 = rn.x0; n.x1 = rn.x1; n.s0 = rn.x0; n.s1 = rn.x1; n.t0 = n.t1 = 0;
      n.t0 = kbl0 * rn.xl + kbd0 * rn.x0 + kbu0 * rn.x1;  // becomes the centre row next iteration
      n.t1 = kbl1 * rn.x0 + kbd1 * rn.x1 + kbu1 * rn.xr;
    } else {
      n = horiz(rn);
    }
    const int gi = L.row0 + i;
    const double kal = L.ka_lo[gi], kad = L.ka_di[gi], kau = L.ka_up[gi];
    double av0, av1, mad = 1.0;
    if (FIVE) {
      av0 = kal * p.x0 + (c.t0 + kad * c.x0) + kau * n.x0 - shift * c.x0;
      av1 = kal * p.x1 + (c.t1 + kad * c.x1) + kau * n.x1 - shift * c.x1;
    } else {
      const double mal = L.ma_lo[gi], mau = L.ma_up[gi];
      mad = L.ma_di[gi];
      av0 = (mal * p.t0 + kal * p.s0) + (mad * c.t0 + kad * c.s0) + (mau * n.t0 + kau * n.s0) - shift * c.x0;
      av1 = (mal * p.t1 + kal * p.s1) + (mad * c.t1 + kad * c.s1) + (mau * n.t1 + kau * n.s1) - shift * c.x1;
    }
    double2 o;
    if (OP == OP_JACOBI) {
      if (!have_w || kad != last_kad || mad != last_mad) {  // block-uniform: first and last rows only
        const double d0 = FIVE ? (kad + kbd0) - shift : (mad * kbd0 + kad * mbd0) - shift;
        const double d1 = FIVE ? (kad + kbd1) - shift : (mad * kbd1 + kad * mbd1) - shift;
        winv0 = omega / d0;
        winv1 = omega / d1;
        last_kad = kad; last_mad = mad; have_w = true;
      }
      o.x = c.x0 + winv0 * (ff.x - av0);
      o.y = c.x1 + winv1 * (ff.y - av1);
    } else if (OP == OP_RESIDUAL) {
      o.x = ff.x - av0;
      o.y = ff.y - av1;
    } else {
      o.x = av0;
      o.y = av1;
    }
    if (active) *reinterpret_cast<double2 *>(out + (size_t)i * L.ncols + j0) = o;
    p = c;
    c = n;
  }
}

static inline int pick_rows_per_cta(int nrows, int col_blocks) {
  // enough CTAs to fill 148 SMs x 16 resident 128-thread CTAs, but >= 16 rows per chunk so the
  // re-read halo rows stay <= 12.5 % (and those mostly hit L2)
  int r = 64;
  while (r > 16 && (long long)col_blocks * ((nrows + r - 1) / r) < 148LL * 8) r >>= 1;
  if (r > nrows) r = nrows;
  return r < 1 ? 1 : r;
}

template <int OP>
static cudaError_t launch_march(const LevelDev &L, double shift, double omega, const double *v,
                                const double *f, double *out, const double *halo_top,
                                const double *halo_bot, cudaStream_t s) {
  int threads = L.ncols / 2;
  if (threads > 128) threads = 128;
  if (threads < 32) threads = 32;
  const int col_blocks = (L.ncols / 2 + threads - 1) / threads;
  const int rpc = pick_rows_per_cta(L.nrows, col_blocks);
  dim3 grid(col_blocks, (L.nrows + rpc - 1) / rpc);
  if (L.five)
    stencil_march_kernel<OP, true><<<grid, threads, 0, s>>>(L, shift, omega, v, f, out, halo_top, halo_bot, rpc);
  else
    stencil_march_kernel<OP, false><<<grid, threads, 0, s>>>(L, shift, omega, v, f, out, halo_top, halo_bot, rpc);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_jacobi_sweep(const LevelDev &L, double shift, double omega, const double *v_in,
                                const double *f, double *v_out, const double *halo_top,
                                const double *halo_bot, cudaStream_t s) {
  return launch_march<OP_JACOBI>(L, shift, omega, v_in, f, v_out, halo_top, halo_bot, s);
}
cudaError_t launch_residual(const LevelDev &L, double shift, const double *v, const double *f, double *r,
                            const double *halo_top, const double *halo_bot, cudaStream_t s) {
  return launch_march<OP_RESIDUAL>(L, shift, 0.0, v, f, r, halo_top, halo_bot, s);
}
cudaError_t launch_apply(const LevelDev &L, double shift, const double *x, double *y, const double *halo_top,
                         const double *halo_bot, cudaStream_t s) {
  return launch_march<OP_APPLY>(L, shift, 0.0, x, nullptr, y, halo_top, halo_bot, s);
}

}  // namespace mgcmt
