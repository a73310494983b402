// gs.cu -- Gauss-Seidel family.
//
//  * four-colour ("red-black") Gauss-Seidel / SOR: the working replacement of the reference's dead
//    gseidelrb (MGCMTSolver.py:248-279).  Colours (i%2, j%2) in the order (0,0),(1,1),(0,1),(1,0): for
//    the 5-point finest stencil that is the classical red-black sweep, and it stays a true
//    Gauss-Seidel for the 9-point Galerkin stencils of the coarse levels (no two points of one colour
//    are coupled).  CPU twin: oracle/mgcmt_oracle.py:Solver.rbgs.
//  * lexicographic Gauss-Seidel / SOR with the reference's exact semantics (MGCMTSolver.py:210-246,
//    including quirk Q6: SOR adds w (D-L)^-1 f, not w (D-wL)^-1 f).  Inherently sequential; done as a
//    wavefront sweep (t = 2i + j) by one CTA.  Kept for drop-in parity on the small grids the
//    reference can run it on, not for speed.
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

// stencil coefficient a(di,dj) at (gi, j)
template <bool FIVE>
__device__ __forceinline__ double coef(const LevelDev &L, int gi, int j, int di, int dj) {
  const double ka = di < 0 ? L.ka_lo[gi] : (di == 0 ? L.ka_di[gi] : L.ka_up[gi]);
  const double kb = dj < 0 ? L.kb_lo[j] : (dj == 0 ? L.kb_di[j] : L.kb_up[j]);
  if (FIVE) {
    if (di == 0 && dj == 0) return ka + kb;
    if (di == 0) return kb;
    if (dj == 0) return ka;
    return 0.0;
  }
  const double ma = di < 0 ? L.ma_lo[gi] : (di == 0 ? L.ma_di[gi] : L.ma_up[gi]);
  const double mb = dj < 0 ? L.mb_lo[j] : (dj == 0 ? L.mb_di[j] : L.mb_up[j]);
  return ma * kb + ka * mb;
}

template <bool FIVE>
__device__ __forceinline__ double nb(const LevelDev &L, const double *v, int i, int j, int di, int dj) {
  const int ii = i + di, jj = j + dj;
  if (ii < 0 || ii >= L.nrows || jj < 0 || jj >= L.ncols) return 0.0;
  if (FIVE && di != 0 && dj != 0) return 0.0;
  return coef<FIVE>(L, L.row0 + i, j, di, dj) * v[(size_t)ii * L.ncols + jj];
}

// ---- four-colour GS: one launch per colour (v1: simple, strided) --------------------------------
template <bool FIVE>
__global__ void rbgs_colour_kernel(LevelDev L, double shift, double omega, int pa, int pb, double *v,
                                   const double *__restrict__ f) {
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int I = blockIdx.y * blockDim.y + threadIdx.y;
  // local row with global parity pa
  const int i = 2 * I + ((pa - L.row0) & 1);
  const int j = 2 * J + pb;
  if (i >= L.nrows || j >= L.ncols) return;
  double off = 0.0;
#pragma unroll
  for (int di = -1; di <= 1; ++di)
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj)
      if (di != 0 || dj != 0) off += nb<FIVE>(L, v, i, j, di, dj);
  const double d = coef<FIVE>(L, L.row0 + i, j, 0, 0) - shift;
  const double x = v[(size_t)i * L.ncols + j];
  const double av = off + d * x;
  v[(size_t)i * L.ncols + j] = x + omega * (f[(size_t)i * L.ncols + j] - av) / d;
}

cudaError_t launch_rbgs(const LevelDev &L, double shift, double omega, int nu, double *v, const double *f,
                        cudaStream_t s) {
  static const int order[4][2] = {{0, 0}, {1, 1}, {0, 1}, {1, 0}};
  dim3 block(64, 4);
  dim3 grid((L.ncols / 2 + 63) / 64 > 0 ? (L.ncols / 2 + 63) / 64 : 1, ((L.nrows + 1) / 2 + 3) / 4);
  for (int it = 0; it < nu; ++it) {
    for (int c = 0; c < 4; ++c) {
      if (L.nrows_glob == 1 && order[c][0] == 1) continue;  // 1-D: colours are even, odd
      if (L.five)
        rbgs_colour_kernel<true><<<grid, block, 0, s>>>(L, shift, omega, order[c][0], order[c][1], v, f);
      else
        rbgs_colour_kernel<false><<<grid, block, 0, s>>>(L, shift, omega, order[c][0], order[c][1], v, f);
      count_launch();
    }
  }
  return cudaGetLastError();
}

// ---- lexicographic GS / SOR: single CTA, wavefront t = 2i + j -----------------------------------
// mode 0: y = (D - L)^-1 f * omega  -> written to c      (the f-term of SOR, quirk Q6)
// mode 1: in-place sweep  y = (D - wL)^-1 ((1-w) D + w U) v  [+ f folded in when w == 1]
template <bool FIVE>
__device__ void lex_wavefront(const LevelDev &L, double shift, double omega, int mode, double *v,
                              const double *__restrict__ f, double *c) {
  const int nr = L.nrows, nc = L.ncols;
  const int nt = 2 * (nr - 1) + nc;
  for (int t = 0; t < nt; ++t) {
    // points (i, j = t - 2i), 0 <= j < nc
    const int i_lo = max(0, (t - nc + 2) / 2), i_hi = min(nr - 1, t / 2);
    for (int i = i_lo + (int)threadIdx.x; i <= i_hi; i += blockDim.x) {
      const int j = t - 2 * i;
      if (j < 0 || j >= nc) continue;
      const size_t p = (size_t)i * nc + j;
      const double d = coef<FIVE>(L, L.row0 + i, j, 0, 0) - shift;
      if (mode == 0) {
        // (D - L) y = f  with L = -strict_lower(A):  d y_p + sum_{q<p} a_pq y_q = f_p
        double low = nb<FIVE>(L, c, i, j, -1, -1) + nb<FIVE>(L, c, i, j, -1, 0) + nb<FIVE>(L, c, i, j, -1, 1) +
                     nb<FIVE>(L, c, i, j, 0, -1);
        c[p] = (f[p] - low) / d;
      } else {
        const double upv = nb<FIVE>(L, v, i, j, 0, 1) + nb<FIVE>(L, v, i, j, 1, -1) + nb<FIVE>(L, v, i, j, 1, 0) +
                           nb<FIVE>(L, v, i, j, 1, 1);
        const double low = nb<FIVE>(L, v, i, j, -1, -1) + nb<FIVE>(L, v, i, j, -1, 0) + nb<FIVE>(L, v, i, j, -1, 1) +
                           nb<FIVE>(L, v, i, j, 0, -1);
        double b = (1.0 - omega) * d * v[p] - omega * upv;
        if (omega == 1.0) b = f[p] - upv;
        v[p] = (b - omega * low) / d;
      }
    }
    __syncthreads();
  }
}

template <bool FIVE>
__global__ void __launch_bounds__(1024)
gs_lex_kernel(LevelDev L, double shift, double omega, int nu, double *v, const double *__restrict__ f,
              double *c) {
  const size_t n = (size_t)L.nrows * L.ncols;
  const bool fold = (omega == 1.0);
  if (!fold) {
    lex_wavefront<FIVE>(L, shift, omega, 0, v, f, c);
    __syncthreads();
  }
  for (int it = 0; it < nu; ++it) {
    lex_wavefront<FIVE>(L, shift, omega, 1, v, f, c);
    if (!fold) {
      for (size_t p = threadIdx.x; p < n; p += blockDim.x) v[p] += omega * c[p];
      __syncthreads();
    }
  }
}

cudaError_t launch_gs_lex(const LevelDev &L, double shift, double omega, int nu, double *v, const double *f,
                          double *scratch, cudaStream_t s) {
  int threads = L.nrows >= 1024 ? 1024 : (L.nrows < 32 ? 32 : L.nrows);
  if (L.five)
    gs_lex_kernel<true><<<1, threads, 0, s>>>(L, shift, omega, nu, v, f, scratch);
  else
    gs_lex_kernel<false><<<1, threads, 0, s>>>(L, shift, omega, nu, v, f, scratch);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mgcmt
