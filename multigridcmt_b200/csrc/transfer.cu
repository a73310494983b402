// transfer.cu -- grid-transfer kernels and the Galerkin coarsening of the tridiagonal factors.
//
// Reference operators restated (MGCMTStencilMaker.py:27-78, single-level jump):
//   interpolation P (1-D):  coarse j sits on fine 2j+1;  P[2j,j] = 1/2, P[2j+1,j] = 1, P[2j+2,j] = 1/2
//                           (the entry 2j+2 = n_fine of the last column does not exist)
//   restriction   R (1-D):  1/2 P^T  => rows [1/4 1/2 1/4], last row truncated to [1/4 1/2]
//   2-D:                    P2 = P (x) P,  R2 = 1/4 P2^T = R (x) R
//   Galerkin:               A_c = R A P   (MGCMTSolver.py:318) -- for A = Ma (x) Kb + Ka (x) Mb this is
//                           (R Ma P) (x) (R Kb P) + (R Ka P) (x) (R Mb P): one tridiagonal product per factor.
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

// ---------------------------------------------------------------------------------------------
// coarse tridiagonal = R T P for tridiagonal T = (lo, di, up); one thread per coarse row
// ---------------------------------------------------------------------------------------------
__global__ void galerkin_tridiag_kernel(int nf, const double *__restrict__ lo, const double *__restrict__ di,
                                        const double *__restrict__ up, double *__restrict__ clo,
                                        double *__restrict__ cdi, double *__restrict__ cup) {
  const int nc = nf / 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nc) return;
  const double wr[3] = {0.25, 0.5, 0.25};
  const double wp[3] = {0.5, 1.0, 0.5};
  double res[3];
  for (int dj = -1; dj <= 1; ++dj) {
    const int jp = j + dj;
    double acc = 0.0;
    if (jp >= 0 && jp < nc) {
      // (R T)[j, b] for the fine columns b that P's column jp touches, then times P[b, jp]
      for (int s = 0; s < 3; ++s) {
        const int b = 2 * jp + s;
        if (b >= nf) continue;
        double rt = 0.0;
        for (int q = 0; q < 3; ++q) {
          const int a = 2 * j + q;
          if (a >= nf) continue;
          double t = 0.0;
          if (b == a - 1) t = lo[a];
          else if (b == a) t = di[a];
          else if (b == a + 1) t = up[a];
          else continue;
          rt += wr[q] * t;
        }
        acc += rt * wp[s];
      }
    }
    res[dj + 1] = acc;
  }
  clo[j] = (j > 0) ? res[0] : 0.0;
  cdi[j] = res[1];
  cup[j] = (j + 1 < nc) ? res[2] : 0.0;
}

cudaError_t launch_galerkin_tridiag(int n_fine, const double *lo, const double *di, const double *up,
                                    double *clo, double *cdi, double *cup, cudaStream_t s) {
  const int nc = n_fine / 2;
  const int threads = 128;
  galerkin_tridiag_kernel<<<(nc + threads - 1) / threads, threads, 0, s>>>(n_fine, lo, di, up, clo, cdi, cup);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// plain restriction: coarse = R fine
// ---------------------------------------------------------------------------------------------
__global__ void restrict_kernel(int nrf, int ncf, int coarsen_rows, const double *__restrict__ fine,
                                double *__restrict__ coarse) {
  const int ncc = ncf / 2;
  const int nrc = coarsen_rows ? nrf / 2 : nrf;
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int I = blockIdx.y * blockDim.y + threadIdx.y;
  if (J >= ncc || I >= nrc) return;
  const double w[3] = {0.25, 0.5, 0.25};
  double acc = 0.0;
  if (coarsen_rows) {
    for (int a = 0; a < 3; ++a) {
      const int i = 2 * I + a;
      if (i >= nrf) continue;
      double rowacc = 0.0;
      for (int b = 0; b < 3; ++b) {
        const int j = 2 * J + b;
        if (j >= ncf) continue;
        rowacc += w[b] * fine[(size_t)i * ncf + j];
      }
      acc += w[a] * rowacc;
    }
  } else {
    for (int b = 0; b < 3; ++b) {
      const int j = 2 * J + b;
      if (j >= ncf) continue;
      acc += w[b] * fine[(size_t)I * ncf + j];
    }
  }
  coarse[(size_t)I * ncc + J] = acc;
}

cudaError_t launch_restrict(const LevelDev &Lf, bool coarsen_rows, const double *fine, double *coarse,
                            cudaStream_t s) {
  const int ncc = Lf.ncols / 2;
  const int nrc = coarsen_rows ? Lf.nrows / 2 : Lf.nrows;
  dim3 block(64, 4);
  dim3 grid((ncc + 63) / 64, (nrc + 3) / 4);
  restrict_kernel<<<grid, block, 0, s>>>(Lf.nrows, Lf.ncols, coarsen_rows ? 1 : 0, fine, coarse);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// fused  r_c = R (f - (A - shift I) v)   (MGCMTSolver.py:315)
// A CTA owns a CI x CJ coarse tile: it stages the (2CI+3) x (2CJ+3) fine v tile in shared memory,
// forms the (2CI+1) x (2CJ+1) fine residuals there, and full-weights them.
// ---------------------------------------------------------------------------------------------
constexpr int RR_CI = 8;
constexpr int RR_CJ = 64;
constexpr int RR_VR = 2 * RR_CI + 3;  // v rows staged
constexpr int RR_VC = 2 * RR_CJ + 3;  // v cols staged
constexpr int RR_VP = RR_VC + 1;      // padded pitch
constexpr int RR_RR = 2 * RR_CI + 1;
constexpr int RR_RC = 2 * RR_CJ + 1;
constexpr int RR_RP = RR_RC + 1;

template <bool FIVE>
__global__ void __launch_bounds__(256)
residual_restrict_kernel(LevelDev L, double shift, const double *__restrict__ v, const double *__restrict__ f,
                         double *__restrict__ rc) {
  __shared__ double sv[RR_VR * RR_VP];
  __shared__ double sr[RR_RR * RR_RP];
  const int ncc = L.ncols / 2;
  const int nrc = L.nrows / 2;
  const int I0 = blockIdx.y * RR_CI, J0 = blockIdx.x * RR_CJ;
  const int fi0 = 2 * I0 - 1, fj0 = 2 * J0 - 1;  // fine coords of sv[0][0]
  const int tid = threadIdx.x;

  for (int idx = tid; idx < RR_VR * RR_VC; idx += 256) {
    const int r = idx / RR_VC, c = idx - r * RR_VC;
    const int i = fi0 + r, j = fj0 + c;
    double x = 0.0;
    if (i >= 0 && i < L.nrows && j >= 0 && j < L.ncols) x = v[(size_t)i * L.ncols + j];
    sv[r * RR_VP + c] = x;
  }
  __syncthreads();

  for (int idx = tid; idx < RR_RR * RR_RC; idx += 256) {
    const int r = idx / RR_RC, c = idx - r * RR_RC;
    const int i = 2 * I0 + r, j = 2 * J0 + c;  // fine point; sv index (r+1, c+1)
    double res = 0.0;
    if (i < L.nrows && j < L.ncols) {
      const double *s = &sv[(r + 1) * RR_VP + (c + 1)];
      const int gi = L.row0 + i;
      const double kbl = L.kb_lo[j], kbd = L.kb_di[j], kbu = L.kb_up[j];
      const double kal = L.ka_lo[gi], kad = L.ka_di[gi], kau = L.ka_up[gi];
      double av;
      if (FIVE) {
        av = kal * s[-RR_VP] + ((kbl * s[-1] + kbd * s[0] + kbu * s[1]) + kad * s[0]) + kau * s[RR_VP] - shift * s[0];
      } else {
        const double mbl = L.mb_lo[j], mbd = L.mb_di[j], mbu = L.mb_up[j];
        const double mal = L.ma_lo[gi], mad = L.ma_di[gi], mau = L.ma_up[gi];
        const double *sp = s - RR_VP, *sn = s + RR_VP;
        const double tp = kbl * sp[-1] + kbd * sp[0] + kbu * sp[1];
        const double tc = kbl * s[-1] + kbd * s[0] + kbu * s[1];
        const double tn = kbl * sn[-1] + kbd * sn[0] + kbu * sn[1];
        const double up_ = mbl * sp[-1] + mbd * sp[0] + mbu * sp[1];
        const double uc = mbl * s[-1] + mbd * s[0] + mbu * s[1];
        const double un = mbl * sn[-1] + mbd * sn[0] + mbu * sn[1];
        av = (mal * tp + kal * up_) + (mad * tc + kad * uc) + (mau * tn + kau * un) - shift * s[0];
      }
      res = f[(size_t)i * L.ncols + j] - av;
    }
    sr[r * RR_RP + c] = res;
  }
  __syncthreads();

  for (int idx = tid; idx < RR_CI * RR_CJ; idx += 256) {
    const int r = idx / RR_CJ, c = idx - r * RR_CJ;
    const int I = I0 + r, J = J0 + c;
    if (I < nrc && J < ncc) {
      const double *s = &sr[(2 * r) * RR_RP + 2 * c];
      const double a0 = 0.25 * s[0] + 0.5 * s[1] + 0.25 * s[2];
      const double a1 = 0.25 * s[RR_RP] + 0.5 * s[RR_RP + 1] + 0.25 * s[RR_RP + 2];
      const double a2 = 0.25 * s[2 * RR_RP] + 0.5 * s[2 * RR_RP + 1] + 0.25 * s[2 * RR_RP + 2];
      rc[(size_t)I * ncc + J] = 0.25 * a0 + 0.5 * a1 + 0.25 * a2;
    }
  }
}

// 1-D (single row, no row coarsening): one thread per coarse point
__global__ void residual_restrict_1d_kernel(LevelDev L, double shift, const double *__restrict__ v,
                                            const double *__restrict__ f, double *__restrict__ rc) {
  const int ncc = L.ncols / 2;
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  if (J >= ncc) return;
  const double kad = L.ka_di[L.row0];  // single row: A = ma_di*Kb + ka_di*Mb
  const double mad = L.five ? 1.0 : L.ma_di[L.row0];
  const double w[3] = {0.25, 0.5, 0.25};
  double acc = 0.0;
  for (int b = 0; b < 3; ++b) {
    const int j = 2 * J + b;
    if (j >= L.ncols) continue;
    const double xl = j > 0 ? v[j - 1] : 0.0, x0 = v[j], xr = j + 1 < L.ncols ? v[j + 1] : 0.0;
    double av = mad * (L.kb_lo[j] * xl + L.kb_di[j] * x0 + L.kb_up[j] * xr);
    if (L.five)
      av += kad * x0;
    else
      av += kad * (L.mb_lo[j] * xl + L.mb_di[j] * x0 + L.mb_up[j] * xr);
    av -= shift * x0;
    acc += w[b] * (f[j] - av);
  }
  rc[J] = acc;
}

cudaError_t launch_residual_restrict(const LevelDev &Lf, bool coarsen_rows, double shift, const double *v,
                                     const double *f, double *rc, cudaStream_t s) {
  if (!coarsen_rows) {
    const int ncc = Lf.ncols / 2;
    residual_restrict_1d_kernel<<<(ncc + 127) / 128, 128, 0, s>>>(Lf, shift, v, f, rc);
    count_launch();
  return cudaGetLastError();
  }
  const int ncc = Lf.ncols / 2, nrc = Lf.nrows / 2;
  dim3 grid((ncc + RR_CJ - 1) / RR_CJ, (nrc + RR_CI - 1) / RR_CI);
  if (Lf.five)
    residual_restrict_kernel<true><<<grid, 256, 0, s>>>(Lf, shift, v, f, rc);
  else
    residual_restrict_kernel<false><<<grid, 256, 0, s>>>(Lf, shift, v, f, rc);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// prolongation: fine (+)= P coarse.  One thread per coarse point writes its 2x2 (or 1x2) fine block:
//   fine(2I  ,2J  ) = 1/4 (e[I-1,J-1] + e[I-1,J] + e[I,J-1] + e[I,J])
//   fine(2I  ,2J+1) = 1/2 (e[I-1,J] + e[I,J])
//   fine(2I+1,2J  ) = 1/2 (e[I,J-1] + e[I,J])
//   fine(2I+1,2J+1) = e[I,J]
// ---------------------------------------------------------------------------------------------
template <bool ACC>
__global__ void prolong_kernel(int nrf, int ncf, int coarsen_rows, const double *__restrict__ e,
                               double *__restrict__ fine) {
  const int ncc = ncf / 2;
  const int nrc = coarsen_rows ? nrf / 2 : nrf;
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int I = blockIdx.y * blockDim.y + threadIdx.y;
  if (J >= ncc || I >= nrc) return;
  const double e11 = e[(size_t)I * ncc + J];
  const double e10 = J > 0 ? e[(size_t)I * ncc + J - 1] : 0.0;
  if (coarsen_rows) {
    const double e01 = I > 0 ? e[(size_t)(I - 1) * ncc + J] : 0.0;
    const double e00 = (I > 0 && J > 0) ? e[(size_t)(I - 1) * ncc + J - 1] : 0.0;
    double2 top, bot;
    top.x = 0.25 * ((e00 + e01) + (e10 + e11));
    top.y = 0.5 * (e01 + e11);
    bot.x = 0.5 * (e10 + e11);
    bot.y = e11;
    double2 *pt = reinterpret_cast<double2 *>(fine + (size_t)(2 * I) * ncf + 2 * J);
    double2 *pb = reinterpret_cast<double2 *>(fine + (size_t)(2 * I + 1) * ncf + 2 * J);
    if (ACC) {
      double2 a = *pt, b = *pb;
      top.x += a.x; top.y += a.y; bot.x += b.x; bot.y += b.y;
    }
    *pt = top;
    *pb = bot;
  } else {
    double2 o;
    o.x = 0.5 * (e10 + e11);
    o.y = e11;
    double2 *pt = reinterpret_cast<double2 *>(fine + (size_t)I * ncf + 2 * J);
    if (ACC) {
      double2 a = *pt;
      o.x += a.x; o.y += a.y;
    }
    *pt = o;
  }
}

cudaError_t launch_prolong(const LevelDev &Lf, bool coarsen_rows, bool accumulate, const double *coarse,
                           double *fine, cudaStream_t s) {
  const int ncc = Lf.ncols / 2;
  const int nrc = coarsen_rows ? Lf.nrows / 2 : Lf.nrows;
  dim3 block(64, 4);
  dim3 grid((ncc + 63) / 64, (nrc + 3) / 4);
  if (accumulate)
    prolong_kernel<true><<<grid, block, 0, s>>>(Lf.nrows, Lf.ncols, coarsen_rows ? 1 : 0, coarse, fine);
  else
    prolong_kernel<false><<<grid, block, 0, s>>>(Lf.nrows, Lf.ncols, coarsen_rows ? 1 : 0, coarse, fine);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mgcmt
