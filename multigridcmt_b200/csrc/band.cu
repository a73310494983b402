// band.cu -- kernels of the general banded complex128 path (1-D multiband Hamiltonians).
//
// The reference's ThesisProblem driver (ThesisProblem.py:26-104) hands MGCMTSolver.vcycle a complex sparse matrix
// assembled by PotWellSolver.makeMatrix (PotWellSolver.py:54-233): a 4x4 (or 6x6) block matrix of tridiagonal
// blocks, i.e. a handful of diagonals at offsets {0, +-1, +-n +- {0,1}, +-2n ...}.  MGCMTSolver treats it as a 1-D
// problem of 4n unknowns (transfer operators smear across band boundaries, as in the reference).  Operators here are
// kept by DIAGONALS:  vals[k*n + i] = A[i, i + offs[k]]  (0 where the column falls outside), complex interleaved.
// The Galerkin product R A P of such a matrix with the 1-D full-weighting / linear-interpolation pair is again a
// few diagonals (offset d -> {d/2 - 1, d/2, d/2 + 1}), so every level is a coalesced streaming stencil.
//
// Nothing here is bandwidth-critical at the sizes the reference runs (1024 unknowns): these kernels are
// latency-bound and written for clarity; the 2-D well path (fused.cu) is the throughput path.
#include "kernels.h"

namespace mgcmt {

namespace {

using cplx = double2;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return make_double2(s * a.x, s * a.y); }
// c -= a * b
__device__ __forceinline__ void cfms(cplx &c, cplx a, cplx b) {
  c.x -= a.x * b.x - a.y * b.y;
  c.y -= a.x * b.y + a.y * b.x;
}
__device__ __forceinline__ cplx crecip(cplx a) {
  // Smith's algorithm: no overflow for large |a|
  if (fabs(a.x) >= fabs(a.y)) {
    const double r = a.y / a.x, d = a.x + a.y * r;
    return make_double2(1.0 / d, -r / d);
  }
  const double r = a.x / a.y, d = a.x * r + a.y;
  return make_double2(r / d, -1.0 / d);
}

// (A - shift I) x at row i
__device__ __forceinline__ cplx row_apply(const BandDev &L, double shift, const cplx *x, int i) {
  cplx acc = make_double2(0.0, 0.0);
  for (int k = 0; k < L.ndiag; ++k) {
    const int j = i + L.offs[k];
    if (j < 0 || j >= L.n) continue;
    const cplx a = L.vals[(size_t)k * L.n + i];
    const cplx xv = x[j];
    acc.x += a.x * xv.x - a.y * xv.y;
    acc.y += a.x * xv.y + a.y * xv.x;
  }
  const cplx xi = x[i];
  acc.x -= shift * xi.x;
  acc.y -= shift * xi.y;
  return acc;
}

__global__ void band_apply_kernel(BandDev L, double shift, const cplx *x, cplx *y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n; i += gridDim.x * blockDim.x)
    y[i] = row_apply(L, shift, x, i);
}

// v_out = v_in + omega (f - (A - shift) v_in) / (d - shift)     (MGCMTSolver.py:182-208, one sweep)
__global__ void band_jacobi_kernel(BandDev L, double shift, double omega, const cplx *vin, const cplx *f, cplx *vout) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n; i += gridDim.x * blockDim.x) {
    const cplx r = csub(f[i], row_apply(L, shift, vin, i));
    cplx d = L.vals[(size_t)L.idiag * L.n + i];
    d.x -= shift;
    vout[i] = cadd(vin[i], cscale(omega, cmul(r, crecip(d))));
  }
}

// r_c[J] = 1/4 r[2J] + 1/2 r[2J+1] + 1/4 r[2J+2],  r = f - (A - shift) v   (MGCMTSolver.py:315; last row truncated)
__global__ void band_residual_restrict_kernel(BandDev L, double shift, const cplx *v, const cplx *f, cplx *rc) {
  const int nc = L.n >> 1;
  for (int J = blockIdx.x * blockDim.x + threadIdx.x; J < nc; J += gridDim.x * blockDim.x) {
    const int i0 = 2 * J;
    const cplx r0 = csub(f[i0], row_apply(L, shift, v, i0));
    const cplx r1 = csub(f[i0 + 1], row_apply(L, shift, v, i0 + 1));
    cplx acc = make_double2(0.25 * r0.x + 0.5 * r1.x, 0.25 * r0.y + 0.5 * r1.y);
    if (i0 + 2 < L.n) {
      const cplx r2 = csub(f[i0 + 2], row_apply(L, shift, v, i0 + 2));
      acc.x += 0.25 * r2.x;
      acc.y += 0.25 * r2.y;
    }
    rc[J] = acc;
  }
}

// v += P e:  odd fine 2J+1 <- e[J];  even fine 2J <- (e[J-1] + e[J]) / 2   (MGCMTStencilMaker.py:27-54)
__global__ void band_prolong_correct_kernel(int n, const cplx *ec, cplx *v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int J = i >> 1;
    cplx e = ec[J];
    if (!(i & 1)) {
      e = cscale(0.5, e);
      if (J > 0) {
        const cplx w = ec[J - 1];
        e.x += 0.5 * w.x;
        e.y += 0.5 * w.y;
      }
    }
    v[i] = cadd(v[i], e);
  }
}

// Galerkin product: coarse diagonal kc, row J:  A_c[J, J+D] = sum_{a,b} R[J,2J+a] A[2J+a, 2(J+D)+b] P[2(J+D)+b, J+D]
// with R weights (1/4,1/2,1/4), P weights (1/2,1,1/2).  lut[kc*5 + (b-a+2)] = fine diagonal holding offset 2D+b-a,
// or -1.  (MGCMTSolver.py:318: restriction_matrix * A * interpolation_matrix)
__global__ void band_galerkin_kernel(BandDev F, int nc, int ndiag_c, const int *offs_c, const int *lut, cplx *vals_c) {
  const long long total = (long long)nc * ndiag_c;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(t / nc), J = (int)(t % nc);
    const int K = J + offs_c[kc];
    cplx acc = make_double2(0.0, 0.0);
    if (K >= 0 && K < nc) {
      for (int a = 0; a < 3; ++a) {
        const int ia = 2 * J + a;
        if (ia >= F.n) continue;
        const double wr = (a == 1) ? 0.5 : 0.25;
        for (int b = 0; b < 3; ++b) {
          const int ib = 2 * K + b;
          if (ib >= F.n) continue;
          const int kf = lut[kc * 5 + (b - a + 2)];
          if (kf < 0) continue;
          const double w = wr * ((b == 1) ? 1.0 : 0.5);
          const cplx av = F.vals[(size_t)kf * F.n + ia];
          acc.x += w * av.x;
          acc.y += w * av.y;
        }
      }
    }
    vals_c[(size_t)kc * nc + J] = acc;
  }
}

// Forward substitution sweep (one warp):  solve  (D + wl * strict_lower(A_s)) y = cf f + cd D v - cu strict_upper(A_s) v
// with A_s = A - shift I, D = diag(A_s); then v_out = oscale * y + g.   gseidel (MGCMTSolver.py:210-227) is
// (wl, cf, cd, cu, oscale) = (1, 1, 0, 1, 1); the two solves of sor (:229-246, quirk Q6) are (1, 1, 0, 0, w) and
// (w, 0, 1-w, w, 1).  Rows are taken 32 at a time: everything that refers to rows before the chunk (already final, in
// y) or to old values is summed in parallel, the 32x32 triangular block left over is solved by broadcasting one
// finished unknown per step.  v_in may alias v_out and y (in-place Gauss-Seidel).
__global__ void __launch_bounds__(32) band_lower_solve_kernel(BandDev L, double shift, double wl, double cf, double cd,
                                                              double cu, double oscale, const cplx *vin, const cplx *f,
                                                              cplx *y, const cplx *g, cplx *vout, int scan) {
  __shared__ cplx coef[32][33];  // coef[r][lane]: wl * A[c0+lane, c0+r] for r < lane
  const int lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  for (int c0 = 0; c0 < L.n; c0 += 32) {
    const int i = c0 + lane;
    const bool valid = i < L.n;
    cplx t = make_double2(0.0, 0.0), dinv = make_double2(1.0, 0.0);
    unsigned mask = 0u;
    if (valid) {
      cplx d = L.vals[(size_t)L.idiag * L.n + i];
      d.x -= shift;
      dinv = crecip(d);
      if (cf != 0.0) t = cscale(cf, f[i]);
      if (cd != 0.0) t = cadd(t, cscale(cd, cmul(d, vin[i])));
      for (int k = 0; k < L.ndiag; ++k) {
        const int off = L.offs[k];
        if (off == 0) continue;
        const int j = i + off;
        if (j < 0 || j >= L.n) continue;
        const cplx a = L.vals[(size_t)k * L.n + i];
        if (off > 0) {
          if (cu != 0.0) cfms(t, cscale(cu, a), vin[j]);
        } else if (j < c0) {
          cfms(t, cscale(wl, a), y[j]);
        } else {
          coef[j - c0][lane] = cscale(wl, a);
          mask |= 1u << (j - c0);
        }
      }
    }
    __syncwarp(full);
    cplx mine = make_double2(0.0, 0.0);
    // Usual case (every level of the multiband hierarchies except the few coarsest): inside the chunk a row couples
    // only to its predecessor, so the block is the first-order recurrence x_l = p_l + q_l x_{l-1} -- solved by a
    // 5-step inclusive scan over the affine maps instead of 32 dependent steps.
    const unsigned pred = lane > 0 ? (1u << (lane - 1)) : 0u;
    if (scan && __all_sync(full, (mask & ~pred) == 0u)) {
      cplx p = cmul(t, dinv);
      cplx q = make_double2(0.0, 0.0);
      if (mask) q = cscale(-1.0, cmul(coef[lane - 1][lane], dinv));
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        cplx pp, qq;
        pp.x = __shfl_up_sync(full, p.x, s);
        pp.y = __shfl_up_sync(full, p.y, s);
        qq.x = __shfl_up_sync(full, q.x, s);
        qq.y = __shfl_up_sync(full, q.y, s);
        if (lane >= s) {
          p = cadd(p, cmul(q, pp));
          q = cmul(q, qq);
        }
      }
      mine = p;
    } else {
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const cplx cand = cmul(t, dinv);
        cplx xr;
        xr.x = __shfl_sync(full, cand.x, r);
        xr.y = __shfl_sync(full, cand.y, r);
        if (lane == r) mine = cand;
        if ((mask >> r) & 1u) cfms(t, coef[r][lane], xr);
      }
    }
    if (valid) {
      y[i] = mine;
      cplx o = cscale(oscale, mine);
      if (g) o = cadd(o, g[i]);
      vout[i] = o;
    }
    __syncwarp(full);
    __threadfence_block();
  }
}

// ---- the same sweep in two launches (default): everything that only needs OLD values first, fully parallel ...
//   rhs_i = cf f_i + cd d_i v_i - cu sum_{off > 0} A_s[i, i+off] v[i+off]
__global__ void band_upper_rhs_kernel(BandDev L, double shift, double cf, double cd, double cu, const cplx *vin,
                                      const cplx *f, cplx *rhs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n; i += gridDim.x * blockDim.x) {
    cplx t = make_double2(0.0, 0.0);
    if (cf != 0.0) t = cscale(cf, f[i]);
    if (cd != 0.0) {
      cplx d = L.vals[(size_t)L.idiag * L.n + i];
      d.x -= shift;
      t = cadd(t, cscale(cd, cmul(d, vin[i])));
    }
    if (cu != 0.0)
      for (int k = L.idiag + 1; k < L.ndiag; ++k) {   // offsets ascend: the diagonals after the main one are the upper ones
        const int j = i + L.offs[k];
        if (j >= L.n) break;
        cfms(t, cscale(cu, L.vals[(size_t)k * L.n + i]), vin[j]);
      }
    rhs[i] = t;
  }
}

// ... then the forward substitution proper (one warp): (D + wl strict_lower(A_s)) y = rhs, v_out = oscale y + g.  Only
// the lower diagonals are touched here; their offsets sit in shared memory and their loads are issued four at a time.
__global__ void __launch_bounds__(32) band_lower_solve2_kernel(BandDev L, double shift, double wl, double oscale,
                                                               const cplx *rhs, cplx *y, const cplx *g, cplx *vout, int scan) {
  __shared__ cplx coef[32][33];
  __shared__ int s_off[kBandMaxDiags];
  const int lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int nlow = L.idiag;   // offsets ascend: diagonals 0 .. idiag-1 are the lower ones
  for (int k = lane; k < nlow; k += 32) s_off[k] = L.offs[k];
  __syncwarp(full);
  for (int c0 = 0; c0 < L.n; c0 += 32) {
    const int i = c0 + lane;
    const bool valid = i < L.n;
    cplx t = make_double2(0.0, 0.0), dinv = make_double2(1.0, 0.0);
    unsigned mask = 0u;
    if (valid) {
      cplx d = L.vals[(size_t)L.idiag * L.n + i];
      d.x -= shift;
      dinv = crecip(d);
      t = rhs[i];
    }
    for (int base = 0; base < nlow; base += 4) {
      cplx a[4], xv[4];
      int jj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = base + u;
        int j = -1;
        if (valid && k < nlow) j = i + s_off[k];
        if (j < 0) j = -1;
        jj[u] = j;
        a[u] = (j >= 0) ? L.vals[(size_t)k * L.n + i] : make_double2(0.0, 0.0);
        xv[u] = (j >= 0 && j < c0) ? y[j] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (jj[u] < 0) continue;
        if (jj[u] < c0) {
          cfms(t, cscale(wl, a[u]), xv[u]);
        } else {
          coef[jj[u] - c0][lane] = cscale(wl, a[u]);
          mask |= 1u << (jj[u] - c0);
        }
      }
    }
    __syncwarp(full);
    cplx mine = make_double2(0.0, 0.0);
    const unsigned pred = lane > 0 ? (1u << (lane - 1)) : 0u;
    if (scan && __all_sync(full, (mask & ~pred) == 0u)) {
      cplx p = cmul(t, dinv);
      cplx q = make_double2(0.0, 0.0);
      if (mask) q = cscale(-1.0, cmul(coef[lane - 1][lane], dinv));
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        cplx pp, qq;
        pp.x = __shfl_up_sync(full, p.x, s);
        pp.y = __shfl_up_sync(full, p.y, s);
        qq.x = __shfl_up_sync(full, q.x, s);
        qq.y = __shfl_up_sync(full, q.y, s);
        if (lane >= s) {
          p = cadd(p, cmul(q, pp));
          q = cmul(q, qq);
        }
      }
      mine = p;
    } else {
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const cplx cand = cmul(t, dinv);
        cplx xr;
        xr.x = __shfl_sync(full, cand.x, r);
        xr.y = __shfl_sync(full, cand.y, r);
        if (lane == r) mine = cand;
        if ((mask >> r) & 1u) cfms(t, coef[r][lane], xr);
      }
    }
    if (valid) {
      y[i] = mine;
      cplx o = cscale(oscale, mine);
      if (g) o = cadd(o, g[i]);
      vout[i] = o;
    }
    __syncwarp(full);
    __threadfence_block();
  }
}

// dense (A - shift I) | I, row-major n x 2n, for the coarsest solve (MGCMTSolver.py:306: spsolve(shifted_matrix, f))
__global__ void band_build_dense_kernel(BandDev L, double shift, cplx *aug) {
  const int n = L.n;
  const long long total = (long long)n * 2 * n;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / (2 * n)), c = (int)(t % (2 * n));
    cplx val = make_double2(0.0, 0.0);
    if (c >= n) {
      if (c - n == i) val.x = 1.0;
    } else {
      for (int k = 0; k < L.ndiag; ++k)
        if (i + L.offs[k] == c) val = L.vals[(size_t)k * n + i];
      if (c == i) val.x -= shift;
    }
    aug[t] = val;
  }
}

// Gauss-Jordan with partial pivoting on the augmented matrix, one CTA (n <= kBandMaxCoarse).  status != 0: singular.
constexpr int kGjThreads = 256;
__global__ void __launch_bounds__(kGjThreads) band_gauss_jordan_kernel(int n, cplx *aug, int *status) {
  __shared__ cplx fac[kBandMaxCoarse];
  __shared__ double best_v[kGjThreads];
  __shared__ int best_i[kGjThreads];
  const int tid = threadIdx.x, w = 2 * n;
  for (int p = 0; p < n; ++p) {
    double bv = -1.0;
    int bi = p;
    for (int r = p + tid; r < n; r += kGjThreads) {
      const cplx a = aug[(size_t)r * w + p];
      const double m = a.x * a.x + a.y * a.y;
      if (m > bv) { bv = m; bi = r; }
    }
    best_v[tid] = bv;
    best_i[tid] = bi;
    __syncthreads();
    for (int s = kGjThreads / 2; s > 0; s >>= 1) {
      if (tid < s) {
        const double ov = best_v[tid + s];
        const int oi = best_i[tid + s];
        if (ov > best_v[tid] || (ov == best_v[tid] && oi < best_i[tid])) { best_v[tid] = ov; best_i[tid] = oi; }
      }
      __syncthreads();
    }
    const int piv = best_i[0];
    const double pm = best_v[0];
    __syncthreads();
    if (!(pm > 0.0)) {
      if (tid == 0) *status = p + 1;
      return;
    }
    if (piv != p)
      for (int c = tid; c < w; c += kGjThreads) {
        const cplx a = aug[(size_t)p * w + c];
        aug[(size_t)p * w + c] = aug[(size_t)piv * w + c];
        aug[(size_t)piv * w + c] = a;
      }
    __syncthreads();
    const cplx inv = crecip(aug[(size_t)p * w + p]);
    for (int r = tid; r < n; r += kGjThreads) fac[r] = aug[(size_t)r * w + p];
    __syncthreads();
    for (int c = tid; c < w; c += kGjThreads) aug[(size_t)p * w + c] = cmul(aug[(size_t)p * w + c], inv);
    __syncthreads();
    for (long long t = tid; t < (long long)n * w; t += kGjThreads) {
      const int r = (int)(t / w), c = (int)(t % w);
      if (r == p) continue;
      cfms(aug[t], fac[r], aug[(size_t)p * w + c]);
    }
    __syncthreads();
  }
}

// y = Ainv f, Ainv = right half of the augmented matrix; one warp per row
__global__ void band_gemv_kernel(int n, const cplx *aug, const cplx *f, cplx *y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  cplx acc = make_double2(0.0, 0.0);
  for (int c = lane; c < n; c += 32) {
    const cplx a = aug[(size_t)warp * 2 * n + n + c], x = f[c];
    acc.x += a.x * x.x - a.y * x.y;
    acc.y += a.x * x.y + a.y * x.x;
  }
  acc.x = warp_sum(acc.x);
  acc.y = warp_sum(acc.y);
  if (lane == 0) y[warp] = acc;
}

}  // namespace
int g_band_gs_scan = 1;   // mgcmt_set_option("band_gs_scan", 0|1): scan form of the in-chunk recurrence (A/B, tests)
int g_band_gs_split = 1;  // mgcmt_set_option("band_gs_split", 0|1): old-value part of the sweep in its own parallel launch
namespace {

int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

cudaError_t launch_band_apply(const BandDev &L, double shift, const double *x, double *y, cudaStream_t s) {
  band_apply_kernel<<<blocks_for(L.n), 256, 0, s>>>(L, shift, (const cplx *)x, (cplx *)y);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_band_jacobi(const BandDev &L, double shift, double omega, const double *vin, const double *f,
                               double *vout, cudaStream_t s) {
  band_jacobi_kernel<<<blocks_for(L.n), 256, 0, s>>>(L, shift, omega, (const cplx *)vin, (const cplx *)f, (cplx *)vout);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_band_residual_restrict(const BandDev &L, double shift, const double *v, const double *f, double *rc,
                                          cudaStream_t s) {
  band_residual_restrict_kernel<<<blocks_for(L.n / 2), 256, 0, s>>>(L, shift, (const cplx *)v, (const cplx *)f, (cplx *)rc);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_band_prolong_correct(int n_fine, const double *ec, double *v, cudaStream_t s) {
  band_prolong_correct_kernel<<<blocks_for(n_fine), 256, 0, s>>>(n_fine, (const cplx *)ec, (cplx *)v);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_band_galerkin(const BandDev &F, int nc, int ndiag_c, const int *offs_c, const int *lut, double *vals_c,
                                 cudaStream_t s) {
  band_galerkin_kernel<<<blocks_for((long long)nc * ndiag_c), 256, 0, s>>>(F, nc, ndiag_c, offs_c, lut, (cplx *)vals_c);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_band_lower_solve(const BandDev &L, double shift, double wl, double cf, double cd, double cu,
                                    double oscale, const double *vin, const double *f, double *y, const double *g,
                                    double *vout, double *rhs_scratch, cudaStream_t s) {
  if (g_band_gs_split && rhs_scratch) {
    band_upper_rhs_kernel<<<blocks_for(L.n), 256, 0, s>>>(L, shift, cf, cd, cu, (const cplx *)vin, (const cplx *)f,
                                                          (cplx *)rhs_scratch);
    band_lower_solve2_kernel<<<1, 32, 0, s>>>(L, shift, wl, oscale, (const cplx *)rhs_scratch, (cplx *)y, (const cplx *)g,
                                              (cplx *)vout, g_band_gs_scan);
    count_launch(2);
    return cudaGetLastError();
  }
  band_lower_solve_kernel<<<1, 32, 0, s>>>(L, shift, wl, cf, cd, cu, oscale, (const cplx *)vin, (const cplx *)f, (cplx *)y,
                                           (const cplx *)g, (cplx *)vout, g_band_gs_scan);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_band_inverse(const BandDev &L, double shift, double *aug, int *status, cudaStream_t s) {
  if (L.n > kBandMaxCoarse) return cudaErrorInvalidValue;
  band_build_dense_kernel<<<blocks_for((long long)L.n * 2 * L.n), 256, 0, s>>>(L, shift, (cplx *)aug);
  band_gauss_jordan_kernel<<<1, kGjThreads, 0, s>>>(L.n, (cplx *)aug, status);
  count_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_band_gemv(int n, const double *aug, const double *f, double *y, cudaStream_t s) {
  band_gemv_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(n, (const cplx *)aug, (const cplx *)f, (cplx *)y);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mgcmt
