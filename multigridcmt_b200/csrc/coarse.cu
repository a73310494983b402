// coarse.cu -- exact solve on the coarsest level (replaces `spsolve(shifted_matrix, f)`,
// MGCMTSolver.py:305-308).
//
// The coarsest operator (A_L - shift I) is small (lowest_level^2 unknowns in 2-D) and, for shifts
// inside the spectrum, indefinite -- so it is inverted with partial pivoting.  The dense inverse is
// formed ONCE per (hierarchy, shift) by Gauss-Jordan on [A | I] on the device and cached; the solve
// inside every V-cycle is then one dense mat-vec, which is bandwidth-trivial and has no sequential
// dependency chain (a triangular solve would serialise thousands of steps inside every cycle).
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

// aug is n x 2n row-major: [A - shift I | I]
__global__ void build_dense_kernel(LevelDev L, double shift, double *__restrict__ aug) {
  const int n = L.nrows * L.ncols;
  const long long total = (long long)n * 2 * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / (2 * n));
    const int c = (int)(idx - (long long)r * 2 * n);
    double val = 0.0;
    if (c >= n) {
      val = (c - n == r) ? 1.0 : 0.0;
    } else {
      const int i = r / L.ncols, j = r - i * L.ncols;
      const int i2 = c / L.ncols, j2 = c - i2 * L.ncols;
      const int di = i2 - i, dj = j2 - j;
      if (di >= -1 && di <= 1 && dj >= -1 && dj <= 1) {
        const int gi = L.row0 + i;
        const double ka = di < 0 ? L.ka_lo[gi] : (di == 0 ? L.ka_di[gi] : L.ka_up[gi]);
        const double kb = dj < 0 ? L.kb_lo[j] : (dj == 0 ? L.kb_di[j] : L.kb_up[j]);
        if (L.five) {
          if (di == 0 && dj == 0) val = ka + kb;
          else if (di == 0) val = kb;
          else if (dj == 0) val = ka;
        } else {
          const double ma = di < 0 ? L.ma_lo[gi] : (di == 0 ? L.ma_di[gi] : L.ma_up[gi]);
          const double mb = dj < 0 ? L.mb_lo[j] : (dj == 0 ? L.mb_di[j] : L.mb_up[j]);
          val = ma * kb + ka * mb;
        }
        if (di == 0 && dj == 0) val -= shift;
      }
    }
    aug[idx] = val;
  }
}

cudaError_t launch_build_dense(const LevelDev &L, double shift, double *aug, cudaStream_t s) {
  const long long total = (long long)L.nrows * L.ncols * 2 * L.nrows * L.ncols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  build_dense_kernel<<<blocks, 256, 0, s>>>(L, shift, aug);
  count_launch();
  return cudaGetLastError();
}

// --- Gauss-Jordan with partial pivoting, one pivot column per step ---------------------------------
// step kernel A (one CTA): find pivot row p >= k maximising |aug[p][k]|; record it.
__global__ void gj_pivot_kernel(int n, int k, const double *__restrict__ aug, int *__restrict__ piv,
                                int *__restrict__ status) {
  __shared__ double sval[1024];
  __shared__ int sidx[1024];
  double best = -1.0;
  int bi = k;
  for (int r = k + threadIdx.x; r < n; r += blockDim.x) {
    const double a = fabs(aug[(size_t)r * 2 * n + k]);
    if (a > best) { best = a; bi = r; }
  }
  sval[threadIdx.x] = best;
  sidx[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double b = sval[threadIdx.x + o];
      const int ib = sidx[threadIdx.x + o];
      if (b > sval[threadIdx.x] || (b == sval[threadIdx.x] && ib < sidx[threadIdx.x])) {
        sval[threadIdx.x] = b;
        sidx[threadIdx.x] = ib;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *piv = sidx[0];
    if (!(sval[0] > 0.0)) *status = 1;  // singular (or NaN)
  }
}

// step kernel B0/B: capture the pivot value (so no thread reads it while another overwrites it), then
// swap rows k and p and scale row k by 1/pivot
__global__ void gj_capture_kernel(int n, int k, const double *__restrict__ aug, const int *__restrict__ piv,
                                  double *__restrict__ pivot_val) {
  *pivot_val = aug[(size_t)(*piv) * 2 * n + k];
}
__global__ void gj_swap_scale_kernel(int n, int k, double *__restrict__ aug, const int *__restrict__ piv,
                                      const double *__restrict__ pivot_val) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * n) return;
  const int p = *piv;
  double *rk = aug + (size_t)k * 2 * n, *rp = aug + (size_t)p * 2 * n;
  const double pivot = *pivot_val;
  const double a = rp[c];
  if (p != k) rp[c] = rk[c];
  rk[c] = a / pivot;
}
// step kernel C: gather multipliers mult[r] = aug[r][k]
__global__ void gj_gather_kernel(int n, int k, const double *__restrict__ aug, double *__restrict__ mult) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) mult[r] = aug[(size_t)r * 2 * n + k];
}
// step kernel D: row_r -= mult[r] * row_k for r != k
__global__ void gj_eliminate_kernel(int n, int k, double *__restrict__ aug, const double *__restrict__ mult) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * n) return;
  const double pk = aug[(size_t)k * 2 * n + c];
  for (int r = blockIdx.y; r < n; r += gridDim.y) {
    if (r == k) continue;
    const double m = mult[r];
    if (m != 0.0) aug[(size_t)r * 2 * n + c] -= m * pk;
  }
}

cudaError_t launch_gauss_jordan(int n, double *aug, int *status, double *mult, cudaStream_t s) {
  // mult: n doubles + 1 double (pivot value) + 1 int slot (pivot row) carved from the tail
  double *pivot_val = mult + n;
  int *piv = reinterpret_cast<int *>(mult + n + 1);
  int pthreads = 32;
  while (pthreads < n && pthreads < 1024) pthreads <<= 1;
  const int cthreads = 128;
  const int cblocks = (2 * n + cthreads - 1) / cthreads;
  int yblocks = n < 64 ? n : 64;
  for (int k = 0; k < n; ++k) {
    gj_pivot_kernel<<<1, pthreads, 0, s>>>(n, k, aug, piv, status);
    gj_capture_kernel<<<1, 1, 0, s>>>(n, k, aug, piv, pivot_val);
    gj_swap_scale_kernel<<<cblocks, cthreads, 0, s>>>(n, k, aug, piv, pivot_val);
    gj_gather_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, k, aug, mult);
    gj_eliminate_kernel<<<dim3(cblocks, yblocks), cthreads, 0, s>>>(n, k, aug, mult);
    count_launch(5);
  }
  return cudaGetLastError();
}

__global__ void extract_inverse_kernel(int n, const double *__restrict__ aug, double *__restrict__ inv) {
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / n), c = (int)(idx - (long long)r * n);
    inv[idx] = aug[(size_t)r * 2 * n + n + c];
  }
}
cudaError_t launch_extract_inverse(int n, const double *aug, double *inv, cudaStream_t s) {
  long long total = (long long)n * n;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  extract_inverse_kernel<<<blocks, 256, 0, s>>>(n, aug, inv);
  count_launch();
  return cudaGetLastError();
}

// y = inv x ; one warp per row, fixed summation order
__global__ void gemv_kernel(int n, const double *__restrict__ inv, const double *__restrict__ x,
                            double *__restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double *row = inv + (size_t)warp * n;
  double acc = 0.0;
  for (int c = lane; c < n; c += 32) acc += row[c] * x[c];
  acc = warp_sum(acc);
  if (lane == 0) y[warp] = acc;
}
cudaError_t launch_gemv(int n, const double *inv, const double *x, double *y, cudaStream_t s) {
  const int threads = 128;
  const int blocks = (n * 32 + threads - 1) / threads;
  gemv_kernel<<<blocks, threads, 0, s>>>(n, inv, x, y);
  count_launch();
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// Banded route to the same dense inverse, for coarsest levels of thousands of unknowns (lowest_level = 64 in 2-D:
// 4096 unknowns, the "7 levels" of a 4096^2 grid).  The 9-point operator has half-bandwidth kl = ncols + 1, so LU with
// partial pivoting touches kl rows x (kl + ku + kl) columns per pivot instead of n x 2n (the Gauss-Jordan above: ~20 000
// launches and 2 TB of traffic at n = 4096).  Window storage as in the CPU oracle / LAPACK gbtrf: row i keeps columns
// i - kl .. i + ku + kl; multipliers are not permuted retroactively, so the solves interleave interchanges and
// eliminations.  Three launches: build the band, factor it (one CTA, the kl + 1 active rows live in shared memory and
// slide down the matrix), then one thread per column of the identity runs the forward and backward substitution on the
// n x n result in place (rows k .. k + kl of all columns are the working set: L2-resident).
// ---------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ double level_entry(const LevelDev &L, double shift, int r, int c) {
  const int i = r / L.ncols, j = r - i * L.ncols;
  const int i2 = c / L.ncols, j2 = c - i2 * L.ncols;
  const int di = i2 - i, dj = j2 - j;
  if (di < -1 || di > 1 || dj < -1 || dj > 1) return 0.0;
  const int gi = L.row0 + i;
  const double ka = di < 0 ? L.ka_lo[gi] : (di == 0 ? L.ka_di[gi] : L.ka_up[gi]);
  const double kb = dj < 0 ? L.kb_lo[j] : (dj == 0 ? L.kb_di[j] : L.kb_up[j]);
  double val;
  if (L.five) {
    val = (di == 0 && dj == 0) ? ka + kb : (di == 0 ? kb : (dj == 0 ? ka : 0.0));
  } else {
    const double ma = di < 0 ? L.ma_lo[gi] : (di == 0 ? L.ma_di[gi] : L.ma_up[gi]);
    const double mb = dj < 0 ? L.mb_lo[j] : (dj == 0 ? L.mb_di[j] : L.mb_up[j]);
    val = ma * kb + ka * mb;
  }
  if (di == 0 && dj == 0) val -= shift;
  return val;
}

// band[i * wp + (c - i + kl)] = (A - shift I)[i][c] for |c - i| <= ku, 0 in the fill columns
__global__ void band_build_kernel(LevelDev L, double shift, int n, int kl, int ku, int wp, double *__restrict__ band) {
  const long long total = (long long)n * wp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / wp), w = (int)(idx - (long long)i * wp);
    const int c = i - kl + w;
    double v = 0.0;
    if (c >= 0 && c < n && c - i <= ku && i - c <= kl) v = level_entry(L, shift, i, c);
    band[idx] = v;
  }
}

// One CTA.  Shared memory: R = kl + 1 row windows (circular: row i lives in slot i % R), each wp doubles.
__global__ void __launch_bounds__(1024)
band_factor_kernel(int n, int kl, int ku, int wp, double *__restrict__ band, int *__restrict__ piv, int *__restrict__ status) {
  extern __shared__ double S[];
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_p;
  const int R = kl + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < min(R, n) * wp; idx += nt) S[idx] = band[idx];
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const int rhi = min(k + kl, n - 1), chi = min(k + ku + kl, n - 1);
    const int nr = rhi - k + 1;
    // 1. pivot: largest |A[r][k]|, r = k .. rhi (first wins ties, as in the oracle)
    {
      double best = -1.0;
      int bi = k;
      for (int t = tid; t < nr; t += nt) {
        const double a = fabs(S[((k + t) % R) * wp + (kl - t)]);
        if (a > best) { best = a; bi = k + t; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
      __syncthreads();
      if (tid < 32) {
        const int nw = (nt + 31) >> 5;
        best = tid < nw ? s_val[tid] : -1.0;
        bi = tid < nw ? s_idx[tid] : k;
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (tid == 0) {
          s_p = bi;
          piv[k] = bi;
          if (!(best > 0.0)) *status = 1;  // singular (or NaN)
        }
      }
      __syncthreads();
    }
    const int p = s_p;
    double *rowk = S + (k % R) * wp;  // element (k, c) at rowk[c - k + kl]
    // 2. interchange rows k and p in columns k .. chi
    if (p != k) {
      double *rowp = S + (p % R) * wp;
      for (int c = k + tid; c <= chi; c += nt) {
        const double a = rowk[c - k + kl], b = rowp[c - p + kl];
        rowk[c - k + kl] = b;
        rowp[c - p + kl] = a;
      }
      __syncthreads();
    }
    // 3. multipliers, 4. elimination: one warp per row r, lanes over the columns
    const double pv = rowk[kl];
    const double ipv = (pv != 0.0) ? 1.0 / pv : 0.0;
    (void)ipv;
    for (int r = k + 1 + (tid >> 5); r <= rhi; r += (nt >> 5)) {
      double *rowr = S + (r % R) * wp;
      const double m = (pv != 0.0) ? rowr[k - r + kl] / pv : 0.0;
      __syncwarp();
      if ((tid & 31) == 0) rowr[k - r + kl] = m;
      if (m != 0.0)
        for (int c = k + 1 + (tid & 31); c <= chi; c += 32) rowr[c - r + kl] -= m * rowk[c - k + kl];
    }
    __syncthreads();
    // 5. row k is final: back to global; its slot takes row k + R
    for (int w = tid; w < wp; w += nt) band[(size_t)k * wp + w] = rowk[w];
    if (k + R < n)
      for (int w = tid; w < wp; w += nt) rowk[w] = band[(size_t)(k + R) * wp + w];
    __syncthreads();
  }
}

// One thread per column j of the identity: X[:, j] = A^-1 e_j, X row-major n x n (X[k * n + j]).
__global__ void band_inverse_solve_kernel(int n, int kl, int ku, int wp, const double *__restrict__ band,
                                          const int *__restrict__ piv, double *__restrict__ X) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  for (int k = 0; k < n; ++k) X[(size_t)k * n + j] = (k == j) ? 1.0 : 0.0;
  // forward: everything above row j - kl is still zero
  for (int k = max(0, j - kl); k < n; ++k) {
    const int p = piv[k];
    double yk = X[(size_t)k * n + j];
    if (p != k) {
      const double yp = X[(size_t)p * n + j];
      X[(size_t)p * n + j] = yk;
      X[(size_t)k * n + j] = yp;
      yk = yp;
    }
    if (yk != 0.0) {
      const int rhi = min(k + kl, n - 1);
      for (int r = k + 1; r <= rhi; ++r) X[(size_t)r * n + j] -= band[(size_t)r * wp + (k - r + kl)] * yk;
    }
  }
  // backward
  for (int k = n - 1; k >= 0; --k) {
    const int chi = min(k + ku + kl, n - 1);
    const double *u = band + (size_t)k * wp + kl;  // u[c - k] = U[k][c]
    double t0 = X[(size_t)k * n + j], t1 = 0.0, t2 = 0.0, t3 = 0.0;
    int c = k + 1;
    for (; c + 3 <= chi; c += 4) {
      t0 -= u[c - k] * X[(size_t)c * n + j];
      t1 -= u[c + 1 - k] * X[(size_t)(c + 1) * n + j];
      t2 -= u[c + 2 - k] * X[(size_t)(c + 2) * n + j];
      t3 -= u[c + 3 - k] * X[(size_t)(c + 3) * n + j];
    }
    for (; c <= chi; ++c) t0 -= u[c - k] * X[(size_t)c * n + j];
    X[(size_t)k * n + j] = ((t0 + t1) + (t2 + t3)) / u[0];
  }
}

}  // namespace

size_t band_workspace_doubles(int n, int kl, int ku) { return (size_t)n * (2 * kl + ku + 1) + (size_t)(n + 1) / 2 + 2; }

// inv (n x n, row-major) = (A_L - shift I)^-1 through the banded LU.  work: band_workspace_doubles(n, kl, ku) doubles.
cudaError_t launch_band_inverse2d(const LevelDev &L, double shift, double *inv, int *status, double *work, cudaStream_t s) {
  const int n = L.nrows * L.ncols;
  const int kl = (L.nrows > 1) ? min(L.ncols + 1, n - 1) : min(1, n - 1), ku = kl;
  const int wp = 2 * kl + ku + 1;
  double *band = work;
  int *piv = reinterpret_cast<int *>(work + (size_t)n * wp);
  const long long total = (long long)n * wp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  band_build_kernel<<<blocks, 256, 0, s>>>(L, shift, n, kl, ku, wp, band);
  const size_t smem = sizeof(double) * (size_t)(kl + 1) * wp;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(band_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  band_factor_kernel<<<1, 1024, smem, s>>>(n, kl, ku, wp, band, piv, status);
  band_inverse_solve_kernel<<<(n + 63) / 64, 64, 0, s>>>(n, kl, ku, wp, band, piv, inv);
  count_launch(3);
  return cudaGetLastError();
}

bool band_inverse_fits(const LevelDev &L) {
  const int n = L.nrows * L.ncols;
  if (n < 4) return false;
  const int kl = (L.nrows > 1) ? min(L.ncols + 1, n - 1) : 1;
  return sizeof(double) * (size_t)(kl + 1) * (3 * kl + 1) <= 200 * 1024;
}

}  // namespace mgcmt
