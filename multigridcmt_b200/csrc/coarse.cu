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

}  // namespace mgcmt
