// reduce.cu -- deterministic reductions and the vector updates around them.
//
// Replaces the numpy dots/norms of the reference's drivers and MGCMTProcessor:
//   Rayleigh quotient  v^T (H v) / v^T v          e.g. 2DPotGS.py:103, MGCMTSolver.py:19
//   normalisation      w / ||w||                   e.g. 2DPotGS.py:96, MGCMTProcessor.py:52-63
//   projections        (<v,u>/<u,u>) u             MGCMTProcessor.py:10-20
// Every reduction is a fixed two-stage tree (per-thread grid-stride partial -> warp shuffle -> block ->
// ordered sum of the block partials), so a result depends only on n, never on scheduling; repeated runs
// and different GPU counts give bit-identical scalars ("Reductions").
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

constexpr int kRedThreads = 256;

// partials[m * gridDim.x + b] = sum over this block's elements of X_m[i] * y[i],  X_m = x0 + m*stride
// VEC: all operands 16-byte aligned (base pointers and stride), so pairs can be loaded as double2;
// the summation order is the same either way.
template <int M, bool VEC>
__global__ void __launch_bounds__(kRedThreads)
multidot_partial_kernel(long long n, const double *__restrict__ x0, long long stride,
                        const double *__restrict__ y, double *__restrict__ partials) {
  double acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0.0;
  const long long n2 = n >> 1;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += step) {
    double2 yy;
    if (VEC) yy = reinterpret_cast<const double2 *>(y)[i];
    else yy = make_double2(y[2 * i], y[2 * i + 1]);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      double2 xx;
      if (VEC) xx = reinterpret_cast<const double2 *>(x0 + m * stride)[i];
      else xx = make_double2(x0[m * stride + 2 * i], x0[m * stride + 2 * i + 1]);
      acc[m] += xx.x * yy.x;
      acc[m] += xx.y * yy.y;
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] += x0[m * stride + n - 1] * y[n - 1];
  }
  __shared__ double sm[M][kRedThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const double s = warp_sum(acc[m]);
    if (lane == 0) sm[m][w] = s;
  }
  __syncthreads();
  if (threadIdx.x < M) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kRedThreads / 32; ++k) s += sm[threadIdx.x][k];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

// out[m] = ordered tree sum of partials[m*B .. m*B+B)
__global__ void __launch_bounds__(256) finish_kernel(int B, const double *__restrict__ partials,
                                                     double *__restrict__ out) {
  __shared__ double sm[256];
  const int m = blockIdx.x;
  double s = 0.0;
  for (int b = threadIdx.x; b < B; b += 256) s += partials[(size_t)m * B + b];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[m] = sm[0];
}

static int blocks_for(long long n) {
  long long b = (n / 2 + kRedThreads - 1) / kRedThreads;
  if (b < 1) b = 1;
  if (b > kReduceBlocks) b = kReduceBlocks;
  return (int)b;
}

cudaError_t launch_multidot(long long n, int M, const double *x0, long long stride, const double *y,
                            double *partials, double *out, cudaStream_t s) {
  const int B = blocks_for(n);
  const bool vec = ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                   (M == 1 || (stride & 1) == 0);
  switch (M) {
#define CASE(MM)                                                                                 \
  case MM:                                                                                       \
    if (vec) multidot_partial_kernel<MM, true><<<B, kRedThreads, 0, s>>>(n, x0, stride, y, partials);  \
    else multidot_partial_kernel<MM, false><<<B, kRedThreads, 0, s>>>(n, x0, stride, y, partials);     \
    break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
    CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    default:
      return cudaErrorInvalidValue;
  }
  finish_kernel<<<M, 256, 0, s>>>(B, partials, out);
  count_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_finish(int M, int B, const double *partials, double *out, cudaStream_t s) {
  finish_kernel<<<M, 256, 0, s>>>(B, partials, out);
  count_launch();
  return cudaGetLastError();
}

// ---- fused modified Gram-Schmidt (MGCMTProcessor.py:44-50) -----------------------------------------
// pass 1 of column i:  q = w_i / sqrt(*sumsq) (written in place);  partial sums of <q,q> (slot 0) and of
//                      <w_j, q> for the m later columns j (slots 1..m)
template <int M>
__global__ void __launch_bounds__(kRedThreads)
mgs_scale_dots_kernel(long long n, double *__restrict__ wi, const double *__restrict__ sumsq,
                      const double *__restrict__ wj0, long long stride, double *__restrict__ partials) {
  const double nrm = sqrt(*sumsq);
  double acc[M + 1];
#pragma unroll
  for (int m = 0; m <= M; ++m) acc[m] = 0.0;
  const long long n2 = n >> 1;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += step) {
    double2 q = reinterpret_cast<const double2 *>(wi)[i];
    q.x = q.x / nrm;
    q.y = q.y / nrm;
    reinterpret_cast<double2 *>(wi)[i] = q;
    acc[0] += q.x * q.x;
    acc[0] += q.y * q.y;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const double2 w = reinterpret_cast<const double2 *>(wj0 + m * stride)[i];
      acc[m + 1] += w.x * q.x;
      acc[m + 1] += w.y * q.y;
    }
  }
  __shared__ double sm[M + 1][kRedThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m <= M; ++m) {
    const double s = warp_sum(acc[m]);
    if (lane == 0) sm[m][w] = s;
  }
  __syncthreads();
  if (threadIdx.x <= M) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kRedThreads / 32; ++k) s += sm[threadIdx.x][k];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

// pass 2 of column i:  w_j -= (dots[1+j] / dots[0]) q  for the m later columns;  partial sums of the new
//                      ||w_{i+1}||^2 (the next column to be normalised)
template <int M>
__global__ void __launch_bounds__(kRedThreads)
mgs_update_kernel(long long n, const double *__restrict__ qi, const double *__restrict__ dots,
                  double *__restrict__ wj0, long long stride, double *__restrict__ partials) {
  double c[M];
  const double qq = dots[0];
#pragma unroll
  for (int m = 0; m < M; ++m) c[m] = dots[1 + m] / qq;
  double acc = 0.0;
  const long long n2 = n >> 1;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += step) {
    const double2 q = reinterpret_cast<const double2 *>(qi)[i];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      double2 w = reinterpret_cast<double2 *>(wj0 + m * stride)[i];
      w.x -= c[m] * q.x;
      w.y -= c[m] * q.y;
      reinterpret_cast<double2 *>(wj0 + m * stride)[i] = w;
      if (m == 0) {
        acc += w.x * w.x;
        acc += w.y * w.y;
      }
    }
  }
  __shared__ double sm[kRedThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double s = warp_sum(acc);
  if (lane == 0) sm[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kRedThreads / 32; ++k) t += sm[k];
    partials[blockIdx.x] = t;
  }
}

// both need n even and 16-byte aligned columns (checked by the caller, which otherwise uses the un-fused path)
cudaError_t launch_mgs_scale_dots(long long n, int m, double *wi, const double *sumsq, const double *wj0,
                                  long long stride, double *partials, double *out, cudaStream_t s) {
  const int B = blocks_for(n);
  switch (m) {
#define CASE(MM) case MM: mgs_scale_dots_kernel<MM><<<B, kRedThreads, 0, s>>>(n, wi, sumsq, wj0, stride, partials); break;
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
#undef CASE
    default: return cudaErrorInvalidValue;
  }
  finish_kernel<<<m + 1, 256, 0, s>>>(B, partials, out);
  count_launch(2);
  return cudaGetLastError();
}
cudaError_t launch_mgs_update(long long n, int m, const double *qi, const double *dots, double *wj0, long long stride,
                              double *partials, double *out_sumsq, cudaStream_t s) {
  const int B = blocks_for(n);
  switch (m) {
#define CASE(MM) case MM: mgs_update_kernel<MM><<<B, kRedThreads, 0, s>>>(n, qi, dots, wj0, stride, partials); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
#undef CASE
    default: return cudaErrorInvalidValue;
  }
  finish_kernel<<<1, 256, 0, s>>>(B, partials, out_sumsq);
  count_launch(2);
  return cudaGetLastError();
}

// ---- Gram-matrix (Cholesky-QR) orthonormalisation ----------------------------------------------------------
// Q = W R^-1 with W^T W = R^T R is the same Q as Gram-Schmidt produces (QR with positive diagonal is unique);
// it needs one pass for the Gram matrix and one for the triangular combination: 3k vector passes instead
// of ~k^2 + 3k for column-by-column MGS.  Rounding differs from MGS by O(cond(W)^2 eps): meant for the nearly
// orthonormal blocks of the eigen-iteration (tests compare it with MGS at 1e-12).
template <int K>
__global__ void __launch_bounds__(kRedThreads)
gram_partial_kernel(long long n, const double *__restrict__ w0, long long stride, double *__restrict__ partials) {
  constexpr int NP = K * (K + 1) / 2;
  double acc[NP];
#pragma unroll
  for (int m = 0; m < NP; ++m) acc[m] = 0.0;
  const long long n2 = n >> 1;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += step) {
    double2 w[K];
#pragma unroll
    for (int a = 0; a < K; ++a) w[a] = reinterpret_cast<const double2 *>(w0 + a * stride)[i];
    int m = 0;
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
      for (int b = a; b < K; ++b) {
        acc[m] += w[a].x * w[b].x;
        acc[m] += w[a].y * w[b].y;
        ++m;
      }
  }
  __shared__ double sm[NP][kRedThreads / 32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < NP; ++m) {
    const double s = warp_sum(acc[m]);
    if (lane == 0) sm[m][wp] = s;
  }
  __syncthreads();
  if (threadIdx.x < NP) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < kRedThreads / 32; ++q) s += sm[threadIdx.x][q];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

// packed upper Gram matrix g (row-major over a <= b) -> rinv (K x K row-major, upper): inverse of the Cholesky factor
__global__ void chol_inverse_kernel(int K, const double *__restrict__ g, double *__restrict__ rinv, int *__restrict__ status) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double G[8][8], R[8][8], X[8][8];
  int m = 0;
  for (int a = 0; a < K; ++a)
    for (int b = a; b < K; ++b) { G[a][b] = G[b][a] = g[m++]; }
  for (int a = 0; a < K; ++a)
    for (int b = 0; b < K; ++b) { R[a][b] = 0.0; X[a][b] = 0.0; }
  for (int j = 0; j < K; ++j) {  // G = R^T R, R upper
    double d = G[j][j];
    for (int q = 0; q < j; ++q) d -= R[q][j] * R[q][j];
    if (!(d > 0.0)) { *status = 1; d = 1.0; }
    R[j][j] = sqrt(d);
    for (int b = j + 1; b < K; ++b) {
      double v = G[j][b];
      for (int q = 0; q < j; ++q) v -= R[q][j] * R[q][b];
      R[j][b] = v / R[j][j];
    }
  }
  for (int j = 0; j < K; ++j) {  // X = R^-1 (upper), column by column
    X[j][j] = 1.0 / R[j][j];
    for (int a = j - 1; a >= 0; --a) {
      double v = 0.0;
      for (int q = a + 1; q <= j; ++q) v -= R[a][q] * X[q][j];
      X[a][j] = v / R[a][a];
    }
  }
  for (int a = 0; a < K; ++a)
    for (int b = 0; b < K; ++b) rinv[a * K + b] = X[a][b];
}

// q_j = sum_{i <= j} rinv[i][j] w_i, in place (descending j so every w_i is still the input when it is read)
template <int K>
__global__ void __launch_bounds__(kRedThreads)
cholqr_apply_kernel(long long n, double *__restrict__ w0, long long stride, const double *__restrict__ rinv) {
  double c[K][K];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < K; ++b) c[a][b] = rinv[a * K + b];
  const long long n2 = n >> 1;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += step) {
    double2 w[K];
#pragma unroll
    for (int a = 0; a < K; ++a) w[a] = reinterpret_cast<const double2 *>(w0 + a * stride)[i];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double2 q = make_double2(0.0, 0.0);
#pragma unroll
      for (int a = 0; a <= j; ++a) {
        q.x += c[a][j] * w[a].x;
        q.y += c[a][j] * w[a].y;
      }
      reinterpret_cast<double2 *>(w0 + j * stride)[i] = q;
    }
  }
}

cudaError_t launch_gram(long long n, int k, const double *w0, long long stride, double *partials, double *out, cudaStream_t s) {
  const int B = blocks_for(n);
  switch (k) {
#define CASE(KK) case KK: gram_partial_kernel<KK><<<B, kRedThreads, 0, s>>>(n, w0, stride, partials); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6)
#undef CASE
    default: return cudaErrorInvalidValue;
  }
  finish_kernel<<<k * (k + 1) / 2, 256, 0, s>>>(B, partials, out);
  count_launch(2);
  return cudaGetLastError();
}
cudaError_t launch_chol_inverse(int k, const double *g, double *rinv, int *status, cudaStream_t s) {
  chol_inverse_kernel<<<1, 32, 0, s>>>(k, g, rinv, status);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_cholqr_apply(long long n, int k, double *w0, long long stride, const double *rinv, cudaStream_t s) {
  long long blk = (n / 2 + kRedThreads - 1) / kRedThreads;
  if (blk > 148 * 8) blk = 148 * 8;
  if (blk < 1) blk = 1;
  switch (k) {
#define CASE(KK) case KK: cholqr_apply_kernel<KK><<<(int)blk, kRedThreads, 0, s>>>(n, w0, stride, rinv); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6)
#undef CASE
    default: return cudaErrorInvalidValue;
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_dot(long long n, const double *x, const double *y, double *partials, double *out,
                       cudaStream_t s) {
  return launch_multidot(n, 1, x, 0, y, partials, out, s);
}

// x[i] = x[i] / sqrt(*sumsq)     (division, as numpy does)
__global__ void scale_by_inv_norm_kernel(long long n, const double *__restrict__ x, const double *__restrict__ sumsq,
                                         double *__restrict__ y) {
  const double nrm = sqrt(*sumsq);
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) y[i] = x[i] / nrm;
}
cudaError_t launch_scale_to(long long n, const double *x, const double *sumsq, double *y, cudaStream_t s) {
  long long b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  scale_by_inv_norm_kernel<<<(int)b, 256, 0, s>>>(n, x, sumsq, y);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_scale_by_inv_norm(long long n, double *x, const double *sumsq, cudaStream_t s) {
  return launch_scale_to(n, x, sumsq, x, s);
}

// y += sign * (alpha / denom) * x      alpha, denom on the device (denom may be null => 1)
__global__ void axpy_dev_kernel(long long n, const double *__restrict__ alpha, const double *__restrict__ denom,
                                double sign, const double *__restrict__ x, double *__restrict__ y) {
  double a = *alpha;
  if (denom) a = a / *denom;
  a *= sign;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) y[i] += a * x[i];
}
cudaError_t launch_axpy_dev(long long n, const double *alpha, const double *denom, double sign,
                            const double *x, double *y, cudaStream_t s) {
  long long b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  axpy_dev_kernel<<<(int)b, 256, 0, s>>>(n, alpha, denom, sign, x, y);
  count_launch();
  return cudaGetLastError();
}

__global__ void rq_unshift_kernel(double *out2, double shift) { out2[0] += shift * out2[1]; }
cudaError_t launch_rq_unshift(double *out2, double shift, cudaStream_t s) {
  rq_unshift_kernel<<<1, 1, 0, s>>>(out2, shift);
  count_launch();
  return cudaGetLastError();
}

// out = a x + b y   (host scalars)
__global__ void axpby_kernel(long long n, double a, const double *__restrict__ x, double b, const double *__restrict__ y,
                             double *__restrict__ out) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) out[i] = a * x[i] + b * y[i];
}
cudaError_t launch_axpby(long long n, double a, const double *x, double b, const double *y, double *out, cudaStream_t s) {
  long long blk = (n + 255) / 256;
  if (blk > 148 * 16) blk = 148 * 16;
  if (blk < 1) blk = 1;
  axpby_kernel<<<(int)blk, 256, 0, s>>>(n, a, x, b, y, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mgcmt
