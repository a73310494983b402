// rq.cu -- Rayleigh-quotient minimisation (MGCMTSolver.rqmin, MGCMTSolver.py:17-57) resident on the device.
//
// The reference's step: search direction p = -g + (g^T M g / g_old^T M g_old) p (:34-36), the 2 x 2 pencil
// R = [x p]^T A [x p], RM = [x p]^T M [x p] filled with 8 mat-vecs and 8 dots (:38-46), scipy.linalg.eig(R, b=RM) (:48),
// step delta = z[1] / z[0] of the eigenvector of the smallest eigenvalue (:49-51), new rho and gradient
// g = 2 (A x - rho M x) (:52-54).  Here: A p and M p once per step (A x, M x follow by linearity: x <- x + delta p), the
// 8 pencil entries in ONE pass over the six vectors, the 2 x 2 generalised eigenproblem in closed form in a one-thread
// kernel, the update fused with the sums of the new Rayleigh quotient, the gradient fused with g^T g -- and no scalar
// ever visits the host.  All sums are the fixed two-stage trees of reduce.cu.
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

namespace {

constexpr int kRqThreads = 256;

template <int M>
__device__ __forceinline__ void block_partials(double (&acc)[M], double *__restrict__ partials) {
  __shared__ double sm[M][kRqThreads / 32];
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
    for (int k = 0; k < kRqThreads / 32; ++k) s += sm[threadIdx.x][k];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

// sums 0..7: (x,Ax) (x,Ap) (p,Ax) (p,Ap) (x,Mx) (x,Mp) (p,Mx) (p,Mp)      (MGCMTSolver.py:38-46)
__global__ void __launch_bounds__(kRqThreads)
rq_pencil_kernel(long long n, const double *__restrict__ x, const double *__restrict__ p, const double *__restrict__ Ax,
                 const double *__restrict__ Ap, const double *__restrict__ Mx, const double *__restrict__ Mp,
                 double *__restrict__ partials) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) {
    const double xi = x[i], pi = p[i], ax = Ax[i], ap = Ap[i], mx = Mx[i], mp = Mp[i];
    acc[0] += xi * ax; acc[1] += xi * ap; acc[2] += pi * ax; acc[3] += pi * ap;
    acc[4] += xi * mx; acc[5] += xi * mp; acc[6] += pi * mx; acc[7] += pi * mp;
  }
  block_partials<8>(acc, partials);
}

// smallest eigenvalue of R z = lambda RM z (2 x 2, real) and delta = z[1] / z[0]; s = the 8 sums above.
// out[0] = delta, out[1] = lambda.  A complex pair (cannot happen for a symmetric-definite pencil up to rounding) is
// treated as a double root.
__global__ void rq_eig2_kernel(const double *__restrict__ s, double *__restrict__ out) {
  const double r00 = s[0], r01 = s[1], r10 = s[2], r11 = s[3], b00 = s[4], b01 = s[5], b10 = s[6], b11 = s[7];
  const double a = b00 * b11 - b01 * b10;
  const double bq = -(r00 * b11 + r11 * b00 - r01 * b10 - r10 * b01);
  const double c = r00 * r11 - r01 * r10;
  double lam;
  if (a == 0.0) {
    lam = (bq != 0.0) ? -c / bq : 0.0;
  } else {
    double disc = bq * bq - 4.0 * a * c;
    if (disc < 0.0) disc = 0.0;
    const double q = -0.5 * (bq + (bq >= 0.0 ? sqrt(disc) : -sqrt(disc)));
    const double l1 = q / a, l2 = (q != 0.0) ? c / q : l1;
    lam = l1 < l2 ? l1 : l2;
  }
  const double m00 = r00 - lam * b00, m01 = r01 - lam * b01, m10 = r10 - lam * b10, m11 = r11 - lam * b11;
  double delta;
  if (fabs(m01) >= fabs(m11)) delta = (m01 != 0.0) ? -m00 / m01 : 0.0;
  else delta = -m10 / m11;
  out[0] = delta;
  out[1] = lam;
}

// x += delta p, Ax += delta Ap, Mx += delta Mp (mass == identity: Mx is x, Mp is p, not touched twice);
// partial sums of x^T A x and x^T M x of the new x
template <bool MASS>
__global__ void __launch_bounds__(kRqThreads)
rq_update_kernel(long long n, const double *__restrict__ dl, double *__restrict__ x, const double *__restrict__ p,
                 double *__restrict__ Ax, const double *__restrict__ Ap, double *__restrict__ Mx,
                 const double *__restrict__ Mp, double *__restrict__ partials) {
  const double delta = dl[0];
  double acc[2] = {0, 0};
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) {
    const double xn = x[i] + delta * p[i];
    const double an = Ax[i] + delta * Ap[i];
    double mn = xn;
    if (MASS) { mn = Mx[i] + delta * Mp[i]; Mx[i] = mn; }
    x[i] = xn;
    Ax[i] = an;
    acc[0] += xn * an;
    acc[1] += xn * mn;
  }
  block_partials<2>(acc, partials);
}

// partial sums of x^T Ax and x^T Mx (start of rqmin)
__global__ void __launch_bounds__(kRqThreads)
rq_sums_kernel(long long n, const double *__restrict__ x, const double *__restrict__ Ax, const double *__restrict__ Mx,
               double *__restrict__ partials) {
  double acc[2] = {0, 0};
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) {
    acc[0] += x[i] * Ax[i];
    acc[1] += x[i] * Mx[i];
  }
  block_partials<2>(acc, partials);
}

// g = 2 (Ax - rho Mx), rho = rq[0] / rq[1]; partial sums of g^T g (the g^T M g of a unit mass matrix)
__global__ void __launch_bounds__(kRqThreads)
rq_grad_kernel(long long n, const double *__restrict__ rq, const double *__restrict__ Ax, const double *__restrict__ Mx,
               double *__restrict__ g, double *__restrict__ partials) {
  const double rho = rq[0] / rq[1];
  double acc[1] = {0};
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) {
    const double gi = 2.0 * (Ax[i] - rho * Mx[i]);
    g[i] = gi;
    acc[0] += gi * gi;
  }
  block_partials<1>(acc, partials);
}

// p = -g (first) or -g + (gmg[0] / gmg[1]) p
__global__ void __launch_bounds__(kRqThreads)
rq_dir_kernel(long long n, int first, const double *__restrict__ gmg, const double *__restrict__ g, double *__restrict__ p) {
  const double beta = first ? 0.0 : gmg[0] / gmg[1];
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += step) p[i] = first ? -g[i] : -g[i] + beta * p[i];
}

__global__ void rq_copy2_kernel(const double *__restrict__ src, double *__restrict__ dst) { dst[0] = src[0]; }

int rq_blocks(long long n) {
  long long b = (n + kRqThreads - 1) / kRqThreads;
  if (b < 1) b = 1;
  if (b > kReduceBlocks) b = kReduceBlocks;
  return (int)b;
}

}  // namespace

// scal layout (device, >= 32 doubles): [0..7] pencil sums, [8] delta, [9] lambda, [10] g^T M g, [11] previous g^T M g,
// [12..13] x^T A x, x^T M x.  partials: 8 * kReduceBlocks doubles.
cudaError_t launch_rq_pencil(long long n, const double *x, const double *p, const double *Ax, const double *Ap,
                             const double *Mx, const double *Mp, double *partials, double *scal, cudaStream_t s) {
  const int B = rq_blocks(n);
  rq_pencil_kernel<<<B, kRqThreads, 0, s>>>(n, x, p, Ax, Ap, Mx, Mp, partials);
  cudaError_t e = launch_finish(8, B, partials, scal, s);
  if (e != cudaSuccess) return e;
  rq_eig2_kernel<<<1, 1, 0, s>>>(scal, scal + 8);
  count_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_rq_update(long long n, bool mass, double *x, const double *p, double *Ax, const double *Ap, double *Mx,
                             const double *Mp, double *partials, double *scal, cudaStream_t s) {
  const int B = rq_blocks(n);
  if (mass) rq_update_kernel<true><<<B, kRqThreads, 0, s>>>(n, scal + 8, x, p, Ax, Ap, Mx, Mp, partials);
  else rq_update_kernel<false><<<B, kRqThreads, 0, s>>>(n, scal + 8, x, p, Ax, Ap, Mx, Mp, partials);
  count_launch();
  return launch_finish(2, B, partials, scal + 12, s);
}

cudaError_t launch_rq_sums(long long n, const double *x, const double *Ax, const double *Mx, double *partials, double *scal,
                           cudaStream_t s) {
  const int B = rq_blocks(n);
  rq_sums_kernel<<<B, kRqThreads, 0, s>>>(n, x, Ax, Mx, partials);
  count_launch();
  return launch_finish(2, B, partials, scal + 12, s);
}

// g = 2 (Ax - rho Mx); scal[11] <- scal[10]; scal[10] <- g^T g when the mass matrix is the identity (otherwise the
// caller applies M to g and takes the dot)
cudaError_t launch_rq_grad(long long n, bool mass, const double *Ax, const double *Mx, double *g, double *partials, double *scal,
                           cudaStream_t s) {
  const int B = rq_blocks(n);
  rq_copy2_kernel<<<1, 1, 0, s>>>(scal + 10, scal + 11);
  rq_grad_kernel<<<B, kRqThreads, 0, s>>>(n, scal + 12, Ax, Mx, g, partials);
  count_launch(2);
  if (!mass) return launch_finish(1, B, partials, scal + 10, s);
  return cudaGetLastError();
}

cudaError_t launch_rq_dir(long long n, bool first, const double *scal, const double *g, double *p, cudaStream_t s) {
  const int B = rq_blocks(n);
  rq_dir_kernel<<<B, kRqThreads, 0, s>>>(n, first ? 1 : 0, scal + 10, g, p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mgcmt
