// stream_pattern.cu -- which per-warp access pattern lets a row-streaming kernel reach HBM speed on B200?
// Every warp owns a 128-column strip and walks down a chunk of rows, out = v + f, rows prefetched with cp.async
// (ring of 4).  LOAD 0: lane l copies the 16-byte granules 2l and 2l+1 of the row (two instructions, each touching
// every other granule = half of every 32-byte sector: the pattern of fused.cu).  LOAD 1: lane l copies granules l and
// 32+l (each instruction 512 contiguous bytes), XOR-swizzled in shared memory, __syncwarp, then reads its 4 columns.
// STORE 0: two 16-byte stores per lane (half sectors per instruction).  STORE 1: one 32-byte store per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void cpa16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void st32(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void st16(double *p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

template <int LOAD, int STORE, int NARR>
__global__ void __launch_bounds__(128) k(const double *__restrict__ v, const double *__restrict__ f, double *__restrict__ out,
                                         int ncols, int nrows, int rpc) {
  __shared__ __align__(128) double2 ring[2][4][4][64];  // [array][slot][warp][granule]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strip = blockIdx.x * 4 + warp;
  const int col0 = strip * 128;
  if (col0 >= ncols) return;
  const int r0 = blockIdx.y * rpc, r1 = min(r0 + rpc, nrows);
  auto issue = [&](int t) {
    if (t < r1) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int G = LOAD ? (32 * g + lane) : (2 * lane + g);
        const int P = LOAD ? (G ^ ((G >> 3) & 1)) : G;
        cpa16(&ring[0][t & 3][warp][P], v + (size_t)t * ncols + col0 + 2 * G);
        if (NARR > 1) cpa16(&ring[1][t & 3][warp][P], f + (size_t)t * ncols + col0 + 2 * G);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int d = 0; d < 4; ++d) issue(r0 + d);
  for (int t = r0; t < r1; ++t) {
    asm volatile("cp.async.wait_group 3;" ::: "memory");
    if (LOAD) __syncwarp();
    double x[4];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int G = 2 * lane + g;
      const int P = LOAD ? (G ^ ((G >> 3) & 1)) : G;
      double2 a = ring[0][t & 3][warp][P];
      if (NARR > 1) { double2 b = ring[1][t & 3][warp][P]; a.x += b.x; a.y += b.y; }
      x[2 * g] = a.x; x[2 * g + 1] = a.y;
    }
    if (LOAD) __syncwarp();
    issue(t + 4);
    double *dst = out + (size_t)t * ncols + col0 + 4 * lane;
    if (STORE) st32(dst, x[0], x[1], x[2], x[3]);
    else { st16(dst, x[0], x[1]); st16(dst + 2, x[2], x[3]); }
  }
}

template <int LOAD, int STORE, int NARR>
void run(const char *name, const double *v, const double *f, double *out, int N, char *flush, size_t fb) {
  dim3 grid(N / 512, N / 128);
  float best = 1e9f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 8; ++it) {
    cudaMemsetAsync(flush, it, fb);
    cudaEventRecord(e0);
    k<LOAD, STORE, NARR><<<grid, 128>>>(v, f, out, N, N, 128);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 2 && ms < best) best = ms;
  }
  const double bytes = (double)N * N * 8 * (NARR + 1);
  printf("%-34s %7.1f us  %6.0f GB/s  (%s)\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int N = 4096;
  double *v, *f, *out; char *flush; const size_t fb = 256u << 20;
  cudaMalloc(&v, sizeof(double) * N * N); cudaMalloc(&f, sizeof(double) * N * N); cudaMalloc(&out, sizeof(double) * N * N);
  cudaMalloc(&flush, fb);
  cudaMemset(v, 0, sizeof(double) * N * N); cudaMemset(f, 0, sizeof(double) * N * N);
  run<0, 0, 2>("2in: half-sector ld, 2x16B st", v, f, out, N, flush, fb);
  run<1, 0, 2>("2in: contiguous ld,  2x16B st", v, f, out, N, flush, fb);
  run<0, 1, 2>("2in: half-sector ld, 32B st", v, f, out, N, flush, fb);
  run<1, 1, 2>("2in: contiguous ld,  32B st", v, f, out, N, flush, fb);
  run<0, 0, 1>("1in: half-sector ld, 2x16B st", v, f, out, N, flush, fb);
  run<1, 1, 1>("1in: contiguous ld,  32B st", v, f, out, N, flush, fb);
  cudaMemcpy(out, v, sizeof(double) * N * N, cudaMemcpyDeviceToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0); cudaMemcpyAsync(out, v, sizeof(double) * N * N, cudaMemcpyDeviceToDevice); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  printf("cudaMemcpy D2D 134 MB: %.1f us %.0f GB/s\n", best * 1e3, 2.0 * N * N * 8 / (best * 1e-3) / 1e9);
  return 0;
}
