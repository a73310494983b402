// Microbenchmark: achievable DFMA issue rate per SM sub-partition on this GPU, for 1/2/4/8 warps per SMSP and 4/8
// independent chains per thread.  Used to interpret sm__pipe_fp64_cycles_active of the fused legs.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void dfma(double *out, int iters, double b, double c) {
  double a[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) a[j] = threadIdx.x * 1e-3 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CH; ++j) a[j] = fma(a[j], b, c);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < CH; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int warps_per_smsp) {
  const int threads = 32 * 4 * warps_per_smsp;  // one CTA per SM
  const int blocks = 148, iters = 20000;
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma<CH><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  dfma<CH><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr_per_smsp = (double)iters * CH * warps_per_smsp;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("chains %d warps/SMSP %d: %.3f ms, %.2f cycles per DFMA warp-instruction per SMSP (at %d MHz nominal), %.1f TFLOP/s\n", CH,
         warps_per_smsp, ms, cycles / instr_per_smsp, clk / 1000, 2.0 * 32 * instr_per_smsp * 4 * 148 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() {
  for (int w : {1, 2, 4, 8}) { run<4>(w); run<8>(w); }
  return 0;
}
