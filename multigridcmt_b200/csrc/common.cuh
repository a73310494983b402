// common.cuh -- shared device-side types for libmgcmt_b200 (sm_100a, fp64, HBM-bound path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mgcmt {

// One grid level as the kernels see it.  Operator:  A = Ma (x) Kb + Ka (x) Mb  (tridiagonal factors;
// row factors indexed by GLOBAL row, column factors by column), see include/mgcmt_b200.h.
struct LevelDev {
  int nrows;       // rows held locally
  int ncols;       // columns (full width)
  int row0;        // global index of local row 0 (0 unless slab-decomposed)
  int nrows_glob;  // global row count
  int five;        // 1: Ma = Mb = I (finest level: 5-point stencil)
  int crow_shift;    // row-slab decomposition: coarse local row = (fine local row >> 1) + crow_shift (0 when not
  int nrows_coarse;  // decomposed); rows of the coarse array the transfers address (0 => nrows / 2)
  int rq_lo, rq_hi;  // local rows [rq_lo, rq_hi) whose w^T A w, w^T w the fused Rayleigh stage sums (slab piece: the owned rows)
  int uni;           // 1: constant 5-point stencil (all off-diagonals == uni_c, ka_di + kb_di == uni_d on every point): fused_uni.cu
                     // 2: 9-point level whose four tridiagonal factors are constant but for their last diagonal entry: fused_uni9.cu
  double uni_c, uni_d;
  double u9[12];     // uni == 2: ka_off, ka_di, ma_off, ma_di, ka_di_last, ma_di_last, then the same of kb / mb
  const double *ka_lo, *ka_di, *ka_up, *ma_lo, *ma_di, *ma_up;  // length nrows_glob
  const double *kb_lo, *kb_di, *kb_up, *mb_lo, *mb_di, *mb_up;  // length ncols
};

// One level of the general banded complex128 path (band.cu): A kept by diagonals, vals[k*n + i] = A[i, i + offs[k]]
// (0 where i + offs[k] falls outside [0, n)), complex interleaved (re, im).
struct BandDev {
  int n;        // unknowns
  int ndiag;    // stored diagonals
  int idiag;    // index of the main diagonal (offs[idiag] == 0)
  const int *offs;      // device, ascending
  const double2 *vals;  // device, ndiag * n
};

constexpr int kWarp = 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (read-once) 16-byte load / store: keep L1 for the halo re-reads
__device__ __forceinline__ double2 ld_stream2(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream2(double *p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---- per-warp row rings in shared memory (fused_uni.cu, fused_uni9.cu): one ring row = 64 double2 = 1024 bytes, addressed
// as 32-bit shared-space byte offsets so that a slot costs an add and a mask (power-of-two rings) or an add, a compare
// and a select (others) instead of an index multiply chain
__device__ __forceinline__ void cpa16s(unsigned s, const void *gmem, bool valid) {
  const int bytes = valid ? 16 : 0;   // 0: nothing is read from gmem, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ double2 lds2(unsigned s) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(s) : "memory");
  return r;
}
__device__ __forceinline__ void sts2(unsigned s, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(s), "d"(a), "d"(b) : "memory");
}
// byte offset of the ring row `back` bytes behind / `ahead` bytes in front of `cur` in a ring of R rows
template <int R>
__device__ __forceinline__ unsigned ring_back(unsigned cur, unsigned back) {
  if ((R & (R - 1)) == 0) return (cur - back) & (unsigned)(R * 1024 - 1);
  const int s = (int)cur - (int)back;
  return (unsigned)(s + ((s < 0) ? R * 1024 : 0));
}
template <int R>
__device__ __forceinline__ unsigned ring_fwd(unsigned cur, unsigned ahead) {
  if ((R & (R - 1)) == 0) return (cur + ahead) & (unsigned)(R * 1024 - 1);
  const unsigned s = cur + ahead;
  return s - ((s >= (unsigned)(R * 1024)) ? (unsigned)(R * 1024) : 0u);
}

}  // namespace mgcmt
