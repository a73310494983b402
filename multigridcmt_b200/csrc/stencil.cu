// stencil.cu -- row-marching stencil kernels: weighted Jacobi, residual, operator apply.
//
// One thread owns two adjacent columns and marches down a chunk of rows with a three-row window in
// registers, so every grid value is read from HBM once per sweep (plus one halo row per chunk, which
// the neighbouring CTA has just pulled through L2).  Rows are prefetched kPF deep with cp.async
// (LDGSTS) into a per-thread shared-memory ring -- each thread only ever reads the slots it filled
// itself, so the ring needs no barrier, costs no registers, and keeps ~kPF x 32 B per thread in
// flight (what hides HBM latency at the low occupancy a register-windowed fp64 kernel has).
// Left/right neighbours come from warp shuffles; the two edge lanes of each warp prefetch the one
// extra value they need through the same ring.
//
// Reference arithmetic being restated (paths relative to the reference root):
//   weighted Jacobi   MGCMTSolver.py:182-208   v <- (I - w D^-1 A) v + w D^-1 f
//   residual          MGCMTSolver.py:315       f - (A - shift I) v
// with A_l = Ma (x) Kb + Ka (x) Mb and the shift applied as -shift*I on every level
// (MGCMTSolver.py:287-288).
#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

constexpr int kPF = 8;  // prefetch depth in rows (power of two)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct Row4 {
  double xl, x0, x1, xr;
};

// horizontal factors applied to a row: t = Kb row, s = Mb row at the thread's two columns
struct HV {
  double t0, t1, s0, s1, x0, x1;
};

template <int OP, bool FIVE, int TPB>
__global__ void __launch_bounds__(TPB)
stencil_march_kernel(LevelDev L, double shift, double omega, const double *__restrict__ v,
                     const double *__restrict__ f, double *__restrict__ out,
                     const double *__restrict__ halo_top, const double *__restrict__ halo_bot,
                     int rows_per_cta) {
  __shared__ double2 ring_v[kPF][TPB];
  __shared__ double2 ring_f[kPF][TPB];
  __shared__ double ring_e[kPF][TPB / 32][2];

  const int tid = threadIdx.x;
  const int j0 = (blockIdx.x * TPB + tid) * 2;
  const bool active = j0 < L.ncols;
  const int lane = tid & 31, warp = tid >> 5;
  const int i_begin = blockIdx.y * rows_per_cta;
  const int i_end = min(i_begin + rows_per_cta, L.nrows);
  const int jc = active ? j0 : 0;

  // issue the asynchronous copies of local row i (v row i, and f row i when it is an interior row of
  // this chunk); rows -1 / nrows come from the halo pointers (or are Dirichlet zeros)
  auto issue = [&](int i) {
    const int slot = (i + 1) & (kPF - 1);
    const double *src;
    if (i < 0) src = halo_top;
    else if (i >= L.nrows) src = halo_bot;
    else src = v + (size_t)i * L.ncols;
    const bool ok = active && src != nullptr && i <= i_end;
    const double *s = ok ? src : v;
    cp_async16(&ring_v[slot][tid], s + jc, ok);
    if (lane == 0) cp_async8(&ring_e[slot][warp][0], s + (jc > 0 ? jc - 1 : 0), ok && jc > 0);
    if (lane == 31) cp_async8(&ring_e[slot][warp][1], s + (jc + 2 < L.ncols ? jc + 2 : 0), ok && jc + 2 < L.ncols);
    if (OP != OP_APPLY && OP != OP_RAYLEIGH) {
      const bool okf = active && i >= i_begin && i < i_end;
      cp_async16(&ring_f[slot][tid], f + (okf ? (size_t)i * L.ncols + jc : 0), okf);
    }
    cp_async_commit();
  };
  auto fetch = [&](int i) {
    const int slot = (i + 1) & (kPF - 1);
    const double2 x = ring_v[slot][tid];
    Row4 r;
    r.x0 = x.x;
    r.x1 = x.y;
    r.xl = __shfl_up_sync(0xffffffffu, x.y, 1);
    r.xr = __shfl_down_sync(0xffffffffu, x.x, 1);
    if (lane == 0) r.xl = ring_e[slot][warp][0];
    if (lane == 31) r.xr = ring_e[slot][warp][1];
    return r;
  };

  const double kbl0 = L.kb_lo[jc], kbd0 = L.kb_di[jc], kbu0 = L.kb_up[jc];
  const double kbl1 = L.kb_lo[jc + 1], kbd1 = L.kb_di[jc + 1], kbu1 = L.kb_up[jc + 1];
  double mbl0 = 0, mbd0 = 1, mbu0 = 0, mbl1 = 0, mbd1 = 1, mbu1 = 0;
  if (!FIVE) {
    mbl0 = L.mb_lo[jc]; mbd0 = L.mb_di[jc]; mbu0 = L.mb_up[jc];
    mbl1 = L.mb_lo[jc + 1]; mbd1 = L.mb_di[jc + 1]; mbu1 = L.mb_up[jc + 1];
  }
  auto horiz = [&](const Row4 &r) {
    HV h;
    h.x0 = r.x0;
    h.x1 = r.x1;
    h.t0 = kbl0 * r.xl + kbd0 * r.x0 + kbu0 * r.x1;
    h.t1 = kbl1 * r.x0 + kbd1 * r.x1 + kbu1 * r.xr;
    if (!FIVE) {
      h.s0 = mbl0 * r.xl + mbd0 * r.x0 + mbu0 * r.x1;
      h.s1 = mbl1 * r.x0 + mbd1 * r.x1 + mbu1 * r.xr;
    } else {
      h.s0 = r.x0;
      h.s1 = r.x1;
    }
    return h;
  };

  // prologue: rows i_begin-1 .. i_begin+kPF-2 in flight (one commit group per row)
#pragma unroll
  for (int d = 0; d < kPF; ++d) issue(i_begin - 1 + d);

  HV p, c, n;
  cp_async_wait<kPF - 2>();  // rows i_begin-1 and i_begin have landed
  p = horiz(fetch(i_begin - 1));
  c = horiz(fetch(i_begin));

  double last_kad = 0, last_mad = 0, winv0 = 0, winv1 = 0;
  bool have_w = false;
  double rq_num = 0.0, rq_den = 0.0;

  for (int i = i_begin; i < i_end; ++i) {
    // slot of row i-1 is free now: refill it with row i-1+kPF, then wait for row i+1
    issue(i - 1 + kPF);
    cp_async_wait<kPF - 2>();
    n = horiz(fetch(i + 1));
    double2 ff = make_double2(0.0, 0.0);
    if (OP != OP_APPLY && OP != OP_RAYLEIGH) ff = ring_f[(i + 1) & (kPF - 1)][tid];
    const int gi = L.row0 + i;
    const double kal = L.ka_lo[gi], kad = L.ka_di[gi], kau = L.ka_up[gi];
    double av0, av1, mad = 1.0;
    if (FIVE) {
      av0 = kal * p.x0 + (c.t0 + kad * c.x0) + kau * n.x0 - shift * c.x0;
      av1 = kal * p.x1 + (c.t1 + kad * c.x1) + kau * n.x1 - shift * c.x1;
    } else {
      const double mal = L.ma_lo[gi], mau = L.ma_up[gi];
      mad = L.ma_di[gi];
      av0 = (mal * p.t0 + kal * p.s0) + (mad * c.t0 + kad * c.s0) + (mau * n.t0 + kau * n.s0) - shift * c.x0;
      av1 = (mal * p.t1 + kal * p.s1) + (mad * c.t1 + kad * c.s1) + (mau * n.t1 + kau * n.s1) - shift * c.x1;
    }
    double2 o;
    if (OP == OP_RAYLEIGH) {
      rq_num += c.x0 * av0;
      rq_num += c.x1 * av1;
      rq_den += c.x0 * c.x0;
      rq_den += c.x1 * c.x1;
      o.x = o.y = 0.0;
    } else if (OP == OP_JACOBI) {
      if (!have_w || kad != last_kad || mad != last_mad) {  // block-uniform: first and last rows only
        const double d0 = FIVE ? (kad + kbd0) - shift : (mad * kbd0 + kad * mbd0) - shift;
        const double d1 = FIVE ? (kad + kbd1) - shift : (mad * kbd1 + kad * mbd1) - shift;
        winv0 = omega / d0;
        winv1 = omega / d1;
        last_kad = kad; last_mad = mad; have_w = true;
      }
      o.x = c.x0 + winv0 * (ff.x - av0);
      o.y = c.x1 + winv1 * (ff.y - av1);
    } else if (OP == OP_RESIDUAL) {
      o.x = ff.x - av0;
      o.y = ff.y - av1;
    } else {
      o.x = av0;
      o.y = av1;
    }
    if (OP != OP_RAYLEIGH && active) st_stream2(out + (size_t)i * L.ncols + j0, o);
    p = c;
    c = n;
  }
  cp_async_wait<0>();
  if (OP == OP_RAYLEIGH) {
    // block partials of x^T A x and x^T x: out[b] and out[nblocks + b], summed in order by finish_kernel
    __shared__ double red[2][TPB / 32];
    const double a = warp_sum(rq_num), b = warp_sum(rq_den);
    if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
    __syncthreads();
    if (tid == 0) {
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int w = 0; w < TPB / 32; ++w) { sa += red[0][w]; sb += red[1][w]; }
      const int nb = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
      out[bid] = sa;
      out[nb + bid] = sb;
    }
  }
}

// launch geometry: CTAs of 32/64/128 threads (2 columns per thread); rows per CTA chosen so the grid
// has a few CTAs per SM even on small levels, but chunks stay long enough on big ones that the two
// re-read halo rows are a few per cent (and those come from L2)
template <int OP, int TPB>
static cudaError_t launch_march_t(const LevelDev &L, double shift, double omega, const double *v,
                                  const double *f, double *out, const double *halo_top,
                                  const double *halo_bot, cudaStream_t s) {
  const int col_blocks = (L.ncols / 2 + TPB - 1) / TPB;
  int rpc = 32;
  while (rpc > 2 && (long long)col_blocks * ((L.nrows + rpc - 1) / rpc) < 148LL * 4) rpc >>= 1;
  if (rpc > L.nrows) rpc = L.nrows;
  dim3 grid(col_blocks, (L.nrows + rpc - 1) / rpc);
  if (L.five)
    stencil_march_kernel<OP, true, TPB><<<grid, TPB, 0, s>>>(L, shift, omega, v, f, out, halo_top, halo_bot, rpc);
  else
    stencil_march_kernel<OP, false, TPB><<<grid, TPB, 0, s>>>(L, shift, omega, v, f, out, halo_top, halo_bot, rpc);
  count_launch();
  return cudaGetLastError();
}

template <int OP>
static cudaError_t launch_march(const LevelDev &L, double shift, double omega, const double *v,
                                const double *f, double *out, const double *halo_top,
                                const double *halo_bot, cudaStream_t s) {
  if (L.ncols >= 1024) return launch_march_t<OP, 128>(L, shift, omega, v, f, out, halo_top, halo_bot, s);
  if (L.ncols >= 256) return launch_march_t<OP, 64>(L, shift, omega, v, f, out, halo_top, halo_bot, s);
  return launch_march_t<OP, 32>(L, shift, omega, v, f, out, halo_top, halo_bot, s);
}

cudaError_t launch_jacobi_sweep(const LevelDev &L, double shift, double omega, const double *v_in,
                                const double *f, double *v_out, const double *halo_top,
                                const double *halo_bot, cudaStream_t s) {
  return launch_march<OP_JACOBI>(L, shift, omega, v_in, f, v_out, halo_top, halo_bot, s);
}
cudaError_t launch_residual(const LevelDev &L, double shift, const double *v, const double *f, double *r,
                            const double *halo_top, const double *halo_bot, cudaStream_t s) {
  return launch_march<OP_RESIDUAL>(L, shift, 0.0, v, f, r, halo_top, halo_bot, s);
}
cudaError_t launch_apply(const LevelDev &L, double shift, const double *x, double *y, const double *halo_top,
                         const double *halo_bot, cudaStream_t s) {
  return launch_march<OP_APPLY>(L, shift, 0.0, x, nullptr, y, halo_top, halo_bot, s);
}
// partials: 2 * (number of CTAs) doubles (see march_grid_blocks); block b writes x^T A x and x^T x partials
cudaError_t launch_rayleigh_partials(const LevelDev &L, const double *x, double *partials, const double *halo_top,
                                     const double *halo_bot, cudaStream_t s) {
  return launch_march<OP_RAYLEIGH>(L, 0.0, 0.0, x, nullptr, partials, halo_top, halo_bot, s);
}
int march_grid_blocks(const LevelDev &L) {
  const int TPB = L.ncols >= 1024 ? 128 : (L.ncols >= 256 ? 64 : 32);
  const int col_blocks = (L.ncols / 2 + TPB - 1) / TPB;
  int rpc = 32;
  while (rpc > 2 && (long long)col_blocks * ((L.nrows + rpc - 1) / rpc) < 148LL * 4) rpc >>= 1;
  if (rpc > L.nrows) rpc = L.nrows;
  return col_blocks * ((L.nrows + rpc - 1) / rpc);
}

}  // namespace mgcmt
