// fused.cu -- one HBM pass per V-cycle leg: NU Jacobi sweeps fused with the transfer that follows /
// precedes them (temporal blocking in registers).
//
//   down leg:  v_out = J^NU(v_in; f),   r_coarse = R (f - A_s v_out)        (MGCMTSolver.py:313-315)
//   up leg:    v_out = J^NU(v_in + P e_coarse; f)                           (MGCMTSolver.py:323-326)
//
// Un-fused, a V(4,4) level costs 4 x 24 B + 18 B per unknown and leg; here a
// leg reads v and f once and writes v once (+ the coarse vector): 26 B per unknown.
//
// How: every WARP owns a strip of 32*C columns (C per lane) and streams down a chunk of rows.  The NU
// sweeps (+ the residual) form a software pipeline along the row direction: stage k lags stage k-1
// by one row, and each stage keeps just three doubles per column -- two partial row sums and the
// centre value -- because an arriving row is *scattered* into the three output rows it touches
// (as upper neighbour of row n-1, centre of row n, lower neighbour of row n+1).  Horizontal
// neighbours come from warp shuffles only; strips overlap by HALO columns and chunks by HALO rows,
// and the overlap is recomputed (trapezoid blocking), so warps never synchronise with each other.
// Input rows are prefetched with cp.async into a per-thread shared-memory ring (no barrier needed:
// a thread only reads slots it filled itself); the f ring doubles as the queue that hands f to the
// later stages.
//
// Arithmetic is the reference's weighted Jacobi v <- v + w D^-1 (f - A_s v) (MGCMTSolver.py:182-208)
// with A_l = Ma (x) Kb + Ka (x) Mb - shift I; only the order of the row-sum additions differs from
// stencil.cu (parity to the oracle is tested at 1e-12).
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

namespace {

#ifndef MGCMT_VRING
#define MGCMT_VRING 4
#endif
#ifndef MGCMT_MIN_CTAS
#define MGCMT_MIN_CTAS 2
#endif
#ifndef MGCMT_GS9_CTAS
#define MGCMT_GS9_CTAS 1   // resident CTAs the 8-stage Gauss-Seidel legs of 9-point levels are compiled for (3 = 168 registers: measured 3 % slower per cycle)
#endif
constexpr int kVRing = MGCMT_VRING;  // prefetch depth of the v ring (rows, power of two)
// Pipeline skew: stage k works kSkew * k rows behind the input row.  With kSkew = 1 a stage consumes what its
// predecessor produced in the SAME time step, so the stages of a step form one serial dependency chain.  With
// kSkew = 2 a stage consumes what its predecessor produced in the PREVIOUS step (kept in C registers per stage):
// all stages of a step are independent of each other; the price is NSTAGE-1 extra rows of pipeline fill per chunk and
// a deeper f queue.  Measured on B200 (4096^2): no gain (down leg 132 vs 131 us) -- the stage-to-stage chain is not
// what limits the kernel -- so the default stays 1; the variant is kept for the next profiling round.
// It is a template parameter (SKEW): 1 for the bandwidth-bound legs, 2 for the legs of small levels, which are bound by
// the length of that serial chain (a 512^2 Gauss-Seidel leg: 8 stages x ~150 cycles per step, 20 us per launch).
// f ring depth = prefetch depth + SKEW * NSTAGE + 2 rows of queue (per instantiation; not a power of two: slots
// are tracked incrementally)
__host__ __device__ constexpr int f_ring_slots(int nstage, int skew) { return kVRing + skew * nstage + 2; }
constexpr int kERing = 4;    // coarse-row ring (PROLONG)
constexpr int kWarps = 4;    // warps per CTA (independent strips)

__device__ __forceinline__ void cpa16(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa8(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// per-row coefficients staged in shared memory for the rows a chunk touches
struct RowCoef {
  double ka_lo, ka_di, ka_up, slow;  // slow: see the table fill in the kernel
  double ma_lo, ma_di, ma_up, pad;
};

template <int C>
struct Stage {
  double a1[C], a2[C], xc[C];
};

}  // namespace

// GS = 0: NU weighted-Jacobi sweeps.  GS = 1: NU colour stages of Gauss-Seidel/SOR -- two per sweep on the 5-point
// level (red, black), four per sweep on 9-point levels ((0,0),(1,1),(0,1),(1,0): the order of gs.cu and of the CPU twin
// oracle Solver.rbgs).  A colour stage is a Jacobi stage that only writes the points of its colour and passes the others
// through; which points those are is known at compile time (row parity from the unrolled step, column parity from the
// lane's even first column), so the skipped half / three quarters of the arithmetic is simply not generated.
template <bool FIVE, int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int C, int GS, int SKEW>
__global__ void __launch_bounds__(kWarps * 32, (C == 4 ? MGCMT_MIN_CTAS : (GS != 0 && !FIVE && NU == 8 ? MGCMT_GS9_CTAS : 1)))
fused_leg_kernel(LevelDev L, double shift, double omega, const double *__restrict__ v_in,
                 const double *__restrict__ f, double *__restrict__ v_out,
                 const double *__restrict__ e_coarse, double *__restrict__ r_coarse, int rows_per_chunk) {
  constexpr int HALO = (NU + 3) & ~1;  // >= NU + 2, even (threads load 16-byte column pairs)
  constexpr int WCOLS = 32 * C;
  constexpr int USEFUL = WCOLS - 2 * HALO;
  constexpr int NSTAGE = NU + (RESTRICT ? 1 : 0);  // pipeline stages after the input
  constexpr int kSkew = SKEW;
  constexpr int kFRing = f_ring_slots(NSTAGE, SKEW);
  constexpr int NCOL = FIVE ? 2 : 4;               // colours of the Gauss-Seidel ordering
  constexpr int CE = C / 2;                        // coarse columns per thread
  static_assert(C == 2 || C == 4, "C must be 2 or 4");
  static_assert(USEFUL % 2 == 0, "strip width must be even");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: v ring [kVRing][128][C], f ring [kFRing][128][C], e ring [kERing][128][CE], row table
  double *ring_v = reinterpret_cast<double *>(smem_raw);
  double *ring_f = ring_v + kVRing * kWarps * 32 * C;
  double *ring_e = ring_f + kFRing * kWarps * 32 * C;
  RowCoef *rowtab = reinterpret_cast<RowCoef *>(ring_e + kERing * kWarps * 32 * (PROLONG ? CE : 0));

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(r0 + rows_per_chunk, L.nrows);
  // first / last input row the useful outputs depend on; the loop starts on an even row at or before
  // t_first (two rows earlier for PROLONG so both coarse rows of the first fine row are in hand)
  const int t_first = r0 - NU - (RESTRICT ? 1 : 0);
  const int t_load_last = r1 - 1 + NU + (RESTRICT ? 2 : 0);                         // last input row that matters
  const int t_last = t_load_last + (kSkew - 1) * (NSTAGE > 0 ? NSTAGE - 1 : 0);     // last time step (inclusive)
  const int t_begin = (t_first - (PROLONG ? 2 : 0)) & ~1;
  const int tab0 = t_begin - kSkew * NSTAGE - 1;        // first row held in rowtab
  const int ntab = t_last + 3 - tab0;                   // ... up to row t_last + 2
  const int nrc = L.nrows_coarse ? L.nrows_coarse : L.nrows / 2, ncc = L.ncols / 2;
  const int cs = L.crow_shift;

  // reference row for the cached omega/diag: the middle of the chunk (global row index)
  const int gref = min(max(L.row0 + (r0 + r1) / 2, 0), L.nrows_glob - 1);
  const double kad_ref = L.ka_di[gref];
  const double mad_ref = FIVE ? 1.0 : L.ma_di[gref];
  // all six row coefficients of the reference row: steps whose rows all look like it (every interior step of a
  // constant-coefficient operator) use these registers instead of the shared-memory row table
  const double kal_ref = L.ka_lo[gref], kau_ref = L.ka_up[gref];
  const double mal_ref = FIVE ? 0.0 : L.ma_lo[gref], mau_ref = FIVE ? 0.0 : L.ma_up[gref];

  for (int i = tid; i < ntab; i += kWarps * 32) {
    const int row = tab0 + i;
    // global row: a slab holds rows row0 .. row0 + nrows - 1 of an nrows_glob-row grid (row0 = 0, nrows_glob =
    // nrows when not decomposed); only rows outside the GLOBAL grid are Dirichlet rows
    const int gg = L.row0 + row;
    const int g = min(max(gg, 0), L.nrows_glob - 1);
    const bool rin = (gg >= 0 && gg < L.nrows_glob);
    RowCoef rc;
    // Rows outside the grid get an all-zero operator row: with zero-filled inputs every stage then
    // reproduces the Dirichlet zeros there by itself (x + w (0 - 0) = 0), so the hot loop needs no masks.
    rc.ka_lo = rin ? L.ka_lo[g] : 0.0; rc.ka_di = rin ? L.ka_di[g] : 0.0; rc.ka_up = rin ? L.ka_up[g] : 0.0;
    if (FIVE) { rc.ma_lo = 0.0; rc.ma_di = rin ? 1.0 : 0.0; rc.ma_up = 0.0; }
    else { rc.ma_lo = rin ? L.ma_lo[g] : 0.0; rc.ma_di = rin ? L.ma_di[g] : 0.0; rc.ma_up = rin ? L.ma_up[g] : 0.0; }
    // slow = 1 if, while row `row` is the newest input, any row the stages touch (row-NSTAGE .. row+1) lies outside
    // the grid or has coefficients different from the reference row's: that step reads the table (and may
    // recompute omega/diag); all other steps run on the reference registers
    double slow = 0.0;
    for (int d = -1; d <= kSkew * NSTAGE; ++d) {
      const int g2 = gg - d;
      if (g2 < 0 || g2 >= L.nrows_glob) { slow = 1.0; continue; }
      bool same = (L.ka_lo[g2] == kal_ref && L.ka_di[g2] == kad_ref && L.ka_up[g2] == kau_ref);
      if (!FIVE) same = same && (L.ma_lo[g2] == mal_ref && L.ma_di[g2] == mad_ref && L.ma_up[g2] == mau_ref);
      if (!same) slow = 1.0;
    }
    rc.slow = slow;
    rc.pad = 0.0;
    rowtab[i] = rc;
  }
  __syncthreads();

  // PROLONG && RESTRICT = "up leg + Rayleigh quotient": the extra stage after the sweeps evaluates A_s w on the final
  // iterate and every warp leaves its partial sums of w^T A_s w and w^T w in r_coarse[slot], r_coarse[nslots + slot]
  constexpr bool RQ = PROLONG && RESTRICT;
  const int rq_slot = (blockIdx.y * gridDim.x + blockIdx.x) * kWarps + warp;
  const int rq_nslots = gridDim.x * gridDim.y * kWarps;
  const int strip = blockIdx.x * kWarps + warp;
  const int u0 = strip * USEFUL;  // first useful fine column of this strip
  if (u0 >= L.ncols) {            // surplus warp (no CTA-wide barrier after this point)
    if (RQ && lane == 0) { r_coarse[rq_slot] = 0.0; r_coarse[rq_nslots + rq_slot] = 0.0; }
    return;
  }
  const int u1 = min(u0 + USEFUL, L.ncols);
  const int c0 = u0 - HALO + C * lane;  // first column of this thread (even)

  // rows for which column q carries a real unknown: [0, rlim[q])  (0 rows for columns outside the grid)
  unsigned rlim[C];
  bool colout[C];
  double kbl[C], kbd[C], kbu[C], mbl[C], mbd[C], mbu[C], wref[C];
#pragma unroll
  for (int q = 0; q < C; ++q) {
    const int j = c0 + q;
    rlim[q] = (j >= 0 && j < L.ncols) ? (unsigned)L.nrows_glob : 0u;
    colout[q] = (j >= u0 && j < u1);
    const int jc = min(max(j, 0), L.ncols - 1);
    const bool cin = (j >= 0 && j < L.ncols);
    // columns outside the grid: zero operator column and zero relaxation weight (see the row table)
    kbl[q] = cin ? L.kb_lo[jc] : 0.0; kbd[q] = cin ? L.kb_di[jc] : 0.0; kbu[q] = cin ? L.kb_up[jc] : 0.0;
    if (FIVE) { mbl[q] = 0.0; mbd[q] = 1.0; mbu[q] = 0.0; }
    else { mbl[q] = cin ? L.mb_lo[jc] : 0.0; mbd[q] = cin ? L.mb_di[jc] : 0.0; mbu[q] = cin ? L.mb_up[jc] : 0.0; }
    wref[q] = cin ? omega / ((mad_ref * L.kb_di[jc] + kad_ref * (FIVE ? 1.0 : L.mb_di[jc])) - shift) : 0.0;
  }
  // 16-byte granules of this thread: both columns of a pair are inside or outside the grid together
  // (c0 and ncols are even)
  bool pairin[C / 2];
#pragma unroll
  for (int g = 0; g < C / 2; ++g) pairin[g] = rlim[2 * g] != 0u;

  // ring layout [slot][granule g][thread] of 16-byte granules: a warp's accesses to one granule are 512
  // contiguous bytes, i.e. bank-conflict free for both the cp.async writes and the LDS.128 reads
  constexpr int NT = kWarps * 32;
  constexpr int SLOT = NT * C;          // doubles per ring slot
  constexpr int ESLOT = NT * CE;
  double2 *my_v = reinterpret_cast<double2 *>(ring_v) + tid;  // granule g of slot s: my_v[(s*(C/2) + g) * NT]
  double2 *my_f = reinterpret_cast<double2 *>(ring_f) + tid;
  double *my_e = ring_e + tid;                                 // coarse value g of slot s: my_e[(s*CE + g) * NT]

  // ---- asynchronous row fetch -------------------------------------------------------------------
  auto issue = [&](int t, int fslot) {
    // rows outside the slab array or outside the global grid are zero-filled
    const bool rowin = (t >= 0 && t < L.nrows) && t <= t_load_last && (unsigned)(t + L.row0) < (unsigned)L.nrows_glob;
    if (!ZEROV) {
      double2 *dst = my_v + (t & (kVRing - 1)) * (C / 2) * NT;
#pragma unroll
      for (int g = 0; g < C / 2; ++g) {
        const bool ok = rowin && pairin[g];
        cpa16(dst + g * NT, v_in + (ok ? (size_t)t * L.ncols + c0 + 2 * g : 0), ok);
      }
    }
    {
      double2 *dst = my_f + fslot * (C / 2) * NT;
#pragma unroll
      for (int g = 0; g < C / 2; ++g) {
        const bool ok = rowin && pairin[g];
        cpa16(dst + g * NT, f + (ok ? (size_t)t * L.ncols + c0 + 2 * g : 0), ok);
      }
    }
    if (PROLONG && (t & 1) == 0) {
      // coarse row I = t/2 is first needed by fine row t (even); coarse columns c0/2 .. c0/2+CE-1
      const int I = (t >> 1) + cs;
      const int Gc = ((t + L.row0) >> 1);  // global coarse row
      const bool rowc = I >= 0 && I < nrc && t <= t_load_last && Gc >= 0 && Gc < (L.nrows_glob >> 1);
      double *dst = my_e + (I & (kERing - 1)) * ESLOT;
#pragma unroll
      for (int g = 0; g < CE; ++g) {
        const int J = (c0 >> 1) + g;
        const bool ok = rowc && J >= 0 && J < ncc;
        cpa8(dst + g * NT, e_coarse + (ok ? (size_t)I * ncc + J : 0), ok);
      }
    }
    cpa_commit();
  };

  Stage<C> st[NSTAGE > 0 ? NSTAGE : 1];
  double pend[NSTAGE > 0 ? NSTAGE : 1][C];  // kSkew == 2: pend[k] = what stage k-1 finalised in the previous step
#pragma unroll
  for (int k = 0; k < NSTAGE; ++k)
#pragma unroll
    for (int q = 0; q < C; ++q) st[k].a1[q] = st[k].a2[q] = st[k].xc[q] = pend[k][q] = 0.0;

  double eprev[C], ecur[C];  // column-interpolated coarse rows I-1 and I (PROLONG)
#pragma unroll
  for (int q = 0; q < C; ++q) eprev[q] = ecur[q] = 0.0;
  double rq_num = 0.0, rq_den = 0.0;  // RQ mode
  double racc[CE];           // running full-weighting row sum (RESTRICT)
#pragma unroll
  for (int g = 0; g < CE; ++g) racc[g] = 0.0;

  // one time step: input row t enters, every stage finalises one row.  ODD (row parity) and SLOW (some
  // finalised row needs omega/diag recomputed) are compile-time so the common path is branch-free.
  int fs = 0;  // f-ring slot of the current input row t: (t - t_begin) mod kFRing
  auto step = [&](int t, auto odd_tag, auto slow_tag) {
    constexpr bool ODD = decltype(odd_tag)::value;
    constexpr bool SLOW = decltype(slow_tag)::value;
    cpa_wait<kVRing - 1>();  // row t has landed (the kVRing-1 younger groups may still be in flight)

    // ---- stage 0: the input row t --------------------------------------------------------------
    double x[C];
    {
      const double2 *src = my_v + (t & (kVRing - 1)) * (C / 2) * NT;
#pragma unroll
      for (int g = 0; g < C / 2; ++g) {
        const double2 xx = ZEROV ? make_double2(0.0, 0.0) : src[g * NT];
        x[2 * g] = xx.x;
        x[2 * g + 1] = xx.y;
      }
    }
    if (PROLONG) {
      if (!ODD) {
        // new coarse row I = t/2: interpolate along columns.  fine col c0+2g (even) = 1/2 (E[J-1] + E[J]),
        // fine col c0+2g+1 = E[J], J = c0/2 + g.  Lane 0 has no left neighbour: its first column is the
        // outermost halo column of the strip; the up leg has no residual stage, so HALO = NU + 2 leaves
        // two columns of slack and that error never reaches a useful column.
        const double *src = my_e + (((t >> 1) + cs) & (kERing - 1)) * ESLOT;
        double e[CE];
#pragma unroll
        for (int g = 0; g < CE; ++g) e[g] = src[g * NT];
        const double eleft = __shfl_up_sync(0xffffffffu, e[CE - 1], 1);
#pragma unroll
        for (int q = 0; q < C; ++q) eprev[q] = ecur[q];
#pragma unroll
        for (int g = 0; g < CE; ++g) {
          const double em = (g == 0) ? eleft : e[g - 1];
          ecur[2 * g] = 0.5 * (em + e[g]);
          ecur[2 * g + 1] = e[g];
        }
#pragma unroll
        for (int q = 0; q < C; ++q) x[q] += 0.5 * (eprev[q] + ecur[q]);
      } else {
#pragma unroll
        for (int q = 0; q < C; ++q) x[q] += ecur[q];
      }
    }
    if (PROLONG) {  // the interpolated correction is the only input that is not already zero outside the grid
#pragma unroll
      for (int q = 0; q < C; ++q) x[q] = ((unsigned)(t + L.row0) < rlim[q]) ? x[q] : 0.0;
    }

    // refill the ring slot that row t just vacated (f slots are vacated NU+2 rows later; the f ring is
    // deep enough: kVRing + NU + 2 <= kFRing)
    {
      int fnew = fs + kVRing;
      fnew -= (fnew >= kFRing) ? kFRing : 0;
      issue(t + kVRing, fnew);
    }

    // ---- stages 1..NSTAGE: row n of stage k-1 arrives, row n-1 of stage k is finalised -----------
    // (kSkew == 2: in descending order, so a stage reads pend[k] before its predecessor overwrites it)
    double xrow[C];
#pragma unroll
    for (int q = 0; q < C; ++q) xrow[q] = x[q];
#pragma unroll
    for (int kk = 0; kk < NSTAGE; ++kk) {
      const int k = (kSkew == 1) ? kk : NSTAGE - 1 - kk;
      const int n = t - kSkew * k;  // arriving row (of stage k's input)
      const int rho = n - 1;        // finalised row
      if (kSkew != 1) {
#pragma unroll
        for (int q = 0; q < C; ++q) x[q] = (k == 0) ? xrow[q] : pend[k][q];
      }
      const RowCoef *tb = rowtab + (rho - tab0);
      const double cr_ka_up = SLOW ? tb[0].ka_up : kau_ref, cr_ma_up = FIVE ? 0.0 : (SLOW ? tb[0].ma_up : mau_ref);
      const double cn_ka_di = SLOW ? tb[1].ka_di : kad_ref, cn_ma_di = FIVE ? 1.0 : (SLOW ? tb[1].ma_di : mad_ref);
      const double cn_ka_lo = SLOW ? tb[1].ka_lo : kal_ref;
      const double cp_ka_lo = SLOW ? tb[2].ka_lo : kal_ref, cp_ma_lo = FIVE ? 0.0 : (SLOW ? tb[2].ma_lo : mal_ref);
      // horizontal part of the arriving row
      const double xl = __shfl_up_sync(0xffffffffu, x[C - 1], 1);
      const double xr = __shfl_down_sync(0xffffffffu, x[0], 1);
      double T[C], S[C];
#pragma unroll
      for (int q = 0; q < C; ++q) {
        const double xm = (q == 0) ? xl : x[q - 1];
        const double xp = (q == C - 1) ? xr : x[q + 1];
        T[q] = kbl[q] * xm + kbd[q] * x[q] + kbu[q] * xp;
        S[q] = FIVE ? x[q] : (mbl[q] * xm + mbd[q] * x[q] + mbu[q] * xp);
      }
      constexpr bool kIsRes = RESTRICT;  // (only the last stage, tested below with the unrolled k)
      const bool is_res = kIsRes && (k == NSTAGE - 1);
      // Gauss-Seidel colour of this stage and the (compile-time) parity of the finalised row rho = t - k - 1
      // (t_begin, the slab row offset and every lane's first column are even)
      const bool gs_stage = (GS != 0) && !is_res;
      const int colour = k % NCOL;
      const int prho = (ODD ? 1 : 0) ^ ((kSkew * k + 1) & 1);
      auto in_colour = [&](int prow, int pcol) {
        if (!gs_stage) return true;
        if (FIVE) return ((prow + pcol) & 1) == colour;
        return colour == 0 ? (prow == 0 && pcol == 0)
             : colour == 1 ? (prow == 1 && pcol == 1)
             : colour == 2 ? (prow == 0 && pcol == 1) : (prow == 1 && pcol == 0);
      };
      int frho = fs - (kSkew * k + 1);  // slot of row rho (slots of rows before t_begin hold garbage that
      frho += (frho < 0) ? kFRing : 0;  // only ever reaches rows outside every valid region)
      const double2 *fsrc = my_f + frho * (C / 2) * NT;
      double ffv[C];
#pragma unroll
      for (int g = 0; g < C / 2; ++g) {
        const double2 f2 = fsrc[g * NT];
        ffv[2 * g] = f2.x;
        ffv[2 * g + 1] = f2.y;
      }
      double w[C];
#pragma unroll
      for (int q = 0; q < C; ++q) w[q] = wref[q];
      if (SLOW && !is_res) {
        const double cr_ka_di = tb[0].ka_di, cr_ma_di = tb[0].ma_di;  // ma_di == 0 marks a row outside the grid
        if (cr_ma_di != 0.0 && (cr_ka_di != kad_ref || cr_ma_di != mad_ref)) {
#pragma unroll
          for (int q = 0; q < C; ++q)
            w[q] = (rlim[q] != 0u) ? omega / ((cr_ma_di * kbd[q] + cr_ka_di * mbd[q]) - shift) : 0.0;
        }
      }
      double out[C];
#pragma unroll
      for (int q = 0; q < C; ++q) {
        if (in_colour(prho, q & 1)) {
          const double acc = st[k].a1[q] + (FIVE ? cr_ka_up * S[q] : (cr_ma_up * T[q] + cr_ka_up * S[q]));
          const double ff = ffv[q];
          out[q] = is_res ? (ff - acc) : (st[k].xc[q] + w[q] * (ff - acc));
          if (RQ && is_res && rho >= r0 && rho < r1 && rho >= L.rq_lo && rho < L.rq_hi && colout[q]) {  // each useful (owned) point exactly once
            rq_num += st[k].xc[q] * acc;
            rq_den += st[k].xc[q] * st[k].xc[q];
          }
        } else {
          out[q] = st[k].xc[q];  // not this stage's colour: passes through
        }
        // scatter the arriving row n = rho + 1 into the rows still open (only where they will be finalised here)
        if (FIVE) {
          // 5-point: the lower-neighbour part of row n is ka_lo[n] * x[n-1] = ka_lo[n] * xc -- no second open sum needed
          if (in_colour(prho ^ 1, q & 1))
            st[k].a1[q] = cn_ka_lo * st[k].xc[q] + ((T[q] + cn_ka_di * x[q]) - shift * x[q]);
        } else {
          if (in_colour(prho ^ 1, q & 1))
            st[k].a1[q] = st[k].a2[q] + ((cn_ma_di * T[q] + cn_ka_di * S[q]) - shift * x[q]);
          if (in_colour(prho, q & 1))
            st[k].a2[q] = cp_ma_lo * T[q] + cp_ka_lo * S[q];
        }
        st[k].xc[q] = x[q];
      }
      // the finalised row is the next stage's arriving row (this step for kSkew == 1, the next one otherwise)
#pragma unroll
      for (int q = 0; q < C; ++q) {
        x[q] = out[q];
        if (kSkew != 1 && k + 1 < NSTAGE) pend[k + 1][q] = out[q];
      }

      if (!is_res && k == NU - 1) {
        // x = row rho of the NU-th sweep: the smoothed iterate
        const bool rowok = (rho >= r0 && rho < r1);
#pragma unroll
        for (int g = 0; g < C / 2; ++g)
          if (rowok && colout[2 * g])
            st_stream2(v_out + (size_t)rho * L.ncols + c0 + 2 * g, make_double2(x[2 * g], x[2 * g + 1]));
      }
      if (is_res && !RQ) {
        // x = residual row rho (zero outside the grid): full weighting.  Columns first:
        //   coarse J = c0/2 + g  <-  1/4 r[2J] + 1/2 r[2J+1] + 1/4 r[2J+2]
        // rho = t - NU - 1 has the parity of t iff NU is odd
        constexpr bool RHO_ODD = (ODD != ((kSkew * NU + 1) % 2 != 0));
        const double rnext = __shfl_down_sync(0xffffffffu, x[0], 1);
        double crr[CE];
#pragma unroll
        for (int g = 0; g < CE; ++g) {
          const double r2 = (g == CE - 1) ? rnext : x[2 * g + 2];
          crr[g] = 0.25 * x[2 * g] + 0.5 * x[2 * g + 1] + 0.25 * r2;
        }
        if (!RHO_ODD) {
          const int I = (rho >> 1) - 1 + cs;  // coarse row completed by this fine row (as its row 2I+2)
          const int G = ((rho + L.row0) >> 1) - 1;  // its global coarse row: rows outside the coarse grid are never written
          const bool rowok = (I >= (r0 >> 1) + cs && I < (r1 >> 1) + cs && I >= 0 && I < nrc && G >= 0 &&
                              G < (L.nrows_glob >> 1));
#pragma unroll
          for (int g = 0; g < CE; ++g) {
            if (rowok && colout[2 * g]) r_coarse[(size_t)I * ncc + (c0 >> 1) + g] = racc[g] + 0.25 * crr[g];
            racc[g] = 0.25 * crr[g];
          }
        } else {
#pragma unroll
          for (int g = 0; g < CE; ++g) racc[g] += 0.5 * crr[g];
        }
      }
    }
    if (NU == 0 && !RESTRICT) {
      // pure prolongation-correction pass: write the corrected iterate
      const bool rowok = (t >= r0 && t < r1);
#pragma unroll
      for (int g = 0; g < C / 2; ++g)
        if (rowok && colout[2 * g])
          st_stream2(v_out + (size_t)t * L.ncols + c0 + 2 * g, make_double2(x[2 * g], x[2 * g + 1]));
    }
    fs = (fs + 1 == kFRing) ? 0 : fs + 1;
  };

#pragma unroll
  for (int d = 0; d < kVRing; ++d) issue(t_begin + d, d);

  using TrueT = std::integral_constant<bool, true>;
  using FalseT = std::integral_constant<bool, false>;
  for (int t = t_begin; t <= t_last; t += 2) {  // t_begin is even
    const bool slow = (rowtab[t - tab0].slow != 0.0) || (rowtab[t + 1 - tab0].slow != 0.0);
    if (!slow) {
      step(t, FalseT{}, FalseT{});
      step(t + 1, TrueT{}, FalseT{});
    } else {
      step(t, FalseT{}, TrueT{});
      step(t + 1, TrueT{}, TrueT{});
    }
  }
  cpa_wait<0>();
  if (RQ) {
    const double a = warp_sum(rq_num), b = warp_sum(rq_den);
    if (lane == 0) { r_coarse[rq_slot] = a; r_coarse[rq_nslots + rq_slot] = b; }
  }
}

// ---------------------------------------------------------------------------------------------------
template <int C>
static size_t fused_smem_bytes(bool prolong, int rows_per_chunk, int nstage, int skew) {
  size_t b = sizeof(double) * (size_t)(kVRing + f_ring_slots(nstage, skew)) * kWarps * 32 * C;
  if (prolong) b += sizeof(double) * (size_t)kERing * kWarps * 32 * (C / 2);
  b += sizeof(RowCoef) * (size_t)(rows_per_chunk + 3 * skew * nstage + 16);  // rows t_begin-skew*NSTAGE-1 .. t_last+2
  return b;
}

// ctas_per_sm: resident CTAs per SM of the instantiation that will run (occupancy query); the chunk height is chosen
// so that the grid fills whole waves (leg_rows_per_chunk, fused_uni.cu) -- a grid a few CTAs over a wave used to cost
// the 9-point levels a second pass of 24 CTAs
template <int NU, int C>
static void fused_geometry(const LevelDev &L, int ctas_per_sm, int nstage, int *gx, int *rpc_out) {
  constexpr int USEFUL = 32 * C - 2 * ((NU + 3) & ~1);
  const int strips = (L.ncols + USEFUL - 1) / USEFUL;
  *gx = (strips + kWarps - 1) / kWarps;
  *rpc_out = leg_rows_per_chunk(L.nrows, *gx, ctas_per_sm * num_sms(), nstage, 128);
}

int g_fused_skew_cols = 0;     // 9-point levels at most this wide run the skew-2 pipeline (0 = never, the default: measured 6-17 % slower per RB-GS cycle on B200)

template <bool FIVE, int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int C, int GS, int SKEW>
static cudaError_t launch_fused_s(const LevelDev &L, double shift, double omega, const double *v_in,
                                  const double *f, double *v_out, const double *e_coarse, double *r_coarse,
                                  cudaStream_t s, int *slots_out) {
  auto kern = fused_leg_kernel<FIVE, NU, PROLONG, RESTRICT, ZEROV, C, GS, SKEW>;
  constexpr int NSTAGE = NU + (RESTRICT ? 1 : 0);
  static int occ = 0;  // per instantiation: resident CTAs per SM with a 128-row chunk's shared memory
  if (!occ) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarps * 32, fused_smem_bytes<C>(PROLONG, 128, NSTAGE, SKEW));
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
  }
  int gx, rpc;
  fused_geometry<NU, C>(L, occ, SKEW * NSTAGE, &gx, &rpc);
  if (slots_out) {
    *slots_out = gx * ((L.nrows + rpc - 1) / rpc) * kWarps;
    return cudaSuccess;
  }
  const size_t smem = fused_smem_bytes<C>(PROLONG, rpc, NSTAGE, SKEW);
  dim3 grid(gx, (L.nrows + rpc - 1) / rpc);
  kern<<<grid, kWarps * 32, smem, s>>>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, rpc);
  count_launch();
  return cudaGetLastError();
}

template <bool FIVE, int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int C, int GS>
static cudaError_t launch_fused_t(const LevelDev &L, double shift, double omega, const double *v_in,
                                  const double *f, double *v_out, const double *e_coarse, double *r_coarse,
                                  cudaStream_t s, int *slots_out = nullptr) {
  // skew 2 (all stages of a step independent) for the 9-point legs of small levels with a real pipeline (>= 3 stages)
  if constexpr (!FIVE && C == 2 && (NU + (RESTRICT ? 1 : 0)) >= 3) {
    if (L.ncols <= g_fused_skew_cols)
      return launch_fused_s<FIVE, NU, PROLONG, RESTRICT, ZEROV, C, GS, 2>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s, slots_out);
  }
  return launch_fused_s<FIVE, NU, PROLONG, RESTRICT, ZEROV, C, GS, 1>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s, slots_out);
}

template <bool FIVE, int NU, int C, int GS = 0>
static cudaError_t dispatch_mode(const LevelDev &L, int mode, double shift, double omega, const double *v_in,
                                 const double *f, double *v_out, const double *e_coarse, double *r_coarse,
                                 cudaStream_t s, int *slots_out = nullptr) {
  switch (mode) {
    case FUSED_SMOOTH:
      if (NU == 0) return cudaErrorInvalidValue;
      return launch_fused_t<FIVE, NU, false, false, false, C, GS>(L, shift, omega, v_in, f, v_out, nullptr, nullptr, s);
    case FUSED_DOWN:
      return launch_fused_t<FIVE, NU, false, true, false, C, GS>(L, shift, omega, v_in, f, v_out, nullptr, r_coarse, s);
    case FUSED_DOWN_ZERO:
      return launch_fused_t<FIVE, NU, false, true, true, C, GS>(L, shift, omega, v_in, f, v_out, nullptr, r_coarse, s);
    case FUSED_UP:
      return launch_fused_t<FIVE, NU, true, false, false, C, GS>(L, shift, omega, v_in, f, v_out, e_coarse, nullptr, s);
    case FUSED_UP_RQ:  // only the Jacobi NU = 4 up leg of the finest level carries the Rayleigh-quotient stage
      if constexpr (GS == 0 && NU == 4 && FIVE && C == 4)
        return launch_fused_t<true, 4, true, true, false, 4, 0>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s, slots_out);
      else
        return cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}

// Gauss-Seidel legs: `sweeps` full colour sweeps (1..4 on the 5-point level, 1..2 on 9-point levels) per pass
cudaError_t launch_fused_gs_leg(const LevelDev &L, int mode, int sweeps, double shift, double omega,
                                const double *v_in, const double *f, double *v_out, const double *e_coarse,
                                double *r_coarse, cudaStream_t s) {
  if (L.nrows < 2) return cudaErrorInvalidValue;
  if (uni5_available(L)) return launch_uni5_leg(L, 1, mode, sweeps, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  if (uni9_available(L) && mode != FUSED_UP_RQ && sweeps <= 2)
    return launch_uni9_leg(L, 1, mode, sweeps, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  if (sweeps == 0) {  // no smoothing (nu1 = 0 or nu2 = 0): the colour order does not matter, the Jacobi leg's transfer is the same
    return L.five ? dispatch_mode<true, 0, 2>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s)
                  : dispatch_mode<false, 0, 2>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  }
  if (L.five) {
    switch (sweeps) {
      case 1: return dispatch_mode<true, 2, 2, 1>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
      case 2: return dispatch_mode<true, 4, 2, 1>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
      case 3: return dispatch_mode<true, 6, 2, 1>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
      case 4: return dispatch_mode<true, 8, 2, 1>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
    }
    return cudaErrorInvalidValue;
  }
  switch (sweeps) {
    case 1: return dispatch_mode<false, 4, 2, 1>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
    case 2: return dispatch_mode<false, 8, 2, 1>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  }
  return cudaErrorInvalidValue;
}

int g_fused_c5 = MGCMT_FUSED_C5;  // columns per lane on the 5-point level (2 or 4), switchable for A/B timing

// number of per-warp partial-sum slots a FUSED_UP_RQ launch on this level writes (2 * slots doubles), or 0 if that
// mode is not available for the level
int fused_rq_slots(const LevelDev &L, int gs) {
  if (uni5_available(L)) return uni5_rq_slots(L, gs);
  if (gs || !L.five || g_fused_c5 != 4 || L.nrows < 2) return 0;
  int slots = 0;
  if (dispatch_mode<true, 4, 4>(L, FUSED_UP_RQ, 0.0, 1.0, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &slots) != cudaSuccess)
    return 0;
  return slots;
}

int g_fused_c9 = 0;  // columns per lane on 9-point levels: 2, 4, or 0 = 4 on levels >= 4096 wide (enough strips to
                     // fill the GPU; measured 8 % faster per cycle at 16384^2, no gain at 4096^2), else 2

cudaError_t launch_fused_leg(const LevelDev &L, int mode, int nu, double shift, double omega,
                             const double *v_in, const double *f, double *v_out, const double *e_coarse,
                             double *r_coarse, cudaStream_t s) {
  if (L.nrows < 2) return cudaErrorInvalidValue;  // 2-D levels only
  if (uni5_available(L)) return launch_uni5_leg(L, 0, mode, nu, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  if (uni9_available(L) && mode != FUSED_UP_RQ)
    return launch_uni9_leg(L, 0, mode, nu, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
#define NU_CASE(NUV)                                                                                         \
  case NUV:                                                                                                  \
    if (L.five && g_fused_c5 == 4)                                                                           \
      return dispatch_mode<true, NUV, 4>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);      \
    if (!L.five && (g_fused_c9 == 4 || (g_fused_c9 == 0 && L.ncols >= 4096)))                                \
      return dispatch_mode<false, NUV, 4>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);     \
    return L.five ? dispatch_mode<true, NUV, 2>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s) \
                  : dispatch_mode<false, NUV, 2>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  switch (nu) {
    NU_CASE(0) NU_CASE(1) NU_CASE(2) NU_CASE(3) NU_CASE(4)
  }
#undef NU_CASE
  return cudaErrorInvalidValue;
}

}  // namespace mgcmt
