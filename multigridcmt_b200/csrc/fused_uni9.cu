// fused_uni9.cu -- the V-cycle legs of fused.cu for the Galerkin (9-point) levels of a constant-coefficient operator:
// every level below the finest of the 2-D wells (MGCMTSolver.py:318, A_c = R A P) has constant tridiagonal factors
// except for the LAST diagonal entry of each (the truncated last row of R, MGCMTStencilMaker.py:66-68), so the stencil
// is constant in the interior and differs only in the last row and the last column.
//
//   A_l = Ma (x) Kb + Ka (x) Mb - shift I;   with h = x(i, j-1) + x(i, j+1):
//   row i's own part     D_i     = ch h_i + d x_i          ch = ma_di kb_off + ka_di mb_off,  d = ma_di kb_di + ka_di mb_di - shift
//   row i to rows i +- 1 U_i     = cc h_i + cv x_i          cc = ma_off kb_off + ka_off mb_off, cv = ma_off kb_di + ka_off mb_di
//   sweep:  out_i = x_i + w (f_i - (U_{i-1} + D_i + U_{i+1})),  w = omega / d
//
// with the coefficients pre-multiplied by -w (c1..c4) this is 7 fp64 instructions per update instead of ~17 in the
// general kernel, which matters because these levels are not HBM-bound there (2048^2: 55 us per leg against 12-17 us
// of traffic).  Two pipeline forms (template LAG).  LAG = 1 (the one in use): the row a stage finishes is handed to the
// next stage within the step, as in fused.cu.  LAG = 2: the row a stage finishes in step t is consumed by the next stage
// in step t + 1, so within a step every stage works on state of the previous step and no stage waits for another one's
// shuffles -- but twice the pipeline fill and one more row of state per stage; measured slower on every level.
// Where it pays (measured inside the 4-stream bench step, where what counts is the SM time a leg takes from the other
// cycles): levels >= 512 wide, 2.56 -> 2.11 ms per RB-GS step at 4096^2.
//
// State per stage and column: `pre1` (row n-1: everything but the contribution of row n), `uprev` (what row n-1
// gives row n), `ready` (row n-1 finished, next stage's input in the next step); red-black (four-colour) Gauss-Seidel
// stages also keep the row itself for the points that pass through.
// Coefficient classes: rows {interior, last, outside} x columns {interior, last}; interior steps run on registers,
// steps that touch the last row or rows outside the grid (SLOW) add warp-uniform row tests: zero rows outside, the last
// row's constants from the kernel parameters for the one stage per step that works on it.
// The interior diagonal weight is carried as hi + lo like in fused_uni.cu (consistent stencil row sum, no systematic
// shift of the spectrum).  Data movement as in fused_uni.cu: coalesced cp.async ring with XOR swizzle, 32-byte stores.
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

namespace {

#ifndef U9_MINCTAS
#define U9_MINCTAS 2
#endif
constexpr int k9C = 4;
constexpr int k9Warps = 4;
constexpr int k9ERing = 4;
__host__ __device__ constexpr int u9_halo(int nu) { return (nu + 2 + 3) & ~3; }
constexpr int k9VR = 4;
// rows of the w f ring.  LAG 1: rounded up to a power of two, so that a stage's slot is one add and one mask away from the
// newest row's; LAG 2 (twice the rows in flight) keeps the exact count, or the ring would cost a resident CTA
__host__ __device__ constexpr int u9_fring(int nstage, int lag) {
  const int need = k9VR + lag * (nstage > 0 ? nstage - 1 : 0) + 1;
  if (lag != 1) return need;
  int p = 1;
  while (p < need) p <<= 1;
  return p;
}
__device__ __forceinline__ void cpa16(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st4(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// constants of one (row class, column class): everything scaled by -w of the TARGET point
struct C9 {
  double c1, c2;   // U: contribution of a neighbouring row to a target row of this class: c1 h + c2 x
  double c3;       // own row: c3 h
  double ares;     // own row: x coefficient in the residual stage (= -w d, hi part)
  double asm_;     // ... in a sweep (1 - w d, hi part)
  double dlo;      // low part of -w d (interior class only)
  double w;        // scale of f
  double rs;       // w(interior) / w(this class): the residual stage leaves w(class) r, the full weighting wants one scale
};
struct K9 {
  C9 c[3][2];      // [row class: 0 interior, 1 last, 2 outside][column class: 0 interior, 1 last]
  double invw;     // 1 / w of the interior (residual scaling of the full weighting)
};

}  // namespace

// LAG = 2: a stage consumes what its predecessor finished in the previous step (stages independent within a step, twice
// the pipeline fill, one more row of state per stage); LAG = 1: in the same step (the shuffle of every stage is on the
// step's dependency chain, half the fill, less state).
template <int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int GS, int LAG>
__global__ void __launch_bounds__(k9Warps * 32, U9_MINCTAS)
uni9_leg_kernel(LevelDev L, K9 K, const double *__restrict__ v_in, const double *__restrict__ f,
                double *__restrict__ v_out, const double *__restrict__ e_coarse, double *__restrict__ r_coarse,
                int rows_per_chunk) {
  constexpr int C = k9C;
  constexpr int HALO = u9_halo(NU);
  constexpr int USEFUL = 32 * C - 2 * HALO;
  constexpr int NSTAGE = NU + (RESTRICT ? 1 : 0);
  constexpr int NS1 = NSTAGE > 0 ? NSTAGE : 1;
  constexpr int kVR = k9VR;
  constexpr int kFR = u9_fring(NSTAGE, LAG);
  constexpr int AHEAD = kVR - 1;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int WARP_GRAN = (ZEROV ? 0 : kVR * 64) + kFR * 64 + (PROLONG ? k9ERing * 32 : 0);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  double2 *ring_v = reinterpret_cast<double2 *>(smem_raw) + warp * WARP_GRAN;
  double2 *ring_f = ring_v + (ZEROV ? 0 : kVR * 64);
  double2 *ring_e = ring_f + kFR * 64;

  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(r0 + rows_per_chunk, L.nrows);
  // stage k's arriving row in step t is t - LAG k; the row it finishes (t - LAG k - 1) is correct from t_begin + k + 1 on
  const int t_last = r1 + LAG * NU + (RESTRICT ? 1 : -1);
  const int t_begin = (r0 - NU - 2 - (PROLONG ? 2 : 0)) & ~1;
  const int nrc = L.nrows_coarse ? L.nrows_coarse : L.nrows / 2, ncc = L.ncols / 2;
  const int cs = L.crow_shift;
  const int nglob = L.nrows_glob;

  const int strip = blockIdx.x * k9Warps + warp;
  const int u0 = strip * USEFUL;
  if (u0 >= L.ncols) return;  // surplus warp (no CTA-wide barrier in this kernel)
  const int u1 = min(u0 + USEFUL, L.ncols);
  const int cstart = u0 - HALO;
  const int c0 = cstart + C * lane;
  const bool quadin = (c0 >= 0 && c0 < L.ncols);
  const bool quadout = (c0 >= u0 && c0 < u1);
  const bool lastq = quadin && (c0 + C == L.ncols);  // this lane's column 3 is the last column of the grid
  const double hm = quadin ? 0.5 : 0.0;
  const bool st32 = ((reinterpret_cast<uintptr_t>(v_out) & 31) == 0);

  // interior-row constants in registers: columns 0..2 (interior class, zero outside the grid) and column 3
  C9 ki = K.c[0][0], kl = K.c[0][lastq ? 1 : 0];
  if (!quadin) {
    ki.c1 = ki.c2 = ki.c3 = ki.ares = ki.asm_ = ki.dlo = ki.w = ki.rs = 0.0;
    kl = ki;
  }
  // Boundary steps (SLOW) run the interior arithmetic on the register constants as well; what differs is decided by
  // warp-uniform row tests, so the extra work is a few selects per stage plus, for the one stage per step whose row is
  // the LAST grid row, that row's constants fetched from the kernel parameters:
  //   * rows outside the grid: their inputs are zero (the loader fills zeros), so they give nothing to their neighbours;
  //     what the interior constants compute FOR them is discarded (`up_out`: the finished row is forced to zero);
  //   * the last row G: its own part (c3, a, w, rs) when it is opened, and what row G - 1 gives it (c1, c2).
  //     What an outside row G + 1 gives it is zero whatever the constants.
  auto cls1 = [&](int q) -> C9 {   // constants of the last row for this lane's column q
    C9 z = K.c[1][(q == C - 1 && lastq) ? 1 : 0];
    if (!quadin) z.c1 = z.c2 = z.c3 = z.ares = z.asm_ = z.dlo = z.w = z.rs = 0.0;
    return z;
  };
  auto cls0 = [&](int q) -> const C9 & { return (q == C - 1) ? kl : ki; };
  const int gl = nglob - 1 - L.row0;   // local index of the last grid row
  const int go = -L.row0;              // local index of grid row 0
  const double q4 = 0.25 * K.invw, q2 = 0.5 * K.invw;

  // Ring addressing: 32-bit shared-space addresses, one ring row = 64 double2 = 1024 bytes, the w f ring of the
  // LAG 1 form a power of two of rows, so a slot is (newest + constant) & mask.  The copy of column pair G = 32 g + lane lands at position
  // G ^ ((G >> 3) & 1) (g = 1: 512 bytes further); this lane's own four columns are the pairs 2 lane, 2 lane + 1.
  static_assert(AHEAD < kVR && AHEAD + LAG * (NS1 - 1) < kFR, "rings too short");
  const unsigned sm_v = (unsigned)__cvta_generic_to_shared(ring_v), sm_f = (unsigned)__cvta_generic_to_shared(ring_f);
  bool ldin[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int j = cstart + 2 * (32 * g + lane);
    ldin[g] = (j >= 0 && j < L.ncols);
  }
  const unsigned ld_off = (unsigned)(lane ^ ((lane >> 3) & 1)) * 16;
  const unsigned pa_off = (unsigned)((2 * lane) ^ ((lane >> 2) & 1)) * 16, pb_off = pa_off ^ 16;
  // rows the copies may touch: inside the array, inside the grid, not beyond what the last step needs
  const int t_lo = max(0, -L.row0);
  const int t_span = max(0, min(min(L.nrows, t_last + 1), nglob - L.row0) - t_lo);
  // running source pointers of this lane's first column pair in the row the next issue() fetches (predicated-off copies
  // read nothing, so rows before / after the array need no clamping)
  const double *pv = ZEROV ? nullptr : v_in + ((ptrdiff_t)t_begin * L.ncols + cstart + 2 * lane);
  const double *pf = f + ((ptrdiff_t)t_begin * L.ncols + cstart + 2 * lane);

  auto issue = [&](int t, unsigned vo, unsigned fo) {   // vo / fo: byte offsets of the ring rows to fill
    const bool rowin = (unsigned)(t - t_lo) < (unsigned)t_span;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const bool ok = rowin && ldin[g];
      if (!ZEROV) cpa16s(sm_v + vo + ld_off + 512 * g, pv + 64 * g, ok);
      if (NSTAGE > 0) cpa16s(sm_f + fo + ld_off + 512 * g, pf + 64 * g, ok);
    }
    if (!ZEROV) pv += L.ncols;
    pf += L.ncols;
    if (PROLONG && (t & 1) == 0) {
      const int I = (t >> 1) + cs;
      const int Gc = ((t + L.row0) >> 1);
      const int J = c0 >> 1;
      const bool ok = I >= 0 && I < nrc && t <= t_last && Gc >= 0 && Gc < (nglob >> 1) && J >= 0 && J < ncc;
      cpa16(ring_e + (I & (k9ERing - 1)) * 32 + lane, e_coarse + (ok ? (size_t)I * ncc + J : 0), ok);
    }
    cpa_commit();
  };

  // state
  double pre1[NS1][C], uprev[NS1][C], ready[NS1][C], xc[NS1][C];
#pragma unroll
  for (int k = 0; k < NS1; ++k)
#pragma unroll
    for (int q = 0; q < C; ++q) pre1[k][q] = uprev[k][q] = ready[k][q] = xc[k][q] = 0.0;
  double eprev[C], ecur[C];
#pragma unroll
  for (int q = 0; q < C; ++q) eprev[q] = ecur[q] = 0.0;
  double racc[2] = {0.0, 0.0};

  using TrueT = std::integral_constant<bool, true>;
  using FalseT = std::integral_constant<bool, false>;
  unsigned vso = 0, fso = 0;   // byte offsets of the ring rows that hold row t
  // running destination of the row the last sweep finishes in this step (row t - LAG (NU - 1) - 1), this lane's columns
  double *pout = v_out + ((ptrdiff_t)(t_begin - LAG * (NU - 1) - 1) * L.ncols + c0);
  const unsigned out_rows = (unsigned)(r1 - r0);
  auto step = [&](int t, auto odd_tag, auto slow_tag) {
    constexpr bool ODD = decltype(odd_tag)::value;   // parity of t
    constexpr bool SLOW = decltype(slow_tag)::value;
    cpa_wait<AHEAD - 1>();
    __syncwarp();
    issue(t + AHEAD, ring_fwd<kVR>(vso, AHEAD * 1024), ring_fwd<kFR>(fso, AHEAD * 1024));
    // ---- the input row t and its scaled right-hand side ------------------------------------------------
    double x0[C];
    if (ZEROV) {
#pragma unroll
      for (int q = 0; q < C; ++q) x0[q] = 0.0;
    } else {
      const double2 xa = lds2(sm_v + vso + pa_off), xb = lds2(sm_v + vso + pb_off);
      x0[0] = xa.x; x0[1] = xa.y; x0[2] = xb.x; x0[3] = xb.y;
    }
    double wfq[NS1][C];
    if (NSTAGE > 0) {
      const double2 fa = lds2(sm_f + fso + pa_off), fb = lds2(sm_f + fso + pb_off);
      if (SLOW && t == gl) {   // (warp-uniform) the last grid row has its own weights
        wfq[0][0] = cls1(0).w * fa.x; wfq[0][1] = cls1(1).w * fa.y;
        wfq[0][2] = cls1(2).w * fb.x; wfq[0][3] = cls1(3).w * fb.y;
      } else {                 // rows outside the grid arrive as zeros
        wfq[0][0] = ki.w * fa.x; wfq[0][1] = ki.w * fa.y;
        wfq[0][2] = ki.w * fb.x; wfq[0][3] = kl.w * fb.y;
      }
      if (NSTAGE > 1) {  // parked for the later stages, de-interleaved by column parity
        sts2(sm_f + fso + pa_off, wfq[0][0], wfq[0][2]);
        sts2(sm_f + fso + pb_off, wfq[0][1], wfq[0][3]);
      }
    }
#pragma unroll
    for (int k = 1; k < NSTAGE; ++k) {  // w f of row t - LAG k
      const bool is_res = RESTRICT && (k == NSTAGE - 1);
      const bool gs_stage = (GS != 0) && !is_res;
      const int ci = k & 3, pr = ci & 1, pc = (ci == 1 || ci == 2) ? 1 : 0;
      if (gs_stage && pr != ((ODD ? 1 : 0) ^ ((LAG * k) & 1))) continue;  // this stage opens no point in a row of this parity
      const unsigned sl = ring_back<kFR>(fso, (unsigned)(LAG * k) * 1024u);
      if (!gs_stage || pc == 0) { const double2 g0 = lds2(sm_f + sl + pa_off); wfq[k][0] = g0.x; wfq[k][2] = g0.y; }
      if (!gs_stage || pc == 1) { const double2 g1 = lds2(sm_f + sl + pb_off); wfq[k][1] = g1.x; wfq[k][3] = g1.y; }
    }
    if (PROLONG) {
      if (!ODD) {
        const double2 e2 = ring_e[(((t >> 1) + cs) & (k9ERing - 1)) * 32 + lane];
        const double eleft = __shfl_up_sync(0xffffffffu, e2.y, 1);
#pragma unroll
        for (int q = 0; q < C; ++q) eprev[q] = ecur[q];
        ecur[0] = hm * (eleft + e2.x);
        ecur[1] = e2.x;
        ecur[2] = hm * (e2.x + e2.y);
        ecur[3] = e2.y;
#pragma unroll
        for (int q = 0; q < C; ++q) x0[q] += 0.5 * (eprev[q] + ecur[q]);
      } else {
#pragma unroll
        for (int q = 0; q < C; ++q) x0[q] += ecur[q];
      }
      if (SLOW) {
        const bool rin = (unsigned)(t + L.row0) < (unsigned)nglob;
#pragma unroll
        for (int q = 0; q < C; ++q) x0[q] = rin ? x0[q] : 0.0;
      }
    }

    // ---- stages.  LAG 2: last one first, a stage reads what its predecessor finished in the PREVIOUS step;
    // LAG 1: first one first, the finished row is handed on within the step ---------------------------------------
    double res[C];
    double xflow[C];
#pragma unroll
    for (int q = 0; q < C; ++q) xflow[q] = x0[q];
#pragma unroll
    for (int kk = 0; kk < NSTAGE; ++kk) {
      const int k = (LAG == 2) ? NSTAGE - 1 - kk : kk;
      const int n = t - LAG * k;  // arriving row; n - 1 is finished
      const bool is_res = RESTRICT && (k == NSTAGE - 1);
      const bool gs_stage = (GS != 0) && !is_res;
      const int ci = k & 3, pr = ci & 1, pc = (ci == 1 || ci == 2) ? 1 : 0;   // colour order (0,0) (1,1) (0,1) (1,0)
      const int pn = (ODD ? 1 : 0) ^ ((LAG * k) & 1);
      double x[C];
#pragma unroll
      for (int q = 0; q < C; ++q) x[q] = (LAG == 2) ? ((k == 0) ? x0[q] : ready[k - 1][q]) : xflow[q];
      const bool open_here = !gs_stage || (pn == pr);    // row n holds points of this stage's colour
      const bool finish_here = !gs_stage || (pn != pr);  // row n - 1 does
      const double xl = __shfl_up_sync(0xffffffffu, x[C - 1], 1);
      const double xr = __shfl_down_sync(0xffffffffu, x[0], 1);
      // warp-uniform row tests of a boundary step: row n - 1 outside the grid; row n / n + 1 / n - 1 the last grid row.
      // The stage body exists twice in a boundary step: without any last-row code (taken by all but at most three
      // stages of a step) and with it; the choice is one uniform branch per stage.
      const bool up_out = SLOW && (n - 1 < go || n - 1 > gl);
      auto body = [&](auto last_tag) {
      constexpr bool LASTROW = decltype(last_tag)::value;
      const bool open_last = LASTROW && (n == gl), down_last = LASTROW && (n + 1 == gl), up_last = LASTROW && (n - 1 == gl);
#pragma unroll
      for (int q = 0; q < C; ++q) {
        const bool colq = !gs_stage || ((q & 1) == pc);
        const double h = ((q == 0) ? xl : x[q - 1]) + ((q == C - 1) ? xr : x[q + 1]);
        double rdy = gs_stage ? xc[k][q] : 0.0;  // Gauss-Seidel: points of other colours pass through
        const double up_old = uprev[k][q];       // what row n - 1 gives row n
        if (colq && finish_here) {
          const C9 &cu = cls0(q);
          const double un_up = fma(cu.c1, h, cu.c2 * x[q]);
          rdy = pre1[k][q] + un_up;
          if (down_last) {   // what row n gives the last row
            const C9 cd = cls1(q);
            uprev[k][q] = fma(cd.c1, h, cd.c2 * x[q]);
          } else {
            uprev[k][q] = un_up;
          }
        }
        if (colq && open_here) {
          C9 cm = cls0(q);
          if (open_last) cm = cls1(q);
          const double a = is_res ? cm.ares : cm.asm_;
          double base;
          if (is_res || k < (GS ? 4 : 1)) base = fma(a, x[q], fma(cm.dlo * (is_res ? 1.0 : (GS ? NU / 4 : NU)), x[q], wfq[k][q]));
          else base = fma(a, x[q], wfq[k][q]);
          pre1[k][q] = fma(cm.c3, h, base) + up_old;
        }
        if (up_out) rdy = 0.0;
        ready[k][q] = rdy;
        xc[k][q] = x[q];
        if (LAG == 1) xflow[q] = rdy;
      }
      const int rho = n - 1;
      if (!is_res && k == NU - 1 && (unsigned)(rho - r0) < out_rows && quadout) {
        double *dst = pout;
        if (st32) st4(dst, ready[k][0], ready[k][1], ready[k][2], ready[k][3]);
        else { st_stream2(dst, make_double2(ready[k][0], ready[k][1])); st_stream2(dst + 2, make_double2(ready[k][2], ready[k][3])); }
      }
      if (is_res) {
#pragma unroll
        for (int q = 0; q < C; ++q)
          res[q] = up_last ? ready[k][q] * cls1(q).rs : ((q == C - 1) ? ready[k][q] * kl.rs : ready[k][q]);
      }
      };  // body
      if (SLOW && (unsigned)(n - gl + 1) <= 2u) body(TrueT{});   // n - 1, n or n + 1 is the last grid row
      else body(FalseT{});
    }
    if (RESTRICT) {
      // res = w * residual of row rho = t - LAG NU - 1 (zero outside the grid): full weighting, columns first
      const int rho = t - LAG * NU - 1;
      constexpr bool RHO_ODD = (ODD != (((LAG * NU + 1) & 1) != 0));
      const double rnext = __shfl_down_sync(0xffffffffu, res[0], 1);
      double crr[2];
      crr[0] = fma(q4, res[0] + res[2], q2 * res[1]);
      crr[1] = fma(q4, res[2] + rnext, q2 * res[3]);
      if (!RHO_ODD) {
        const int I = (rho >> 1) - 1 + cs;
        const int G = ((rho + L.row0) >> 1) - 1;
        const bool rowok = (I >= (r0 >> 1) + cs && I < (r1 >> 1) + cs && I >= 0 && I < nrc && G >= 0 && G < (nglob >> 1));
        if (rowok && quadout)
          st_stream2(r_coarse + (size_t)I * ncc + (c0 >> 1), make_double2(fma(0.25, crr[0], racc[0]), fma(0.25, crr[1], racc[1])));
        racc[0] = 0.25 * crr[0];
        racc[1] = 0.25 * crr[1];
      } else {
        racc[0] = fma(0.5, crr[0], racc[0]);
        racc[1] = fma(0.5, crr[1], racc[1]);
      }
    }
    if (NU == 0 && !RESTRICT && t >= r0 && t < r1 && quadout) {
      double *dst = v_out + (size_t)t * L.ncols + c0;
      if (st32) st4(dst, x0[0], x0[1], x0[2], x0[3]);
      else { st_stream2(dst, make_double2(x0[0], x0[1])); st_stream2(dst + 2, make_double2(x0[2], x0[3])); }
    }
    vso = ring_fwd<kVR>(vso, 1024);
    fso = ring_fwd<kFR>(fso, 1024);
    if (NU > 0) pout += L.ncols;
  };

#pragma unroll
  for (int d = 0; d < AHEAD; ++d) issue(t_begin + d, d * 1024, d * 1024);

  for (int t = t_begin; t <= t_last; t += 2) {
    // rows t - LAG (NSTAGE - 1) - 1 .. t + 2 (global) all interior?
    const int g = t + L.row0;
    const bool slow = (g - LAG * NSTAGE - 1 < 0) || (g + 2 >= nglob - 1);
    if (!slow) {
      step(t, FalseT{}, FalseT{});
      step(t + 1, TrueT{}, FalseT{});
    } else {
      step(t, FalseT{}, TrueT{});
      step(t + 1, TrueT{}, TrueT{});
    }
  }
  cpa_wait<0>();
}

// ---------------------------------------------------------------------------------------------------
namespace {

// ---- double-double helpers for the consistent interior weights ----------------------------------------------
struct dd { double hi, lo; };
dd two_sum(double a, double b) { const double s = a + b, bb = s - a; return {s, (a - (s - bb)) + (b - bb)}; }
dd two_prod(double a, double b) { const double p = a * b; return {p, fma(a, b, -p)}; }
dd dd_add(dd a, dd b) { dd s = two_sum(a.hi, b.hi); const double lo = s.lo + a.lo + b.lo; return two_sum(s.hi, lo); }
dd dd_mul(dd a, dd b) { dd p = two_prod(a.hi, b.hi); const double lo = p.lo + (a.hi * b.lo + a.lo * b.hi); return two_sum(p.hi, lo); }
dd dd_scale(dd a, double s) { return dd_mul(a, {s, 0.0}); }

K9 u9_coef(const LevelDev &L, double shift, double omega) {
  K9 K;
  // classes: [0] interior / [1] last diagonal entry
  const double kad[2] = {L.u9[1], L.u9[4]}, mad[2] = {L.u9[3], L.u9[5]};
  const double kbd[2] = {L.u9[7], L.u9[10]}, mbd[2] = {L.u9[9], L.u9[11]};
  const double ko = L.u9[0], mo = L.u9[2], kbo = L.u9[6], mbo = L.u9[8];
  const double cc = mo * kbo + ko * mbo;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 2; ++c) {
      C9 &z = K.c[r][c];
      if (r == 2) { z.c1 = z.c2 = z.c3 = z.ares = z.asm_ = z.dlo = z.w = z.rs = 0.0; continue; }
      const double d = (mad[r] * kbd[c] + kad[r] * mbd[c]) - shift;
      const double w = omega / d;
      const double cv = mo * kbd[c] + ko * mbd[c];
      const double ch = mad[r] * kbo + kad[r] * mbo;
      z.w = w;
      z.c1 = -w * cc;
      z.c2 = -w * cv;
      z.c3 = -w * ch;
      z.ares = -w * d;
      z.asm_ = 1.0 + z.ares;
      z.dlo = 0.0;
      z.rs = 1.0;
    }
  // interior: x coefficient = -w * rowsum - (4 c1 + 2 c2 + 2 c3) in double-double, so that the stencil the rounded
  // constants define has the exact row sum (-w sigma): the smooth-mode action of the level operator is not shifted
  {
    C9 &z = K.c[0][0];
    const dd sa = dd_add(two_sum(mad[0], 2.0 * mo), {0.0, 0.0}), sb = two_sum(kbd[0], 2.0 * kbo);
    const dd sc = two_sum(kad[0], 2.0 * ko), sd = two_sum(mbd[0], 2.0 * mbo);
    dd sigma = dd_add(dd_mul(sa, sb), dd_mul(sc, sd));
    sigma = dd_add(sigma, {-shift, 0.0});
    dd target = dd_scale(sigma, -z.w);
    dd rest = dd_add(dd_add(two_prod(4.0, z.c1), two_prod(2.0, z.c2)), two_prod(2.0, z.c3));
    dd c4 = dd_add(target, {-rest.hi, -rest.lo});
    if (fabs(c4.hi + 1.0) < 1e-9) {  // a weight a few ulps from -1 must not multiply x (see fused_uni.cu)
      c4.lo += c4.hi + 1.0;
      c4.hi = -1.0;
    }
    z.ares = c4.hi;
    z.asm_ = 1.0 + c4.hi;
    z.dlo = c4.lo + ((z.asm_ - 1.0) - c4.hi) * -1.0;
  }
  K.invw = 1.0 / K.c[0][0].w;
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) K.c[r][c].rs = K.c[0][0].w / K.c[r][c].w;
  return K;
}

size_t u9_smem_bytes(bool prolong, bool zerov, int nstage, int lag) {
  const size_t gran = (size_t)((zerov ? 0 : k9VR) + u9_fring(nstage, lag)) * 64 + (prolong ? k9ERing * 32 : 0);
  return gran * 16 * k9Warps;
}

}  // namespace

int g_uni9_lag = 1;

template <int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int GS, int LAG>
static cudaError_t launch_u9_l(const LevelDev &L, double shift, double omega, const double *v_in, const double *f,
                               double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s) {
  constexpr int NSTAGE = NU + (RESTRICT ? 1 : 0);
  constexpr int USEFUL = 32 * k9C - 2 * u9_halo(NU);
  auto kern = uni9_leg_kernel<NU, PROLONG, RESTRICT, ZEROV, GS, LAG>;
  const size_t smem = u9_smem_bytes(PROLONG, ZEROV, NSTAGE, LAG);
  static int occ = 0;
  if (!occ) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, k9Warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
  }
  const int strips = (L.ncols + USEFUL - 1) / USEFUL;
  const int gx = (strips + k9Warps - 1) / k9Warps;
  const int rpc = leg_rows_per_chunk(L.nrows, gx, occ * num_sms(), LAG * NSTAGE, 1 << 20);
  dim3 grid(gx, (L.nrows + rpc - 1) / rpc);
  kern<<<grid, k9Warps * 32, smem, s>>>(L, u9_coef(L, shift, omega), v_in, f, v_out, e_coarse, r_coarse, rpc);
  count_launch();
  return cudaGetLastError();
}

template <int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int GS>
static cudaError_t launch_u9_t(const LevelDev &L, double shift, double omega, const double *v_in, const double *f,
                               double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s) {
  if (g_uni9_lag == 2) return launch_u9_l<NU, PROLONG, RESTRICT, ZEROV, GS, 2>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  return launch_u9_l<NU, PROLONG, RESTRICT, ZEROV, GS, 1>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
}

template <int NU, int GS>
static cudaError_t u9_dispatch_mode(const LevelDev &L, int mode, double shift, double omega, const double *v_in,
                                    const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s) {
  switch (mode) {
    case FUSED_SMOOTH:
      if (NU == 0) return cudaErrorInvalidValue;
      return launch_u9_t<NU, false, false, false, GS>(L, shift, omega, v_in, f, v_out, nullptr, nullptr, s);
    case FUSED_DOWN:
      return launch_u9_t<NU, false, true, false, GS>(L, shift, omega, v_in, f, v_out, nullptr, r_coarse, s);
    case FUSED_DOWN_ZERO:
      return launch_u9_t<NU, false, true, true, GS>(L, shift, omega, v_in, f, v_out, nullptr, r_coarse, s);
    case FUSED_UP:
      return launch_u9_t<NU, true, false, false, GS>(L, shift, omega, v_in, f, v_out, e_coarse, nullptr, s);
  }
  return cudaErrorInvalidValue;
}

// 0: never; 1: every eligible 9-point level; 2 (default): levels at least 2048 wide (slab pieces included) -- the only
// place where these legs beat the general kernel on B200 (8 % at 2048^2)
int g_fused_uni9 = 2;
int g_uni9_min_cols = 512;   // measured in the 4-stream bench step (what counts is SM time, not the latency of a lone leg): 2048: 2.25 ms, 1024: 2.14, 512: 2.11, all levels: 2.23

bool uni9_available(const LevelDev &L) {
  if (!g_fused_uni9 || L.uni != 2 || L.five || L.nrows < 16 || L.ncols < 16 || (L.ncols & 3)) return false;
  if (g_fused_uni9 == 2) {
    // slab pieces (lock-step multi-GPU step, few rows per rank on the narrow levels): 2048 measured better than 512
    const bool slab_piece = L.row0 != 0 || L.nrows != L.nrows_glob;
    return L.ncols >= (slab_piece && g_uni9_min_cols < 2048 ? 2048 : g_uni9_min_cols);
  }
  return true;
}

// nu = Jacobi sweeps 0..4 (gs = 0) or four-colour sweeps 1..2 (gs = 1)
cudaError_t launch_uni9_leg(const LevelDev &L, int gs, int mode, int nu, double shift, double omega, const double *v_in,
                            const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s) {
  if (!uni9_available(L)) return cudaErrorInvalidValue;
  if (nu == 0) gs = 0;
#define U9_CASE(NUV, STAGES, GSV) \
  case NUV: return u9_dispatch_mode<STAGES, GSV>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s);
  if (gs) {
    switch (nu) { U9_CASE(1, 4, 1) U9_CASE(2, 8, 1) }
  } else {
    switch (nu) { U9_CASE(0, 0, 0) U9_CASE(1, 1, 0) U9_CASE(2, 2, 0) U9_CASE(3, 3, 0) U9_CASE(4, 4, 0) }
  }
#undef U9_CASE
  return cudaErrorInvalidValue;
}

}  // namespace mgcmt
