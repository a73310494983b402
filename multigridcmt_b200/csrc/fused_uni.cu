// fused_uni.cu -- the V-cycle legs of fused.cu for levels whose operator is a CONSTANT 5-point stencil
// (the infinite well of 2DPot*.py on its finest grid: H = (-1/pi^2) laplacian(N, "2d"), MGCMTStencilMaker.py:15-25).
//
//   down leg:  v_out = S^NU(v_in; f),   r_coarse = R (f - A_s v_out)        (MGCMTSolver.py:313-315)
//   up leg:    v_out = S^NU(v_in + P e_coarse; f)  [+ w^T A_s w, w^T w]     (MGCMTSolver.py:323-326, 2DPotGS.py:103)
//
// One HBM pass per leg as in fused.cu (rows streamed through a cp.async ring, the sweeps pipelined along the row
// direction, strips overlapped by HALO columns and recomputed), rebuilt around what ncu showed limits that kernel --
// not fp64 and not HBM but the shared-memory / shuffle path (MIO) and the in-order issue behind it:
//
//  * arithmetic: 5 instead of ~10 fp64 instructions per update, using what is constant:
//      A_s x (i,j) = c (x(i-1,j) + x(i+1,j) + x(i,j-1) + x(i,j+1)) + d x(i,j),        d = ka_di + kb_di - shift
//      sweep:      out = (1 - omega) x + w f - beta S4,     w = omega / d,  beta = w c,  S4 = sum of the 4 neighbours
//      residual:   w r = - omega x + w f - beta S4          (the 1/w goes into the full-weighting constants)
//      Rayleigh:   w A_s x = omega x + beta S4              (the residual stage with f = 0; 1/w applied to the warp sum)
//    Per stage and column the state is the centre row `xc` and `pre` = everything of the open row known one step before
//    it is finalised; when the lower neighbour row arrives the finalised value is ONE fma, out = fma(-beta, x, pre).
//  * a step runs in two phases: phase A is the chain of NSTAGE dependent fmas (plus the shuffle-free part of the new
//    `pre`) and fires the two shuffles per stage; phase B consumes the shuffled neighbours.  A warp issues in order, so
//    this keeps NSTAGE shuffle round trips off the critical path of a step.
//  * global <-> shared: lane l copies the 16-byte granules l and 32+l of the strip row (every cp.async instruction
//    covers 512 contiguous bytes = whole 32-byte sectors; the per-lane column pairs of fused.cu touch half of every
//    sector per instruction, which doubles the L2 -> SM sectors and the shared-memory write wavefronts), the granules
//    are XOR-swizzled in shared memory (bit 0 ^= bit 3) so that both the copies and the consumers' LDS.128 of their own
//    4 columns are bank-conflict free, and one __syncwarp per step makes the copies visible across lanes.  Results leave
//    with one 32-byte store per lane (HALO is a multiple of 4 columns).
//  * w f: Jacobi legs carry it from stage to stage in registers (WFREG: no shared-memory round trip per stage);
//    Gauss-Seidel legs (9 stages) park it in the f ring slot, de-interleaved by column parity so that a colour stage
//    reads one granule.
// Dirichlet zeros: beta = 0 in columns outside the grid (per-lane constant) keeps them zero; rows outside the grid
// are only met in the first / last few steps of the first / last chunk, which run a masked copy of the step (SLOW).
// Weighted Jacobi and the red-black Gauss-Seidel/SOR colour stages (GS = 1: two stages per sweep) share the code; a
// colour stage only generates the arithmetic of its colour (row parity is compile-time through the unrolled step,
// column parity through the first column of every lane being a multiple of 4).
//
// Results differ from fused.cu / the oracle only in the association of the sums (tested at 1e-12).
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace mgcmt {

namespace {

constexpr int kC = 4;        // columns per lane
constexpr int kWarpsU = 4;   // warps per CTA (independent strips)
constexpr int kERingU = 4;   // coarse-row ring (PROLONG)
__host__ __device__ constexpr int uni_halo(int nu) { return (nu + 2 + 3) & ~3; }   // >= nu + 2, multiple of 4
__host__ __device__ constexpr int uni_vring(int nstage) { return nstage > 5 ? 4 : 6; }
// rows of the w f ring; when the later stages read it (no WFREG) it is rounded up to a power of two, so that a stage's
// slot is one add and one mask away from the newest row's
__host__ __device__ constexpr int uni_fring(int nstage, bool wfreg) {
  if (wfreg) return uni_vring(nstage);
  const int need = uni_vring(nstage) + (nstage > 0 ? nstage - 1 : 0);
  int p = 1;
  while (p < need) p <<= 1;
  return p;
}

__device__ __forceinline__ void ucpa16(void *smem, const void *gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ucpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ucpa_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// ---- bulk asynchronous copies (the TMA engine's 1-D form) with mbarrier completion: the BULK variant of the row ring ----
__device__ __forceinline__ void mbar_init(void *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 26) && !done; ++spin) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
  if (!done) __trap();  // a lost copy must fail the launch, not hang the GPU
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void st_stream4(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

struct UStage {
  double xc[kC], pre[kC];
};

// The coefficients of the sweep  out = a x + wf f - beta S4  (a = 1 - om, beta = om c / ds, wf = om / ds, ds = d - shift).
// Rounding a and beta independently perturbs the ratio om : beta, which is the DIAGONAL of the operator that the
// sweeps, the residual and the Rayleigh stage effectively use: a relative 1e-16 there is an absolute ~1e-16 * d ~ 7e-10
// shift of the spectrum at 4096^2 (d = 4 N^2 / pi^2), systematic over all points, so it does not average out like data
// rounding does, and a cycle amplifies it by 1 / |lambda - shift|.  The reference (and the general kernels) never
// round the diagonal against the off-diagonals: they form c S4 + d x - shift x in the data.  Here beta = fl(omega c / ds)
// is taken as given and the matching weight om_e = beta ds / c is carried as a double-double hi + lo (ds = d - shift
// exactly, as hi + lo too): a = 1 - hi and -hi are exact, and the missing -lo x (|lo| < 6e-17) is added with one extra
// fma where it matters -- in the residual / Rayleigh stage, and once per leg for the sweeps (NU * lo in the first
// update of a point: the components it matters for are the smooth ones, which the sweeps barely change).
struct UniCoef {
  double a_smooth;  // 1 - om_hi
  double a_res;     // -om_hi
  double dlo;       // -om_lo
  double nbeta;     // -beta
  double wf;        // scale of f: beta / c
  double invw;      // c / beta
  double drem;      // (d + 4 c) - shift: what is left of the diagonal next to c (S4 - 4 x)
};

UniCoef uni_coef(double c, double du, double shift, double omega) {
  const double hi = du - shift;
  const double bb = hi - du;
  const double lo = (du - (hi - bb)) + (-shift - bb);  // TwoSum: hi + lo == du - shift exactly
  const double beta = omega * c / hi;
  // om_e = beta (hi + lo) / c as q_hi + q_lo
  const double p_hi = beta * hi, p_lo = fma(beta, hi, -p_hi) + beta * lo;
  double q_hi = p_hi / c;
  double q_lo = (fma(-q_hi, c, p_hi) + p_lo) / c;
  // A weight a few ulps from 1 (Gauss-Seidel, omega = 1) must not multiply x: fl((1 + 2e-16) x) is x or its neighbour
  // depending on the mantissa of x alone, a rounding error with a non-zero mean.  Split it as 1 + (om_e - 1): the
  // product with 1 is exact and the remainder joins the low part.
  if (fabs(q_hi - 1.0) < 1e-9) {
    q_lo += q_hi - 1.0;
    q_hi = 1.0;
  }
  UniCoef k;
  k.a_smooth = 1.0 - q_hi;  // exact for q_hi in [1/2, 1]; otherwise the error joins dlo
  k.a_res = -q_hi;
  k.dlo = -q_lo + ((1.0 - k.a_smooth) - q_hi);
  k.nbeta = -beta;
  k.wf = beta / c;
  k.invw = c / beta;
  k.drem = (du + 4.0 * c) - shift;
  return k;
}

}  // namespace

// NU: Jacobi sweeps (GS = 0) or colour stages (GS = 1, two per sweep).  PROLONG && RESTRICT = up leg + Rayleigh sums.
// BULK: the row ring is filled by cp.async.bulk (one elected lane per warp and row, completion on an mbarrier per ring
// slot, linear rows) instead of per-lane cp.async with the XOR swizzle -- the TMA A/B of DESIGN.md section 3.
template <int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int GS, bool WFREG, int MINCTAS, bool BULK = false>
__global__ void __launch_bounds__(kWarpsU * 32, MINCTAS)
uni5_leg_kernel(LevelDev L, UniCoef K, const double *__restrict__ v_in,
                const double *__restrict__ f, double *__restrict__ v_out,
                const double *__restrict__ e_coarse, double *__restrict__ r_coarse, int rows_per_chunk) {
  constexpr int C = kC;
  constexpr int HALO = uni_halo(NU);
  constexpr int WCOLS = 32 * C;
  constexpr int USEFUL = WCOLS - 2 * HALO;
  constexpr int NSTAGE = NU + (RESTRICT ? 1 : 0);
  constexpr int NS1 = NSTAGE > 0 ? NSTAGE : 1;
  constexpr int kVR = uni_vring(NSTAGE);
  constexpr int kFR = uni_fring(NSTAGE, WFREG);
  constexpr int AHEAD = kVR - 1;  // rows in flight ahead of the one being consumed
  constexpr bool RQ = PROLONG && RESTRICT;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  // per warp: v ring [kVR][64 granules], f ring [kFR][64], e ring [kERingU][32]; a granule is 16 bytes
  constexpr int WARP_GRAN = (ZEROV ? 0 : kVR * 64) + kFR * 64 + (PROLONG ? kERingU * 32 : 0) + (BULK ? 4 : 0);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  double2 *ring_v = reinterpret_cast<double2 *>(smem_raw) + warp * WARP_GRAN;
  double2 *ring_f = ring_v + (ZEROV ? 0 : kVR * 64);
  double2 *ring_e = ring_f + kFR * 64;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring_e + (PROLONG ? kERingU * 32 : 0));  // BULK: kVR (<= 8) barriers

  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(r0 + rows_per_chunk, L.nrows);
  const int t_first = r0 - NU - (RESTRICT ? 1 : 0);
  const int t_last = r1 - 1 + NU + (RESTRICT ? 2 : 0);  // last input row that matters = last time step
  const int t_begin = (t_first - (PROLONG ? 2 : 0)) & ~1;
  const int nrc = L.nrows_coarse ? L.nrows_coarse : L.nrows / 2, ncc = L.ncols / 2;
  const int cs = L.crow_shift;
  const unsigned nglob = (unsigned)L.nrows_glob;

  const int rq_slot = (blockIdx.y * gridDim.x + blockIdx.x) * kWarpsU + warp;
  const int rq_nslots = gridDim.x * gridDim.y * kWarpsU;
  const int strip = blockIdx.x * kWarpsU + warp;
  const int u0 = strip * USEFUL;  // first useful fine column of this strip (a multiple of 4)
  if (u0 >= L.ncols) {            // surplus warp (the kernel has no CTA-wide barrier)
    if (RQ && lane == 0) { r_coarse[rq_slot] = 0.0; r_coarse[rq_nslots + rq_slot] = 0.0; }
    return;
  }
  const int u1 = min(u0 + USEFUL, L.ncols);
  const int cstart = u0 - HALO;         // first column of the strip
  const int c0 = cstart + C * lane;     // first of this lane's 4 columns (a multiple of 4)

  const double w = K.wf;
  const double a_smooth = K.a_smooth, a_res = K.a_res;
  const double dres = K.dlo;                                   // low part of the x coefficient, residual / Rayleigh stage
  const double dlump = K.dlo * (GS ? NU / 2 : NU);             // ... of all sweeps of the leg, applied with the first update
  // c0 and ncols are multiples of 4: a lane's 4 columns are inside / outside the grid (and useful or not) together
  const bool quadin = (c0 >= 0 && c0 < L.ncols);
  const bool quadout = (c0 >= u0 && c0 < u1);
  const double nb = quadin ? K.nbeta : 0.0;          // -beta inside the grid, 0 outside
  const double hm = quadin ? 0.5 : 0.0;              // interpolation weight of even columns: 0 outside the grid
  const double q4 = 0.25 * K.invw, q2 = 0.5 * K.invw;  // full weighting with the 1/w of the scaled residual folded in
  const bool st32 = ((reinterpret_cast<uintptr_t>(v_out) & 31) == 0);

  // Ring addressing: 32-bit shared-space byte offsets, one ring row = 64 granules = 1024 bytes (common.cuh).
  // loader: granules lane and 32 + lane of the strip row (512 bytes apart); swizzled position G ^ ((G >> 3) & 1)
  const unsigned sm_v = (unsigned)__cvta_generic_to_shared(ring_v), sm_f = (unsigned)__cvta_generic_to_shared(ring_f);
  bool ldin[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int j = cstart + 2 * (32 * g + lane);
    ldin[g] = (j >= 0 && j < L.ncols);
  }
  const unsigned ld_off = (unsigned)(BULK ? lane : (lane ^ ((lane >> 3) & 1))) * 16;
  // consumer: granules 2 lane, 2 lane + 1 (columns c0 .. c0+3)
  const unsigned pa_off = (unsigned)(BULK ? 2 * lane : ((2 * lane) ^ ((lane >> 2) & 1))) * 16, pb_off = pa_off ^ 16;
  // rows the copies may touch: inside the array, inside the grid, not beyond what the last step needs
  const int t_lo = max(0, -L.row0);
  const int t_span = max(0, min(min(L.nrows, t_last + 1), (int)nglob - L.row0) - t_lo);
  // running source pointers of this lane's first granule in the row the next issue() fetches (a predicated-off copy reads
  // nothing, so rows before / after the array need no clamping), and the destination of the row the last sweep finishes
  const double *pv = ZEROV ? nullptr : v_in + ((ptrdiff_t)t_begin * L.ncols + cstart + 2 * lane);
  const double *pf = f + ((ptrdiff_t)t_begin * L.ncols + cstart + 2 * lane);
  double *pout = v_out + ((ptrdiff_t)(t_begin - NU) * L.ncols + c0);
  const unsigned out_rows = (unsigned)(r1 - r0);
  // BULK: the part of the strip row that lies inside the grid is one contiguous copy; the granules outside stay zero
  const int bj0 = max(cstart, 0), bj1 = min(cstart + WCOLS, L.ncols);
  const unsigned brow_bytes = (unsigned)(bj1 - bj0) * 8u;
  const int bJ0 = bj0 >> 1, bJ1 = min((cstart + WCOLS) >> 1, ncc);
  if (BULK) {
    if (lane == 0)
      for (int i = 0; i < kVR; ++i) mbar_init(bars + i, 1);
    for (int i = lane; i < WARP_GRAN - 4; i += 32) ring_v[i] = make_double2(0.0, 0.0);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_proxy();
    __syncwarp();
  }

  // ---- asynchronous row fetch ---------------------------------------------------------------------
  auto issue = [&](int t, unsigned vo, unsigned fo) {   // vo / fo: byte offsets of the ring rows to fill
    // rows outside the slab array or outside the global grid are zero-filled
    const bool rowin = (unsigned)(t - t_lo) < (unsigned)t_span;
    if (BULK) {
      const int vslot = (int)(vo >> 10), fslot = (int)(fo >> 10);
      const size_t rowoff = (size_t)(rowin ? t : 0) * L.ncols;
      // every lane first clears its granules of a slot that gets no data (row outside the grid), then lane 0 arms the
      // slot's barrier with the bytes to come and issues the copies
      bool ein = false;
      int I = 0;
      if (PROLONG && (t & 1) == 0) {
        I = (t >> 1) + cs;
        const int Gc = ((t + L.row0) >> 1);
        ein = I >= 0 && I < nrc && t <= t_last && Gc >= 0 && Gc < (L.nrows_glob >> 1) && bJ1 > bJ0;
        if (!ein) ring_e[(I & (kERingU - 1)) * 32 + lane] = make_double2(0.0, 0.0);
      }
      if (!rowin) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (!ZEROV) ring_v[vslot * 64 + 32 * g + lane] = make_double2(0.0, 0.0);
          if (NSTAGE > 0) ring_f[fslot * 64 + 32 * g + lane] = make_double2(0.0, 0.0);
        }
      }
      // generic-proxy writes into ring slots (the zero fills above; w f parked by the Gauss-Seidel stage 0) must be
      // ordered before the async-proxy copy that reuses the slot
      if ((!WFREG && NSTAGE > 1) || !rowin || (PROLONG && (t & 1) == 0 && !ein)) fence_async_proxy();
      __syncwarp();
      if (lane == 0) {
        const bool any = rowin && brow_bytes > 0;
        const unsigned bytes = (any ? brow_bytes * ((ZEROV ? 0u : 1u) + (NSTAGE > 0 ? 1u : 0u)) : 0u) + (ein ? (unsigned)(bJ1 - bJ0) * 8u : 0u);
        mbar_expect_tx(bars + vslot, bytes);
        if (any) {
          if (!ZEROV) bulk_g2s(reinterpret_cast<double *>(ring_v + vslot * 64) + (bj0 - cstart), v_in + rowoff + bj0, brow_bytes, bars + vslot);
          if (NSTAGE > 0) bulk_g2s(reinterpret_cast<double *>(ring_f + fslot * 64) + (bj0 - cstart), f + rowoff + bj0, brow_bytes, bars + vslot);
        }
        if (ein)
          bulk_g2s(reinterpret_cast<double *>(ring_e + (I & (kERingU - 1)) * 32) + (bJ0 - (cstart >> 1)), e_coarse + (size_t)I * ncc + bJ0,
                   (unsigned)(bJ1 - bJ0) * 8u, bars + vslot);
      }
      return;
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const bool ok = rowin && ldin[g];
      if (!ZEROV) cpa16s(sm_v + vo + ld_off + 512 * g, pv + 64 * g, ok);
      if (NSTAGE > 0) cpa16s(sm_f + fo + ld_off + 512 * g, pf + 64 * g, ok);
    }
    if (!ZEROV) pv += L.ncols;
    pf += L.ncols;
    if (PROLONG && (t & 1) == 0) {
      // coarse row I = t/2 is first needed by fine row t (even); this lane's coarse columns c0/2, c0/2 + 1
      const int I = (t >> 1) + cs;
      const int Gc = ((t + L.row0) >> 1);  // global coarse row
      const int J = c0 >> 1;
      const bool ok = I >= 0 && I < nrc && t <= t_last && Gc >= 0 && Gc < (L.nrows_glob >> 1) && J >= 0 && J < ncc;
      ucpa16(ring_e + (I & (kERingU - 1)) * 32 + lane, e_coarse + (ok ? (size_t)I * ncc + J : 0), ok);
    }
    ucpa_commit();
  };

  UStage st[NS1];
  double wfq[NS1][C];  // WFREG: w f of the row each stage opens in the current step
#pragma unroll
  for (int k = 0; k < NS1; ++k)
#pragma unroll
    for (int q = 0; q < C; ++q) st[k].xc[q] = st[k].pre[q] = wfq[k][q] = 0.0;

  double eprev[C], ecur[C];  // column-interpolated coarse rows I-1 and I (PROLONG)
#pragma unroll
  for (int q = 0; q < C; ++q) eprev[q] = ecur[q] = 0.0;
  double rq_num[C] = {0.0, 0.0, 0.0, 0.0}, rq_den[C] = {0.0, 0.0, 0.0, 0.0};  // one accumulator per column: no serial fma chain
  double racc[2] = {0.0, 0.0};  // running full-weighting row sums (RESTRICT)

  unsigned vso = 0, fso = 0;  // byte offsets of the ring rows that hold the current input row t
  auto step = [&](int t, auto odd_tag, auto slow_tag) {
    constexpr bool ODD = decltype(odd_tag)::value;
    constexpr bool SLOW = decltype(slow_tag)::value;
    if (BULK) {
      mbar_wait(bars + (vso >> 10), (unsigned)(((t - t_begin) / kVR) & 1));  // row t (and its coarse row) has landed
    } else {
      ucpa_wait<AHEAD - 1>();  // row t has landed (this lane's copies) ...
    }
    __syncwarp();              // ... and every lane's; all lanes are done reading the slots refilled below
    issue(t + AHEAD, ring_fwd<kVR>(vso, AHEAD * 1024), ring_fwd<kFR>(fso, AHEAD * 1024));

    // ---- the input row t ---------------------------------------------------------------------------
    double x[C];
    if (ZEROV) {
#pragma unroll
      for (int q = 0; q < C; ++q) x[q] = 0.0;
    } else {
      const double2 xa = lds2(sm_v + vso + pa_off), xb = lds2(sm_v + vso + pb_off);
      x[0] = xa.x; x[1] = xa.y; x[2] = xb.x; x[3] = xb.y;
    }
    if (NSTAGE > 0) {
      const double2 fa = lds2(sm_f + fso + pa_off), fb = lds2(sm_f + fso + pb_off);
      wfq[0][0] = w * fa.x; wfq[0][1] = w * fa.y; wfq[0][2] = w * fb.x; wfq[0][3] = w * fb.y;
      if (!WFREG && NSTAGE > 1) {
        // later stages read w f from the slot, de-interleaved by column parity: a colour stage needs one granule
        sts2(sm_f + fso + pa_off, wfq[0][0], wfq[0][2]);
        sts2(sm_f + fso + pb_off, wfq[0][1], wfq[0][3]);
      }
    }
    if (!WFREG) {
      // w f of the rows the later stages open, all loads up front (row t - k, parked by stage 0 k steps ago)
#pragma unroll
      for (int k = 1; k < NSTAGE; ++k) {
        const bool is_res = RESTRICT && (k == NSTAGE - 1);
        if (RQ && is_res) continue;
        const unsigned sl = ring_back<kFR>(fso, (unsigned)k * 1024u);
        const bool gs_stage = (GS != 0) && !is_res;
        // columns this stage opens in row n = t - k: parity (colour + row parity); row parity of n = ODD ^ (k & 1)
        const int cpar = ODD ? 1 : 0;  // colour (k & 1) ^ parity of row t - k
        if (!gs_stage || cpar == 0) { const double2 g0 = lds2(sm_f + sl + pa_off); wfq[k][0] = g0.x; wfq[k][2] = g0.y; }
        if (!gs_stage || cpar == 1) { const double2 g1 = lds2(sm_f + sl + pb_off); wfq[k][1] = g1.x; wfq[k][3] = g1.y; }
      }
    }
    if (PROLONG) {
      if (!ODD) {
        // new coarse row I = t/2: interpolate along columns.  fine col c0+2g (even) = 1/2 (E[J-1] + E[J]),
        // fine col c0+2g+1 = E[J], J = c0/2 + g.  Lane 0 has no left neighbour: its first column is the
        // outermost halo column of the strip and HALO >= NSTAGE + 1, so that error never reaches a useful column.
        const double2 e2 = ring_e[(((t >> 1) + cs) & (kERingU - 1)) * 32 + lane];
        const double eleft = __shfl_up_sync(0xffffffffu, e2.y, 1);
#pragma unroll
        for (int q = 0; q < C; ++q) eprev[q] = ecur[q];
        ecur[0] = hm * (eleft + e2.x);  // hm = 0 outside the grid: the first column past the grid stays zero
        ecur[1] = e2.x;
        ecur[2] = hm * (e2.x + e2.y);
        ecur[3] = e2.y;
#pragma unroll
        for (int q = 0; q < C; ++q) x[q] += 0.5 * (eprev[q] + ecur[q]);
      } else {
#pragma unroll
        for (int q = 0; q < C; ++q) x[q] += ecur[q];
      }
      if (SLOW) {  // the interpolated correction is the only input that is not already zero outside the grid
        const bool rin = (unsigned)(t + L.row0) < nglob;
#pragma unroll
        for (int q = 0; q < C; ++q) x[q] = rin ? x[q] : 0.0;
      }
    }

    // ---- phase A: row n = t - k of stage k's input arrives, row n - 1 is finalised (one fma per point) ----
    double xl[NS1], xr[NS1];  // shuffled edge neighbours of the row every stage opened
    double res[C];            // the last stage's output if it is the residual
#pragma unroll
    for (int k = 0; k < NSTAGE; ++k) {
      const int rho = t - k - 1;
      const bool is_res = RESTRICT && (k == NSTAGE - 1);
      const bool gs_stage = (GS != 0) && !is_res;
      const int colour = k & 1;
      const int prho = (ODD ? 1 : 0) ^ ((k + 1) & 1);  // parity of rho (t_begin, row0 even; c0 a multiple of 4)
      auto in_colour = [&](int prow, int pcol) { return !gs_stage || (((prow + pcol) & 1) == colour); };
      const double ak = is_res ? a_res : a_smooth;

      double out[C];
#pragma unroll
      for (int q = 0; q < C; ++q) {
        if (RQ && is_res) out[q] = st[k].pre[q] + x[q];  // Rayleigh stage: S4 - 4 x, formed in the data (see below)
        else out[q] = in_colour(prho, q & 1) ? fma(nb, x[q], st[k].pre[q]) : st[k].xc[q];
      }
      if (SLOW) {
        const bool rin = (unsigned)(rho + L.row0) < nglob;
#pragma unroll
        for (int q = 0; q < C; ++q) out[q] = rin ? out[q] : 0.0;
      }
      if (RQ && is_res) {
        // each useful (owned) point exactly once; branch-free: rows / lanes outside contribute zeros.  BOTH factors are
        // selected: rows in front of a chunk are computed from ring rows nobody has filled yet (whatever bits the last
        // kernel left in shared memory: after an NCCL kernel that can be a NaN pattern), and 0 * NaN is NaN
        const bool mine = rho >= r0 && rho < r1 && rho >= L.rq_lo && rho < L.rq_hi && quadout;
#pragma unroll
        for (int q = 0; q < C; ++q) {
          const double xm_ = mine ? st[k].xc[q] : 0.0;
          const double om_ = mine ? out[q] : 0.0;
          rq_num[q] = fma(xm_, om_, rq_num[q]);
          rq_den[q] = fma(xm_, xm_, rq_den[q]);
        }
      }
      // open row n: the part that needs no shuffle (old centre row = upper neighbour, a x, w f)
#pragma unroll
      for (int q = 0; q < C; ++q) {
        if (in_colour(prho ^ 1, q & 1)) {
          double base;
          if (RQ && is_res) {
            // Rayleigh stage: no rounded coefficient touches the data.  The stage sums x (S4 - 4 x) and x x; c and the
            // diagonal remainder (d + 4 c - shift) multiply the two sums once at the end.  (With the sweep's scaled
            // coefficients the sum carried a relative bias of ~1e-11 at 4096^2: products of x with omega and beta are
            // rounded with a mean that does not vanish when the coefficients are in simple ratios.)
            st[k].pre[q] = fma(-4.0, x[q], st[k].xc[q]);
            continue;
          }
          if (is_res) base = fma(ak, x[q], fma(dres, x[q], wfq[k][q]));
          else if (k < (GS ? 2 : 1)) base = fma(ak, x[q], fma(dlump, x[q], wfq[k][q]));
          else base = fma(ak, x[q], wfq[k][q]);
          st[k].pre[q] = fma(nb, st[k].xc[q], base);
        }
      }
      xl[k] = __shfl_up_sync(0xffffffffu, x[C - 1], 1);
      xr[k] = __shfl_down_sync(0xffffffffu, x[0], 1);
#pragma unroll
      for (int q = 0; q < C; ++q) { st[k].xc[q] = x[q]; x[q] = out[q]; }

      if (!is_res && k == NU - 1 && (unsigned)(rho - r0) < out_rows && quadout) {
        // x = row rho of the last sweep: the smoothed iterate
        double *dst = pout;
        if (st32) st_stream4(dst, x[0], x[1], x[2], x[3]);
        else { st_stream2(dst, make_double2(x[0], x[1])); st_stream2(dst + 2, make_double2(x[2], x[3])); }
      }
      if (is_res) {
#pragma unroll
        for (int q = 0; q < C; ++q) res[q] = x[q];
      }
    }
    double rnext = 0.0;
    if (RESTRICT && !RQ) rnext = __shfl_down_sync(0xffffffffu, res[0], 1);

    // ---- phase B: the horizontal neighbours of the opened rows ------------------------------------------
#pragma unroll
    for (int k = 0; k < NSTAGE; ++k) {
      const bool is_res = RESTRICT && (k == NSTAGE - 1);
      const bool gs_stage = (GS != 0) && !is_res;
      const int colour = k & 1;
      const int pn = (ODD ? 1 : 0) ^ (k & 1);  // parity of the opened row n = t - k
#pragma unroll
      for (int q = 0; q < C; ++q) {
        if (!gs_stage || (((pn + q) & 1) == colour)) {
          const double xm = (q == 0) ? xl[k] : st[k].xc[q - 1];
          const double xp = (q == C - 1) ? xr[k] : st[k].xc[q + 1];
          if (RQ && is_res) st[k].pre[q] += xm + xp;
          else st[k].pre[q] = fma(nb, xm + xp, st[k].pre[q]);
        }
      }
    }
    if (RESTRICT && !RQ) {
      // res = w * residual row rho (zero outside the grid), rho = t - NU - 1: full weighting, columns first
      //   coarse J = c0/2 + g  <-  1/4 r[2J] + 1/2 r[2J+1] + 1/4 r[2J+2]
      const int rho = t - NU - 1;
      constexpr bool RHO_ODD = (ODD != ((NU + 1) % 2 != 0));
      double crr[2];
      crr[0] = fma(q4, res[0] + res[2], q2 * res[1]);
      crr[1] = fma(q4, res[2] + rnext, q2 * res[3]);
      if (!RHO_ODD) {
        const int I = (rho >> 1) - 1 + cs;        // coarse row completed by this fine row (as its row 2I+2)
        const int G = ((rho + L.row0) >> 1) - 1;  // its global coarse row: rows outside the coarse grid are never written
        const bool rowok = (I >= (r0 >> 1) + cs && I < (r1 >> 1) + cs && I >= 0 && I < nrc && G >= 0 &&
                            G < (L.nrows_glob >> 1));
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
      // pure prolongation-correction pass: write the corrected iterate
      double *dst = v_out + (size_t)t * L.ncols + c0;
      if (st32) st_stream4(dst, x[0], x[1], x[2], x[3]);
      else { st_stream2(dst, make_double2(x[0], x[1])); st_stream2(dst + 2, make_double2(x[2], x[3])); }
    }
    if (WFREG) {  // w f moves on with its row
#pragma unroll
      for (int k = NSTAGE - 1; k > 0; --k)
#pragma unroll
        for (int q = 0; q < C; ++q) wfq[k][q] = wfq[k - 1][q];
    }
    vso = ring_fwd<kVR>(vso, 1024);
    fso = ring_fwd<kFR>(fso, 1024);
    if (NU > 0) pout += L.ncols;
  };

#pragma unroll
  for (int d = 0; d < AHEAD; ++d) issue(t_begin + d, d * 1024, d * 1024);

  using TrueT = std::integral_constant<bool, true>;
  using FalseT = std::integral_constant<bool, false>;
  for (int t = t_begin; t <= t_last; t += 2) {  // t_begin is even
    // rows t - NSTAGE .. t + 1 (global) all inside the grid?
    const int g = t + L.row0;
    const bool slow = (g - NSTAGE < 0) || (g + 1 >= L.nrows_glob);
    if (!slow) {
      step(t, FalseT{}, FalseT{});
      step(t + 1, TrueT{}, FalseT{});
    } else {
      step(t, FalseT{}, TrueT{});
      step(t + 1, TrueT{}, TrueT{});
    }
  }
  if (!BULK) ucpa_wait<0>();
  if (RQ) {
    // the stage summed x (w A_s x)' with (w A_s x)' = -(omega x + beta S4): undo sign and scale
    // w^T A_s w = c sum x (S4 - 4 x) + (d + 4 c - shift) sum x x
    const double b = warp_sum((rq_den[0] + rq_den[1]) + (rq_den[2] + rq_den[3]));
    const double a = fma(L.uni_c, warp_sum((rq_num[0] + rq_num[1]) + (rq_num[2] + rq_num[3])), K.drem * b);
    if (lane == 0) { r_coarse[rq_slot] = a; r_coarse[rq_nslots + rq_slot] = b; }
  }
}

// ---------------------------------------------------------------------------------------------------
namespace {

size_t uni_smem_bytes(bool prolong, bool zerov, int nstage, bool wfreg, bool bulk) {
  size_t gran = (size_t)((zerov ? 0 : uni_vring(nstage)) + uni_fring(nstage, wfreg)) * 64 + (prolong ? kERingU * 32 : 0) + (bulk ? 4 : 0);
  return gran * 16 * kWarpsU;
}

int g_num_sms = 0;

}  // namespace

int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Chunk height for a streaming leg: CTAs = gx * chunks should fill whole waves of `slots` resident CTAs (a grid a few
// CTAs over a wave costs a whole extra pass at the tail), chunks no taller than 128 rows and, while the grid allows it,
// no shorter than 32 (every chunk recomputes ~nstage rows of overlap).
int g_leg_min_rpc = 16;  // smallest chunk height (even); measured in the 4-stream bench step: 16: 2.102 ms, 8: 2.110, 32: 2.168

int leg_rows_per_chunk(int nrows, int gx, int slots, int nstage, int max_rpc) {
  if (nrows <= 16) return nrows;
  int best = 0;
  double best_cost = 1e300;
  for (int waves = 1; waves <= 8; ++waves) {
    int chunks = (int)(((long long)waves * slots) / gx);
    if (chunks < 1) continue;
    int rpc = (nrows + chunks - 1) / chunks;
    rpc = (rpc + 1) & ~1;
    if (rpc < g_leg_min_rpc) rpc = g_leg_min_rpc;
    if (rpc > max_rpc) rpc = max_rpc;
    if (rpc > nrows) rpc = nrows;
    const int nch = (nrows + rpc - 1) / rpc;
    const int w = (int)(((long long)gx * nch + slots - 1) / slots);       // waves actually needed
    const double cost = (double)w * (rpc + nstage + 2);                   // time ~ waves x rows streamed per CTA
    if (cost < best_cost - 1e-9) { best_cost = cost; best = rpc; }
  }
  return best > 0 ? best : (nrows < 128 ? nrows : 128);
}

int g_uni_minctas = 0;  // resident CTAs per SM the 4-sweep Jacobi legs are compiled for: 0 = default, 2 or 3 (A/B timing)
int g_uni_wfreg = 1;    // 4-sweep Jacobi legs: w f carried in registers (1) or parked in the f ring (0)
int g_uni_bulk = 0;     // 1: the 4-sweep legs fill their row ring with cp.async.bulk + mbarrier instead of per-lane cp.async

template <int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int GS, bool WFREG, int MINCTAS, bool BULK = false>
static cudaError_t launch_uni_m(const LevelDev &L, double shift, double omega, const double *v_in, const double *f,
                                double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s, int *slots_out) {
  constexpr int NSTAGE = NU + (RESTRICT ? 1 : 0);
  constexpr int USEFUL = 32 * kC - 2 * uni_halo(NU);
  auto kern = uni5_leg_kernel<NU, PROLONG, RESTRICT, ZEROV, GS, WFREG, MINCTAS, BULK>;
  const size_t smem = uni_smem_bytes(PROLONG, ZEROV, NSTAGE, WFREG, BULK);
  static int occ = 0;  // per instantiation
  if (!occ) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsU * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
  }
  const int strips = (L.ncols + USEFUL - 1) / USEFUL;
  const int gx = (strips + kWarpsU - 1) / kWarpsU;
  const int rpc = leg_rows_per_chunk(L.nrows, gx, occ * num_sms(), NSTAGE, 1 << 20);
  dim3 grid(gx, (L.nrows + rpc - 1) / rpc);
  if (slots_out) {  // geometry query only
    *slots_out = grid.x * grid.y * kWarpsU;
    return cudaSuccess;
  }
  kern<<<grid, kWarpsU * 32, smem, s>>>(L, uni_coef(L.uni_c, L.uni_d, shift, omega), v_in, f, v_out, e_coarse, r_coarse, rpc);
  count_launch();
  return cudaGetLastError();
}

template <int NU, bool PROLONG, bool RESTRICT, bool ZEROV, int GS>
static cudaError_t launch_uni_t(const LevelDev &L, double shift, double omega, const double *v_in, const double *f,
                                double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s, int *slots_out) {
#define UNI_GO(WF, MC) \
  return launch_uni_m<NU, PROLONG, RESTRICT, ZEROV, GS, WF, MC>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s, slots_out)
#define UNI_GO_BULK(WF) \
  return launch_uni_m<NU, PROLONG, RESTRICT, ZEROV, GS, WF, 2, true>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s, slots_out)
  if constexpr (GS == 0 && NU == 4) {
    if (g_uni_bulk) UNI_GO_BULK(true);
  } else if constexpr (GS == 1 && NU == 8) {
    if (g_uni_bulk) UNI_GO_BULK(false);
  }
#undef UNI_GO_BULK
  if constexpr (GS == 0 && NU == 4) {
    // the hot Jacobi legs: both w f variants and both register budgets are built (mgcmt_set_option: uni_wfreg, uni_minctas)
    const int m = g_uni_minctas ? g_uni_minctas : 2;
    if (g_uni_wfreg) { if (m == 3) UNI_GO(true, 3); UNI_GO(true, 2); }
    if (m == 3) UNI_GO(false, 3);
    UNI_GO(false, 2);
  } else if constexpr (GS == 0) {
    UNI_GO(true, 3);
  } else {
    UNI_GO(false, 2);
  }
#undef UNI_GO
}

template <int NU, int GS>
static cudaError_t uni_dispatch_mode(const LevelDev &L, int mode, double shift, double omega, const double *v_in,
                                     const double *f, double *v_out, const double *e_coarse, double *r_coarse,
                                     cudaStream_t s, int *slots_out) {
  switch (mode) {
    case FUSED_SMOOTH:
      if (NU == 0) return cudaErrorInvalidValue;
      return launch_uni_t<NU, false, false, false, GS>(L, shift, omega, v_in, f, v_out, nullptr, nullptr, s, slots_out);
    case FUSED_DOWN:
      return launch_uni_t<NU, false, true, false, GS>(L, shift, omega, v_in, f, v_out, nullptr, r_coarse, s, slots_out);
    case FUSED_DOWN_ZERO:
      return launch_uni_t<NU, false, true, true, GS>(L, shift, omega, v_in, f, v_out, nullptr, r_coarse, s, slots_out);
    case FUSED_UP:
      return launch_uni_t<NU, true, false, false, GS>(L, shift, omega, v_in, f, v_out, e_coarse, nullptr, s, slots_out);
    case FUSED_UP_RQ:  // the last up-leg pass of the finest level: nu2 = 4 Jacobi sweeps or 4 red-black sweeps
      if constexpr ((GS == 0 && NU == 4) || (GS == 1 && NU == 8))
        return launch_uni_t<NU, true, true, false, GS>(L, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s, slots_out);
      else
        return cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}

// host-only view of the coefficient choice (tests/test_host_logic.py checks the consistency it promises without a GPU)
void uni5_coefficients(double c, double d, double shift, double omega, double *out7) {
  const UniCoef k = uni_coef(c, d, shift, omega);
  out7[0] = k.a_smooth; out7[1] = k.a_res; out7[2] = k.dlo; out7[3] = k.nbeta; out7[4] = k.wf; out7[5] = k.invw; out7[6] = k.drem;
}

int g_fused_uni = 1;  // 0: constant-coefficient levels use the general kernels of fused.cu too (A/B, parity)

bool uni5_available(const LevelDev &L) { return g_fused_uni && L.uni == 1 && L.five && L.nrows >= 2 && (L.ncols & 1) == 0; }

// nu = Jacobi sweeps 0..4 (gs = 0) or red-black sweeps 1..4 (gs = 1).  slots_out != nullptr: no launch, only the
// number of per-warp partial-sum slots a FUSED_UP_RQ launch would write.
cudaError_t launch_uni5_leg(const LevelDev &L, int gs, int mode, int nu, double shift, double omega, const double *v_in,
                            const double *f, double *v_out, const double *e_coarse, double *r_coarse, cudaStream_t s,
                            int *slots_out) {
  if (!uni5_available(L)) return cudaErrorInvalidValue;
  if (nu == 0) gs = 0;  // transfer only: no colour order involved
#define UNI_CASE(NUV, GSV) \
  case NUV: return uni_dispatch_mode<(GSV ? 2 * NUV : NUV), GSV>(L, mode, shift, omega, v_in, f, v_out, e_coarse, r_coarse, s, slots_out);
  if (gs) {
    switch (nu) { UNI_CASE(1, 1) UNI_CASE(2, 1) UNI_CASE(3, 1) UNI_CASE(4, 1) }
  } else {
    switch (nu) { UNI_CASE(0, 0) UNI_CASE(1, 0) UNI_CASE(2, 0) UNI_CASE(3, 0) UNI_CASE(4, 0) }
  }
#undef UNI_CASE
  return cudaErrorInvalidValue;
}

int uni5_rq_slots(const LevelDev &L, int gs) {
  if (!uni5_available(L)) return 0;
  int slots = 0;
  if (launch_uni5_leg(L, gs, FUSED_UP_RQ, 4, 0.0, 1.0, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &slots) != cudaSuccess)
    return 0;
  return slots;
}

}  // namespace mgcmt
