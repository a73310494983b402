// api.cu -- the C ABI declared in include/mgcmt_b200.h: grid hierarchy + V-cycle orchestration.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mgcmt_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace mgcmt;

namespace mgcmt {
long long g_launch_count = 0;
}

namespace {

thread_local std::string g_err;

// optional CUDA-event timing of the dominant kernel (finest-level smoother sweeps), for bench.py
struct Profile {
  bool on = false;
  std::vector<cudaEvent_t> pool;  // pairs
  std::vector<int> kind;          // per event: MGCMT_PROF_* of the interval it opens / closes
  size_t used = 0;
} g_prof;

// runtime switches (mgcmt_set_option): fused legs on/off, smallest level width that uses them
int g_opt_fused = 1;
int g_opt_fused_min_cols = 64;
int g_opt_tile_max_cols = 256;   // levels this narrow (or narrower) use the shared-memory tile legs
int g_opt_tail_max_cols = 32;    // levels this narrow are collapsed into the single-CTA tail kernel
int g_opt_tile_gs_max_cols = 0;   // Gauss-Seidel legs of levels this narrow run on shared-memory tiles (one launch per leg); off: measured slower (shared-memory bound: 18 loads per update at 1/4 lane efficiency), see DESIGN.md
int g_opt_coarse_banded = 2;     // coarsest inverse through the banded LU: 0 never, 1 whenever it fits, 2 = when n > 256

void prof_mark(cudaStream_t s, int kind = MGCMT_PROF_SWEEP) {
  if (!g_prof.on) return;
  if (g_prof.used == g_prof.pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof.pool.push_back(e);
    g_prof.kind.push_back(0);
  }
  g_prof.kind[g_prof.used] = kind;
  cudaEventRecord(g_prof.pool[g_prof.used++], s);
}

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
}  // namespace
namespace mgcmt {
int set_error(int code, const std::string &msg) { return fail(code, msg); }
}
namespace {
#define CU(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return fail(MGCMT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));               \
  } while (0)

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
// grid vectors are read as double2: they must be 16-byte aligned
bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
#define NEED_ALIGNED(...)                                                                       \
  do {                                                                                          \
    const void *ps__[] = {__VA_ARGS__};                                                         \
    for (const void *p__ : ps__)                                                                \
      if (!p__ || !al16(p__)) return fail(MGCMT_ERR_ARG, "grid vectors must be non-null and 16-byte aligned"); \
  } while (0)

// scratch for reductions that do not belong to a hierarchy (one per device, single-stream use)
struct Scratch {
  double *partials = nullptr;  // 32 * kReduceBlocks
  double *scal = nullptr;      // 64 scalars
  int *status = nullptr;
};
Scratch g_scratch[64];

int get_scratch(Scratch **out) {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(MGCMT_ERR_STATE, "device index out of range");
  Scratch &s = g_scratch[dev];
  if (!s.partials) {
    CU(cudaMalloc(&s.partials, sizeof(double) * 32 * kReduceBlocks));
    CU(cudaMalloc(&s.scal, sizeof(double) * 64));
    CU(cudaMalloc(&s.status, sizeof(int)));
    CU(cudaMemset(s.status, 0, sizeof(int)));
  }
  *out = &s;
  return MGCMT_OK;
}

struct Level {
  LevelDev dev;
  double *coef = nullptr;  // 6*nrows_glob + 6*ncols doubles
  double *v = nullptr, *f = nullptr, *tmp = nullptr;
  double *zrow = nullptr;  // 3*nrows_glob zeros: the "Ka = 0" factor used by the mass-matrix apply
  size_t n = 0;
};

struct InvEntry {
  double shift;
  double *inv;
  uint64_t stamp;
};

}  // namespace

struct mgcmt_hier {
  int nlev = 0;
  bool coarsen_rows = true;
  bool slab = false;   // row-slab piece of a decomposed grid: only single-level operators are valid
  int halo = 0;        // halo rows above and below the owned rows of every slab level
  int first_work = 0;  // levels below this one have no work vectors (replicated coarse part of a slab solver)
  double *small_partials = nullptr;  // reduction scratch of mgcmt_rayleigh on levels too small to hold their own partials
  double *rq_partials = nullptr;  // per-warp Rayleigh partial sums of the finest up leg (mgcmt_vcycle_rq)
  int rq_slots = 0;
  double *rq_out = nullptr;       // set for the duration of a mgcmt_vcycle_rq call
  bool rq_done = false;
  std::vector<Level> lev;
  std::vector<InvEntry> invs;
  uint64_t clock = 0;
  int *status = nullptr;  // device flag for the Gauss-Jordan
  // mgcmt_vcycle_host_block: three rotating (f, v) device slots, a copy-in and a copy-out stream, events per slot
  struct HostPipe {
    double *f[3] = {nullptr, nullptr, nullptr}, *v[3] = {nullptr, nullptr, nullptr};
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t uploaded[3] = {nullptr, nullptr, nullptr}, cycled[3] = {nullptr, nullptr, nullptr},
                downloaded[3] = {nullptr, nullptr, nullptr};
    bool ready = false;
  } pipe;
};

namespace {

void set_coef_ptrs(Level &L) {
  const int nr = L.dev.nrows_glob, nc = L.dev.ncols;
  double *p = L.coef;
  L.dev.ka_lo = p; p += nr;
  L.dev.ka_di = p; p += nr;
  L.dev.ka_up = p; p += nr;
  L.dev.ma_lo = p; p += nr;
  L.dev.ma_di = p; p += nr;
  L.dev.ma_up = p; p += nr;
  L.dev.kb_lo = p; p += nc;
  L.dev.kb_di = p; p += nc;
  L.dev.kb_up = p; p += nc;
  L.dev.mb_lo = p; p += nc;
  L.dev.mb_di = p; p += nc;
  L.dev.mb_up = p; p += nc;
}

int check_level(const mgcmt_hier *h, int level) {
  if (!h) return fail(MGCMT_ERR_ARG, "null hierarchy");
  if (level < 0 || level >= h->nlev) return fail(MGCMT_ERR_ARG, "level out of range");
  return MGCMT_OK;
}

int smooth_impl(mgcmt_hier *h, int level, int smoother, double shift, double omega, int nu, double *v,
                const double *f, double *tmp, cudaStream_t s) {
  Level &L = h->lev[level];
  if (nu <= 0) return MGCMT_OK;
  if (!tmp) tmp = L.tmp;
  if (smoother == MGCMT_SMOOTH_WJACOBI) {
    double *a = v, *b = tmp;
    for (int it = 0; it < nu; ++it) {
      if (level == 0) prof_mark(s);
      CU(launch_jacobi_sweep(L.dev, shift, omega, a, f, b, nullptr, nullptr, s));
      if (level == 0) prof_mark(s);
      double *t = a; a = b; b = t;
    }
    if (a != v) CU(cudaMemcpyAsync(v, a, sizeof(double) * L.n, cudaMemcpyDeviceToDevice, s));
    return MGCMT_OK;
  }
  if (smoother == MGCMT_SMOOTH_RBGS) {
    if (level == 0) prof_mark(s);
    CU(launch_rbgs(L.dev, shift, omega, nu, v, f, s));
    if (level == 0) prof_mark(s);
    return MGCMT_OK;
  }
  if (smoother == MGCMT_SMOOTH_GSLEX) {
    CU(launch_gs_lex(L.dev, shift, omega, nu, v, f, tmp, s));
    return MGCMT_OK;
  }
  return fail(MGCMT_ERR_ARG, "unknown smoother");
}

int get_inverse(mgcmt_hier *h, double shift, cudaStream_t s, double **out) {
  for (auto &e : h->invs) {
    if (memcmp(&e.shift, &shift, sizeof(double)) == 0) {
      e.stamp = ++h->clock;
      *out = e.inv;
      return MGCMT_OK;
    }
  }
  Level &L = h->lev[h->nlev - 1];
  const int n = (int)L.n;
  double *inv = nullptr;
  const size_t kMaxCached = 16;
  if (h->invs.size() >= kMaxCached) {  // evict the least recently used
    size_t victim = 0;
    for (size_t i = 1; i < h->invs.size(); ++i)
      if (h->invs[i].stamp < h->invs[victim].stamp) victim = i;
    CU(cudaStreamSynchronize(s));
    inv = h->invs[victim].inv;
    h->invs.erase(h->invs.begin() + victim);
  } else {
    CU(cudaMalloc(&inv, sizeof(double) * (size_t)n * n));
  }
  double *aug = nullptr, *mult = nullptr;
  CU(cudaMemsetAsync(h->status, 0, sizeof(int), s));
  // banded LU for coarsest levels beyond a few hundred unknowns (g_opt_coarse_banded: 0 never, 1 whenever it fits,
  // 2 = auto: n > 256), dense Gauss-Jordan otherwise; both with partial pivoting, both leave the dense inverse
  const bool banded = band_inverse_fits(L.dev) && (g_opt_coarse_banded == 1 || (g_opt_coarse_banded == 2 && n > 256));
  if (banded) {
    const int kl = (L.dev.nrows > 1) ? (L.dev.ncols + 1 < n - 1 ? L.dev.ncols + 1 : n - 1) : 1;
    CU(cudaMalloc(&aug, sizeof(double) * band_workspace_doubles(n, kl, kl)));
    CU(launch_band_inverse2d(L.dev, shift, inv, h->status, aug, s));
  } else {
    CU(cudaMalloc(&aug, sizeof(double) * (size_t)n * 2 * n));
    CU(cudaMalloc(&mult, sizeof(double) * (n + 4)));
    CU(launch_build_dense(L.dev, shift, aug, s));
    CU(launch_gauss_jordan(n, aug, h->status, mult, s));
    CU(launch_extract_inverse(n, aug, inv, s));
  }
  int st = 0;
  CU(cudaMemcpyAsync(&st, h->status, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaFree(aug));
  CU(cudaFree(mult));
  if (st != 0) {
    cudaFree(inv);
    return fail(MGCMT_ERR_NUMERIC, "coarsest operator is singular for this shift");
  }
  h->invs.push_back({shift, inv, ++h->clock});
  *out = inv;
  return MGCMT_OK;
}

bool use_tile(const mgcmt_hier *h, int l) {
  const Level &L = h->lev[l];
  return L.dev.ncols <= g_opt_tile_max_cols && L.dev.nrows >= 16 && L.dev.ncols >= 16;
}
cudaError_t launch_leg(const mgcmt_hier *h, int l, int mode, int nu, double shift, double omega, const double *vin,
                       const double *f, double *vout, const double *e, double *rc, cudaStream_t s) {
  const Level &L = h->lev[l];
  if (use_tile(h, l)) return launch_tile_leg(L.dev, mode, nu, shift, omega, vin, f, vout, e, rc, s);
  return launch_fused_leg(L.dev, mode, nu, shift, omega, vin, f, vout, e, rc, s);
}

// can levels l .. coarsest run inside the single-CTA tail kernel?
bool use_tail(const mgcmt_hier *h, int l, int smoother, int nu1, int nu2, bool v_zero) {
  if (!g_opt_fused || smoother != MGCMT_SMOOTH_WJACOBI || !v_zero || g_opt_tail_max_cols <= 0) return false;
  if (nu1 != 4 || nu2 != 4) return false;
  const int nl = h->nlev - l;
  if (nl < 2 || nl > kTailMaxLevels) return false;
  const Level &L = h->lev[l];
  if (h->slab || L.dev.row0 != 0) return false;
  if (h->coarsen_rows) {
    if (L.dev.ncols > g_opt_tail_max_cols || L.dev.nrows != L.dev.ncols) return false;
  } else {
    if (L.dev.nrows != 1 || L.dev.ncols > 1024) return false;  // 1-D: the whole cycle of a <= 1024-point grid fits one CTA
  }
  if (h->lev[h->nlev - 1].n > 256) return false;  // the dense inverse is read by one CTA
  LevelDev devs[kTailMaxLevels];
  for (int k = l; k < h->nlev; ++k) devs[k - l] = h->lev[k].dev;
  return tail_smem_bytes(devs, nl) <= kTailMaxSmem;
}

bool use_fused(const mgcmt_hier *h, int l, int smoother) {
  const Level &L = h->lev[l];
  return g_opt_fused && smoother == MGCMT_SMOOTH_WJACOBI && h->coarsen_rows && L.dev.nrows >= 16 &&
         L.dev.ncols >= 16 && (L.dev.ncols >= g_opt_fused_min_cols || use_tile(h, l)) && L.dev.row0 == 0;
}

bool use_tile_gs(const mgcmt_hier *h, int l);
bool use_fused_gs(const mgcmt_hier *h, int l) {
  const Level &L = h->lev[l];
  return g_opt_fused && h->coarsen_rows && L.dev.nrows >= 16 && (L.dev.ncols >= 64 || use_tile_gs(h, l)) && L.dev.row0 == 0;
}

bool use_tile_gs(const mgcmt_hier *h, int l) {
  const Level &L = h->lev[l];
  return L.dev.ncols <= g_opt_tile_gs_max_cols && L.dev.nrows >= 16 && L.dev.ncols >= 16 && L.dev.row0 == 0 && !h->slab;
}
// colour sweeps one fused Gauss-Seidel pass takes on level l
int gs_maxpass(const mgcmt_hier *h, int l) { return use_tile_gs(h, l) ? 4 : (h->lev[l].dev.five ? 4 : 2); }

// one fused pass of `nu` sweeps (Jacobi: streaming or tile legs; gs: colour-stage streaming legs, tile legs on small levels)
cudaError_t launch_pass(const mgcmt_hier *h, int l, bool gs, int mode, int nu, double shift, double omega,
                        const double *vin, const double *f, double *vout, const double *e, double *rc, cudaStream_t s) {
  if (gs && use_tile_gs(h, l)) return launch_tile_gs_leg(h->lev[l].dev, mode, nu, shift, omega, vin, f, vout, e, rc, s);
  if (gs) return launch_fused_gs_leg(h->lev[l].dev, mode, nu, shift, omega, vin, f, vout, e, rc, s);
  return launch_leg(h, l, mode, nu, shift, omega, vin, f, vout, e, rc, s);
}

// down leg with the fused kernels: nu1 sweeps (in passes of <= maxpass) + residual + restriction into C.f.
// Returns in *cur the buffer that holds the smoothed iterate (v or L.tmp).
int fused_down(mgcmt_hier *h, int l, bool gs, double shift, double omega, int nu1, double *v, const double *f,
               bool v_zero, double **cur, cudaStream_t s) {
  Level &L = h->lev[l];
  Level &C = h->lev[l + 1];
  const int maxpass = gs ? gs_maxpass(h, l) : 4;
  double *a = v, *b = L.tmp;
  int left = nu1;
  while (left > maxpass) {
    if (l == 0) prof_mark(s);
    if (v_zero) CU(cudaMemsetAsync(a, 0, sizeof(double) * L.n, s));  // smooth-only passes read their input
    CU(launch_pass(h, l, gs, FUSED_SMOOTH, maxpass, shift, omega, a, f, b, nullptr, nullptr, s));
    if (l == 0) prof_mark(s);
    double *t = a; a = b; b = t;
    left -= maxpass;
    v_zero = false;
  }
  if (v_zero && left == 0) CU(cudaMemsetAsync(a, 0, sizeof(double) * L.n, s));
  const int dkind = (v_zero && left > 0) ? MGCMT_PROF_DOWN_ZERO : MGCMT_PROF_DOWN;
  if (l == 0) prof_mark(s, dkind);
  CU(launch_pass(h, l, gs, (v_zero && left > 0) ? FUSED_DOWN_ZERO : FUSED_DOWN, left, shift, omega, a, f, b, nullptr,
                 C.f, s));
  if (l == 0) prof_mark(s, dkind);
  if (left > 0) { double *t = a; a = b; b = t; }
  *cur = a;
  return MGCMT_OK;
}

// up leg: cur (+ P e) -> nu2 sweeps; result must end in v
int fused_up(mgcmt_hier *h, int l, bool gs, double shift, double omega, int nu2, double *v, const double *f, double *cur,
             const double *e, cudaStream_t s) {
  Level &L = h->lev[l];
  const int maxpass = gs ? gs_maxpass(h, l) : 4;
  double *a = cur, *b = (cur == v) ? L.tmp : v;
  int left = nu2;
  const int first = left > maxpass ? maxpass : left;
  const int ukind = (l == 0 && h->rq_out) ? MGCMT_PROF_UP_RQ : MGCMT_PROF_UP;
  if (l == 0) prof_mark(s, ukind);
  // finest level, Rayleigh quotient requested and this pass is the last one: the up leg also leaves the partial sums
  const int slots = (l == 0 && h->rq_out && first == 4 && left == 4 && !(gs ? use_tile_gs(h, l) : use_tile(h, l)))
                        ? fused_rq_slots(L.dev, gs ? 1 : 0) : 0;
  if (slots > 0) {
    if (slots > h->rq_slots) {
      cudaFree(h->rq_partials);
      h->rq_partials = nullptr;
      CU(cudaMalloc(&h->rq_partials, sizeof(double) * 2 * slots));
      h->rq_slots = slots;
    }
    if (gs) CU(launch_fused_gs_leg(L.dev, FUSED_UP_RQ, 4, shift, omega, a, f, b, e, h->rq_partials, s));
    else CU(launch_fused_leg(L.dev, FUSED_UP_RQ, 4, shift, omega, a, f, b, e, h->rq_partials, s));
    CU(launch_finish(2, slots, h->rq_partials, h->rq_out, s));
    h->rq_done = true;
  } else {
    CU(launch_pass(h, l, gs, FUSED_UP, first, shift, omega, a, f, b, e, nullptr, s));
  }
  if (l == 0) prof_mark(s, ukind);
  { double *t = a; a = b; b = t; }
  left -= first;
  while (left > 0) {
    const int nu = left > maxpass ? maxpass : left;
    if (l == 0) prof_mark(s);
    CU(launch_pass(h, l, gs, FUSED_SMOOTH, nu, shift, omega, a, f, b, nullptr, nullptr, s));
    if (l == 0) prof_mark(s);
    double *t = a; a = b; b = t;
    left -= nu;
  }
  if (a != v) CU(cudaMemcpyAsync(v, a, sizeof(double) * L.n, cudaMemcpyDeviceToDevice, s));
  return MGCMT_OK;
}

int vcycle_level(mgcmt_hier *h, int l, double shift, int nu1, int nu2, int smoother, double omega, double *v,
                 const double *f, bool v_zero, cudaStream_t s) {
  Level &L = h->lev[l];
  if (l == h->nlev - 1) {
    double *inv = nullptr;
    int rc = get_inverse(h, shift, s, &inv);
    if (rc) return rc;
    CU(launch_gemv((int)L.n, inv, f, v, s));
    return MGCMT_OK;
  }
  Level &C = h->lev[l + 1];
  int rc;
  if (use_tail(h, l, smoother, nu1, nu2, v_zero)) {
    double *inv = nullptr;
    rc = get_inverse(h, shift, s, &inv);
    if (rc) return rc;
    LevelDev devs[kTailMaxLevels];
    for (int k = l; k < h->nlev; ++k) devs[k - l] = h->lev[k].dev;
    CU(launch_tail(devs, h->nlev - l, h->coarsen_rows, inv, shift, omega, f, v, s));
    return MGCMT_OK;
  }
  const bool gs = (smoother == MGCMT_SMOOTH_RBGS) && use_fused_gs(h, l);
  if (use_fused(h, l, smoother) || gs) {
    double *cur = nullptr;
    rc = fused_down(h, l, gs, shift, omega, nu1, v, f, v_zero, &cur, s);
    if (rc) return rc;
    // coarse levels always run 4/4 (MGCMTSolver.py:320 does not forward nu1/nu2); their start is zero
    rc = vcycle_level(h, l + 1, shift, 4, 4, smoother, omega, C.v, C.f, true, s);
    if (rc) return rc;
    return fused_up(h, l, gs, shift, omega, nu2, v, f, cur, C.v, s);
  }
  if (v_zero) CU(cudaMemsetAsync(v, 0, sizeof(double) * L.n, s));
  rc = smooth_impl(h, l, smoother, shift, omega, nu1, v, f, nullptr, s);
  if (rc) return rc;
  CU(launch_residual_restrict(L.dev, h->coarsen_rows, shift, v, f, C.f, s));
  rc = vcycle_level(h, l + 1, shift, 4, 4, smoother, omega, C.v, C.f, true, s);
  if (rc) return rc;
  CU(launch_prolong(L.dev, h->coarsen_rows, true, C.v, v, s));
  return smooth_impl(h, l, smoother, shift, omega, nu2, v, f, nullptr, s);
}

}  // namespace

extern "C" {

int mgcmt_abi_version(void) { return MGCMT_ABI_VERSION; }

long long mgcmt_launch_count(void) { return mgcmt::g_launch_count; }

int mgcmt_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.used = 0;
  return MGCMT_OK;
}

int mgcmt_profile_read_kinds(double *ms_by_kind, long long *intervals_by_kind) {
  CU(cudaDeviceSynchronize());
  double tot[MGCMT_PROF_KINDS] = {0};
  long long cnt[MGCMT_PROF_KINDS] = {0};
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, g_prof.pool[i], g_prof.pool[i + 1]));
    const int k = g_prof.kind[i];
    if (k >= 0 && k < MGCMT_PROF_KINDS) { tot[k] += ms; ++cnt[k]; }
  }
  g_prof.used = 0;
  for (int k = 0; k < MGCMT_PROF_KINDS; ++k) {
    if (ms_by_kind) ms_by_kind[k] = tot[k];
    if (intervals_by_kind) intervals_by_kind[k] = cnt[k];
  }
  return MGCMT_OK;
}

int mgcmt_profile_read(double *ms_total, long long *intervals) {
  double tot[MGCMT_PROF_KINDS];
  long long cnt[MGCMT_PROF_KINDS];
  int rc = mgcmt_profile_read_kinds(tot, cnt);
  if (rc) return rc;
  double t = 0.0;
  long long c = 0;
  for (int k = 0; k < MGCMT_PROF_KINDS; ++k) { t += tot[k]; c += cnt[k]; }
  if (ms_total) *ms_total = t;
  if (intervals) *intervals = c;
  return MGCMT_OK;
}
const char *mgcmt_last_error(void) { return g_err.c_str(); }

// Shared builder.  Plain hierarchy: row_begin = 0, nrows_own = nrows_glob, halo = 0, nlev from lowest_level.
// Slab piece: levels 0..nlev-1 hold rows [row_begin >> l, (row_begin + nrows_own) >> l) plus `halo` rows on both
// sides; coefficient arrays are always the full global ones (O(N)).
static int build_hier(mgcmt_hier_t **out, int nrows_glob, int ncols, int coarsen_rows, const double *h_row_lo,
                      const double *h_row_di, const double *h_row_up, const double *h_col_lo, const double *h_col_di,
                      const double *h_col_up, int nlev, bool slab, int row_begin, int nrows_own, int halo,
                      int coarse_full_after, int first_work, cudaStream_t s) {
  mgcmt_hier *h = new mgcmt_hier();
  h->nlev = nlev;
  h->coarsen_rows = coarsen_rows != 0;
  h->slab = slab;
  h->halo = halo;
  h->first_work = first_work;
  h->lev.resize(nlev);
  auto bail = [&](int code, const std::string &msg) {
    mgcmt_hier_destroy(h);
    return fail(code, msg);
  };
#define CUB(expr)                                                                          \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) return bail(MGCMT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

  CUB(cudaMalloc(&h->status, sizeof(int)));
  int nr = nrows_glob, nc = ncols;
  for (int l = 0; l < nlev; ++l) {
    Level &L = h->lev[l];
    const int own = slab ? (nrows_own >> l) : nr;
    const int begin = slab ? (row_begin >> l) : 0;
    L.dev.nrows = slab ? own + 2 * halo : nr;
    L.dev.ncols = nc;
    L.dev.row0 = slab ? begin - halo : 0;
    L.dev.nrows_glob = nr;
    L.dev.five = (l == 0 || !coarsen_rows) ? 1 : 0;
    L.dev.crow_shift = 0;
    L.dev.nrows_coarse = 0;
    L.dev.rq_lo = slab ? halo : 0;
    L.dev.rq_hi = slab ? halo + own : L.dev.nrows;
    L.dev.uni = 0;
    L.dev.uni_c = L.dev.uni_d = 0.0;
    if (l == 0 && coarsen_rows && nr >= 2 && nc >= 2) {
      // constant 5-point stencil (the infinite well, 2DPot.py:25-26)?  Then the legs of this level run in fused_uni.cu
      const double c = h_col_lo[1];
      bool uni = true;
      for (int i = 0; i < nr && uni; ++i)
        uni = h_row_di[i] == h_row_di[0] && (i == 0 || h_row_lo[i] == c) && (i + 1 == nr || h_row_up[i] == c);
      for (int j = 0; j < nc && uni; ++j)
        uni = h_col_di[j] == h_col_di[0] && (j == 0 || h_col_lo[j] == c) && (j + 1 == nc || h_col_up[j] == c);
      if (uni && c != 0.0) {
        L.dev.uni = 1;
        L.dev.uni_c = c;
        L.dev.uni_d = h_row_di[0] + h_col_di[0];
      }
    }
    if (slab) {
      if (l + 1 < nlev) {  // coarse level is a slab piece too
        L.dev.crow_shift = halo / 2;
        L.dev.nrows_coarse = (own >> 1) + 2 * halo;
      } else if (coarse_full_after) {  // coarse level is the full (replicated) grid
        L.dev.crow_shift = (begin - halo) / 2;  // exact: begin and halo are even
        L.dev.nrows_coarse = nr >> 1;
      }
    }
    L.n = (size_t)L.dev.nrows * nc;
    CUB(cudaMalloc(&L.coef, sizeof(double) * (6 * (size_t)nr + 6 * (size_t)nc)));
    set_coef_ptrs(L);
    CUB(cudaMalloc(&L.zrow, sizeof(double) * 3 * (size_t)nr));
    CUB(cudaMemsetAsync(L.zrow, 0, sizeof(double) * 3 * (size_t)nr, s));
    if (l >= first_work) {
      CUB(cudaMalloc(&L.tmp, sizeof(double) * L.n));
      CUB(cudaMemsetAsync(L.tmp, 0, sizeof(double) * L.n, s));
      if (l > 0 && !slab) {  // slab pieces: the caller owns the level vectors (multigridcmt_b200/slab.py)
        CUB(cudaMalloc(&L.v, sizeof(double) * L.n));
        CUB(cudaMalloc(&L.f, sizeof(double) * L.n));
        CUB(cudaMemsetAsync(L.v, 0, sizeof(double) * L.n, s));
        CUB(cudaMemsetAsync(L.f, 0, sizeof(double) * L.n, s));
      }
    }
    if (l == 0) {
      std::vector<double> host(6 * (size_t)nr + 6 * (size_t)nc, 0.0);
      double *p = host.data();
      for (int i = 0; i < nr; ++i) {
        p[i] = (i > 0) ? h_row_lo[i] : 0.0;
        p[nr + i] = h_row_di[i];
        p[2 * nr + i] = (i + 1 < nr) ? h_row_up[i] : 0.0;
        p[4 * nr + i] = 1.0;  // Ma = I
      }
      p += 6 * nr;
      for (int j = 0; j < nc; ++j) {
        p[j] = (j > 0) ? h_col_lo[j] : 0.0;
        p[nc + j] = h_col_di[j];
        p[2 * nc + j] = (j + 1 < nc) ? h_col_up[j] : 0.0;
        p[4 * nc + j] = 1.0;  // Mb = I
      }
      CUB(cudaMemcpyAsync(L.coef, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, s));
      CUB(cudaStreamSynchronize(s));  // host vector goes out of scope
    } else {
      Level &F = h->lev[l - 1];
      const int nrf = F.dev.nrows_glob, ncf = F.dev.ncols;
      if (coarsen_rows) {
        CUB(launch_galerkin_tridiag(nrf, F.dev.ka_lo, F.dev.ka_di, F.dev.ka_up, (double *)L.dev.ka_lo,
                                    (double *)L.dev.ka_di, (double *)L.dev.ka_up, s));
        CUB(launch_galerkin_tridiag(nrf, F.dev.ma_lo, F.dev.ma_di, F.dev.ma_up, (double *)L.dev.ma_lo,
                                    (double *)L.dev.ma_di, (double *)L.dev.ma_up, s));
      } else {
        CUB(cudaMemcpyAsync((void *)L.dev.ka_lo, F.dev.ka_lo, sizeof(double) * 6 * nrf, cudaMemcpyDeviceToDevice, s));
      }
      CUB(launch_galerkin_tridiag(ncf, F.dev.kb_lo, F.dev.kb_di, F.dev.kb_up, (double *)L.dev.kb_lo,
                                  (double *)L.dev.kb_di, (double *)L.dev.kb_up, s));
      CUB(launch_galerkin_tridiag(ncf, F.dev.mb_lo, F.dev.mb_di, F.dev.mb_up, (double *)L.dev.mb_lo,
                                  (double *)L.dev.mb_di, (double *)L.dev.mb_up, s));
    }
    if (l + 1 < nlev) {
      nc >>= 1;
      if (coarsen_rows) nr >>= 1;
    }
  }
  CUB(cudaStreamSynchronize(s));
  // Galerkin levels of a constant-coefficient operator: all four tridiagonal factors constant but for their last
  // diagonal entry (the truncated last row of R)?  Then their legs run in fused_uni9.cu.
  for (int l = 1; l < nlev && coarsen_rows; ++l) {
    Level &L = h->lev[l];
    const int nrl = L.dev.nrows_glob, ncl = L.dev.ncols;
    if (nrl < 8 || ncl < 8) continue;
    std::vector<double> hc(6 * (size_t)nrl + 6 * (size_t)ncl);
    CUB(cudaMemcpy(hc.data(), L.coef, sizeof(double) * hc.size(), cudaMemcpyDeviceToHost));
    auto tri_uniform = [](const double *lo, const double *di, const double *up, int n, double *off, double *d0, double *dlast) {
      const double o = up[0];
      for (int i = 0; i < n; ++i) {
        if (i > 0 && lo[i] != o) return false;
        if (i + 1 < n && up[i] != o) return false;
        if (i + 1 < n && di[i] != di[0]) return false;
      }
      *off = o; *d0 = di[0]; *dlast = di[n - 1];
      return true;
    };
    const double *r = hc.data(), *c = hc.data() + 6 * (size_t)nrl;
    double *u = L.dev.u9;
    const bool ok = tri_uniform(r, r + nrl, r + 2 * nrl, nrl, &u[0], &u[1], &u[4]) &&
                    tri_uniform(r + 3 * nrl, r + 4 * nrl, r + 5 * nrl, nrl, &u[2], &u[3], &u[5]) &&
                    tri_uniform(c, c + ncl, c + 2 * ncl, ncl, &u[6], &u[7], &u[10]) &&
                    tri_uniform(c + 3 * ncl, c + 4 * ncl, c + 5 * ncl, ncl, &u[8], &u[9], &u[11]);
    if (ok) L.dev.uni = 2;
  }
#undef CUB
  *out = h;
  return MGCMT_OK;
}

static int check_create_args(mgcmt_hier_t **out, int nrows, int ncols, int coarsen_rows, const double *a,
                             const double *b, const double *c, const double *d, const double *e, const double *f) {
  if (!out) return fail(MGCMT_ERR_ARG, "out is null");
  *out = nullptr;
  if (!a || !b || !c || !d || !e || !f) return fail(MGCMT_ERR_ARG, "null coefficient array");
  if (!is_pow2(ncols) || ncols < 2) return fail(MGCMT_ERR_ARG, "ncols must be a power of two >= 2");
  if (!is_pow2(nrows)) return fail(MGCMT_ERR_ARG, "nrows must be a power of two (1 for 1-D)");
  if (!coarsen_rows && nrows != 1) return fail(MGCMT_ERR_ARG, "coarsen_rows = 0 needs nrows == 1");
  return MGCMT_OK;
}

int mgcmt_hier_create2(mgcmt_hier_t **out, int nrows, int ncols, int coarsen_rows, const double *h_row_lo,
                       const double *h_row_di, const double *h_row_up, const double *h_col_lo,
                       const double *h_col_di, const double *h_col_up, int lowest_level, int first_work_level,
                       void *stream) {
  int rc = check_create_args(out, nrows, ncols, coarsen_rows, h_row_lo, h_row_di, h_row_up, h_col_lo, h_col_di, h_col_up);
  if (rc) return rc;
  if (!is_pow2(lowest_level) || lowest_level < 2 || lowest_level > ncols)
    return fail(MGCMT_ERR_ARG, "lowest_level must be a power of two in [2, ncols]");
  int nlev = 1;
  for (int c = ncols; c > lowest_level; c >>= 1) ++nlev;
  if (coarsen_rows && (nrows >> (nlev - 1)) < 1)
    return fail(MGCMT_ERR_ARG, "nrows too small for the requested number of levels");
  if (first_work_level < 0 || first_work_level >= nlev) return fail(MGCMT_ERR_ARG, "bad first_work_level");
  {
    const long long nc_rows = coarsen_rows ? (nrows >> (nlev - 1)) : nrows;
    const long long ncoarse = nc_rows * lowest_level;
    if (ncoarse > 4096) return fail(MGCMT_ERR_ARG, "coarsest level larger than 4096 unknowns is not supported");
  }
  return build_hier(out, nrows, ncols, coarsen_rows, h_row_lo, h_row_di, h_row_up, h_col_lo, h_col_di, h_col_up, nlev,
                    false, 0, nrows, 0, 0, first_work_level, (cudaStream_t)stream);
}

int mgcmt_hier_create(mgcmt_hier_t **out, int nrows, int ncols, int coarsen_rows, const double *h_row_lo,
                      const double *h_row_di, const double *h_row_up, const double *h_col_lo,
                      const double *h_col_di, const double *h_col_up, int lowest_level, void *stream) {
  return mgcmt_hier_create2(out, nrows, ncols, coarsen_rows, h_row_lo, h_row_di, h_row_up, h_col_lo, h_col_di,
                            h_col_up, lowest_level, 0, stream);
}

int mgcmt_hier_create_slab(mgcmt_hier_t **out, int nrows_glob, int ncols, int row_begin, int nrows_own, int nlevels,
                           int halo, const double *h_row_lo, const double *h_row_di, const double *h_row_up,
                           const double *h_col_lo, const double *h_col_di, const double *h_col_up, void *stream) {
  int rc = check_create_args(out, nrows_glob, ncols, 1, h_row_lo, h_row_di, h_row_up, h_col_lo, h_col_di, h_col_up);
  if (rc) return rc;
  if (nlevels < 1 || nlevels > 16) return fail(MGCMT_ERR_ARG, "bad number of slab levels");
  if (halo < 6 || (halo & 1)) return fail(MGCMT_ERR_ARG, "halo must be even and >= 6 (NU + 2 rows: 6 for 4 Jacobi sweeps, 10 for 8 Gauss-Seidel colour stages)");
  const int align = 1 << nlevels;  // slab cuts must stay on even rows on every distributed level
  if (row_begin < 0 || nrows_own <= 0 || row_begin + nrows_own > nrows_glob || (row_begin % align) || (nrows_own % align))
    return fail(MGCMT_ERR_ARG, "slab rows must be multiples of 2^nlevels inside the grid");
  if ((nrows_own >> (nlevels - 1)) < halo) return fail(MGCMT_ERR_ARG, "coarsest slab level owns fewer rows than the halo");
  if ((ncols >> (nlevels - 1)) < 64) return fail(MGCMT_ERR_ARG, "slab levels must be at least 64 columns wide");
  return build_hier(out, nrows_glob, ncols, 1, h_row_lo, h_row_di, h_row_up, h_col_lo, h_col_di, h_col_up, nlevels, true,
                    row_begin, nrows_own, halo, 1, 0, (cudaStream_t)stream);
}

int mgcmt_hier_destroy(mgcmt_hier_t *h) {
  if (!h) return MGCMT_OK;
  for (auto &L : h->lev) {
    cudaFree(L.coef);
    cudaFree(L.v);
    cudaFree(L.f);
    cudaFree(L.tmp);
    cudaFree(L.zrow);
  }
  for (auto &e : h->invs) cudaFree(e.inv);
  for (int i = 0; i < 3; ++i) {
    cudaFree(h->pipe.f[i]);
    cudaFree(h->pipe.v[i]);
    if (h->pipe.uploaded[i]) cudaEventDestroy(h->pipe.uploaded[i]);
    if (h->pipe.cycled[i]) cudaEventDestroy(h->pipe.cycled[i]);
    if (h->pipe.downloaded[i]) cudaEventDestroy(h->pipe.downloaded[i]);
  }
  if (h->pipe.in) cudaStreamDestroy(h->pipe.in);
  if (h->pipe.out) cudaStreamDestroy(h->pipe.out);
  cudaFree(h->rq_partials);
  cudaFree(h->small_partials);
  cudaFree(h->status);
  delete h;
  return MGCMT_OK;
}

int mgcmt_hier_num_levels(const mgcmt_hier_t *h) { return h ? h->nlev : 0; }

int mgcmt_hier_level_buffers(mgcmt_hier_t *h, int level, double **d_v, double **d_f, double **d_tmp) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (d_v) *d_v = h->lev[level].v;
  if (d_f) *d_f = h->lev[level].f;
  if (d_tmp) *d_tmp = h->lev[level].tmp;
  return MGCMT_OK;
}

int mgcmt_hier_level_shape(const mgcmt_hier_t *h, int level, int *nrows, int *ncols) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (nrows) *nrows = h->lev[level].dev.nrows;
  if (ncols) *ncols = h->lev[level].dev.ncols;
  return MGCMT_OK;
}

int mgcmt_hier_level_coefs(const mgcmt_hier_t *h, int level, double *h_rowcoef6, double *h_colcoef6) {
  int rc = check_level(h, level);
  if (rc) return rc;
  const Level &L = h->lev[level];
  const size_t nr = L.dev.nrows_glob, nc = L.dev.ncols;
  CU(cudaDeviceSynchronize());
  if (h_rowcoef6) CU(cudaMemcpy(h_rowcoef6, L.coef, sizeof(double) * 6 * nr, cudaMemcpyDeviceToHost));
  if (h_colcoef6) CU(cudaMemcpy(h_colcoef6, L.coef + 6 * nr, sizeof(double) * 6 * nc, cudaMemcpyDeviceToHost));
  return MGCMT_OK;
}

int mgcmt_apply(mgcmt_hier_t *h, int level, double shift, const double *d_x, double *d_y, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  NEED_ALIGNED(d_x, d_y);
  CU(launch_apply(h->lev[level].dev, shift, d_x, d_y, nullptr, nullptr, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_apply_mass(mgcmt_hier_t *h, int level, const double *d_x, double *d_y, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  NEED_ALIGNED(d_x, d_y);
  // M_l = Ma (x) Mb is the operator with Kb := Mb and Ka := 0
  const Level &L = h->lev[level];
  LevelDev m = L.dev;
  m.kb_lo = L.dev.mb_lo; m.kb_di = L.dev.mb_di; m.kb_up = L.dev.mb_up;
  m.ka_lo = L.zrow; m.ka_di = L.zrow + L.dev.nrows_glob; m.ka_up = L.zrow + 2 * (size_t)L.dev.nrows_glob;
  if (L.dev.nrows_glob > 1 && h->coarsen_rows) m.five = 0;  // 2-D: needs the Ma factor (identity on the finest level)
  CU(launch_apply(m, 0.0, d_x, d_y, nullptr, nullptr, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_axpby(long long n, double a, const double *d_x, double b, const double *d_y, double *d_out, void *stream) {
  if (n < 0 || !d_x || !d_y || !d_out) return fail(MGCMT_ERR_ARG, "bad axpby arguments");
  CU(launch_axpby(n, a, d_x, b, d_y, d_out, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_residual(mgcmt_hier_t *h, int level, double shift, const double *d_v, const double *d_f, double *d_r,
                   void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  NEED_ALIGNED(d_v, d_f, d_r);
  CU(launch_residual(h->lev[level].dev, shift, d_v, d_f, d_r, nullptr, nullptr, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_smooth(mgcmt_hier_t *h, int level, int smoother, double shift, double omega, int nu, double *d_v,
                 const double *d_f, double *d_tmp, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  NEED_ALIGNED(d_v, d_f);
  if (d_tmp && !al16(d_tmp)) return fail(MGCMT_ERR_ARG, "d_tmp must be 16-byte aligned");
  if (!d_tmp && !h->lev[level].tmp) return fail(MGCMT_ERR_STATE, "level has no scratch vector in this hierarchy");
  if (h->slab) return fail(MGCMT_ERR_STATE, "slab pieces are driven leg by leg (mgcmt_fused_leg)");
  return smooth_impl(h, level, smoother, shift, omega, nu, d_v, d_f, d_tmp, (cudaStream_t)stream);
}

int mgcmt_restrict(mgcmt_hier_t *h, int level, const double *d_fine, double *d_coarse, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (level + 1 >= h->nlev) return fail(MGCMT_ERR_ARG, "no coarser level");
  NEED_ALIGNED(d_fine, d_coarse);
  CU(launch_restrict(h->lev[level].dev, h->coarsen_rows, d_fine, d_coarse, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_residual_restrict(mgcmt_hier_t *h, int level, double shift, const double *d_v, const double *d_f,
                            double *d_rcoarse, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (level + 1 >= h->nlev) return fail(MGCMT_ERR_ARG, "no coarser level");
  NEED_ALIGNED(d_v, d_f, d_rcoarse);
  CU(launch_residual_restrict(h->lev[level].dev, h->coarsen_rows, shift, d_v, d_f, d_rcoarse,
                              (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_prolong(mgcmt_hier_t *h, int level, const double *d_coarse, double *d_fine, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (level + 1 >= h->nlev) return fail(MGCMT_ERR_ARG, "no coarser level");
  NEED_ALIGNED(d_coarse, d_fine);
  CU(launch_prolong(h->lev[level].dev, h->coarsen_rows, false, d_coarse, d_fine, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_prolong_correct(mgcmt_hier_t *h, int level, const double *d_ecoarse, double *d_v, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (level + 1 >= h->nlev) return fail(MGCMT_ERR_ARG, "no coarser level");
  NEED_ALIGNED(d_ecoarse, d_v);
  CU(launch_prolong(h->lev[level].dev, h->coarsen_rows, true, d_ecoarse, d_v, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_coarse_solve(mgcmt_hier_t *h, double shift, const double *d_f, double *d_v, void *stream) {
  if (!h) return fail(MGCMT_ERR_ARG, "null hierarchy");
  if (!d_f || !d_v || d_f == d_v) return fail(MGCMT_ERR_ARG, "need distinct non-null f and v");
  double *inv = nullptr;
  int rc = get_inverse(h, shift, (cudaStream_t)stream, &inv);
  if (rc) return rc;
  CU(launch_gemv((int)h->lev[h->nlev - 1].n, inv, d_f, d_v, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_debug_uni_coefficients(double c, double d, double shift, double omega, double *h_out7) {
  if (!h_out7) return fail(MGCMT_ERR_ARG, "null output");
  mgcmt::uni5_coefficients(c, d, shift, omega, h_out7);
  return MGCMT_OK;
}

namespace {
// fills the shared memory of every SM with NaN bit patterns and leaves (shared memory is not cleared between kernels)
__global__ void poison_shared_kernel(int granules, double *sink) {
  extern __shared__ double2 poison_smem[];
  const double nan = __longlong_as_double(0x7ff8dead7ff8deadLL);
  for (int i = threadIdx.x; i < granules; i += blockDim.x) poison_smem[i] = make_double2(nan, nan);
  __syncthreads();
  if (sink && poison_smem[(threadIdx.x * 7) % granules].x == 0.0) *sink = 1.0;   // keeps the stores alive; never true
}
}  // namespace

int mgcmt_debug_poison_shared_memory(void *stream) {
  const int bytes = 100 * 1024;   // two such CTAs are resident per SM
  static bool ready = false;
  if (!ready) {
    CU(cudaFuncSetAttribute(poison_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ready = true;
  }
  poison_shared_kernel<<<4 * num_sms(), 128, bytes, (cudaStream_t)stream>>>(bytes / 16, nullptr);
  CU(cudaGetLastError());
  return MGCMT_OK;
}

int mgcmt_debug_leg_rows_per_chunk(int nrows, int gx, int slots, int nstage, int max_rpc) {
  return mgcmt::leg_rows_per_chunk(nrows, gx, slots, nstage, max_rpc);
}

int mgcmt_set_option(const char *name, int value) {
  if (!name) return fail(MGCMT_ERR_ARG, "null option name");
  if (!strcmp(name, "fused")) { g_opt_fused = value; return MGCMT_OK; }
  if (!strcmp(name, "fused_min_cols")) { g_opt_fused_min_cols = value; return MGCMT_OK; }
  if (!strcmp(name, "tile_max_cols")) { g_opt_tile_max_cols = value; return MGCMT_OK; }
  if (!strcmp(name, "tail_max_cols")) { g_opt_tail_max_cols = value; return MGCMT_OK; }
  if (!strcmp(name, "tile_gs_max_cols")) { g_opt_tile_gs_max_cols = value; return MGCMT_OK; }
  if (!strcmp(name, "stage_threads")) { g_stage_threads = value; return MGCMT_OK; }
  if (!strcmp(name, "stage_chunk_kib")) { g_stage_chunk_kib = value; return MGCMT_OK; }
  if (!strcmp(name, "coarse_banded")) {
    if (value < 0 || value > 2) return fail(MGCMT_ERR_ARG, "coarse_banded must be 0, 1 or 2 (auto)");
    g_opt_coarse_banded = value;
    return MGCMT_OK;
  }
  if (!strcmp(name, "fused_uni")) { mgcmt::g_fused_uni = value ? 1 : 0; return MGCMT_OK; }
  if (!strcmp(name, "leg_min_rpc")) {
    if (value < 2 || (value & 1)) return fail(MGCMT_ERR_ARG, "leg_min_rpc must be even and >= 2");
    mgcmt::g_leg_min_rpc = value;
    return MGCMT_OK;
  }
  if (!strcmp(name, "fused_skew_cols")) { mgcmt::g_fused_skew_cols = value; return MGCMT_OK; }
  if (!strcmp(name, "uni9_min_cols")) { mgcmt::g_uni9_min_cols = value; return MGCMT_OK; }
  if (!strcmp(name, "uni9_lag")) {
    if (value != 1 && value != 2) return fail(MGCMT_ERR_ARG, "uni9_lag must be 1 or 2");
    mgcmt::g_uni9_lag = value;
    return MGCMT_OK;
  }
  if (!strcmp(name, "fused_uni9")) {
    if (value < 0 || value > 2) return fail(MGCMT_ERR_ARG, "fused_uni9 must be 0, 1 or 2 (auto)");
    mgcmt::g_fused_uni9 = value;
    return MGCMT_OK;
  }
  if (!strcmp(name, "uni_bulk")) { mgcmt::g_uni_bulk = value ? 1 : 0; return MGCMT_OK; }
  if (!strcmp(name, "uni_wfreg")) { mgcmt::g_uni_wfreg = value ? 1 : 0; return MGCMT_OK; }
  if (!strcmp(name, "uni_minctas")) {
    if (value != 0 && value != 2 && value != 3) return fail(MGCMT_ERR_ARG, "uni_minctas must be 0 (default), 2 or 3");
    mgcmt::g_uni_minctas = value;
    return MGCMT_OK;
  }
  if (!strcmp(name, "fused_c9")) {
    if (value != 0 && value != 2 && value != 4) return fail(MGCMT_ERR_ARG, "fused_c9 must be 0 (auto), 2 or 4");
    mgcmt::g_fused_c9 = value;
    return MGCMT_OK;
  }
  if (!strcmp(name, "fused_c5")) {
    if (value != 2 && value != 4) return fail(MGCMT_ERR_ARG, "fused_c5 must be 2 or 4");
    mgcmt::g_fused_c5 = value;
    return MGCMT_OK;
  }
  if (!strcmp(name, "band_gs_scan")) { mgcmt::g_band_gs_scan = value ? 1 : 0; return MGCMT_OK; }
  if (!strcmp(name, "band_gs_split")) { mgcmt::g_band_gs_split = value ? 1 : 0; return MGCMT_OK; }
  return fail(MGCMT_ERR_ARG, std::string("unknown option ") + name);
}

int mgcmt_fused_leg(mgcmt_hier_t *h, int level, int mode, int nu, double shift, double omega, const double *d_vin,
                    const double *d_f, double *d_vout, const double *d_ecoarse, double *d_rcoarse, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (level + 1 >= h->nlev && mode != FUSED_SMOOTH && !h->slab) return fail(MGCMT_ERR_ARG, "no coarser level");
  if (h->slab && (mode & 16)) return fail(MGCMT_ERR_ARG, "slab levels use the streaming legs");
  if (!h->coarsen_rows || h->lev[level].dev.nrows < 2) return fail(MGCMT_ERR_ARG, "fused legs are 2-D only");
  const bool force_tile = (mode & 16) != 0;  // bit 4: use the shared-memory tile implementation
  const bool gs_leg = (mode & 32) != 0;      // bit 5: nu = Gauss-Seidel colour sweeps instead of Jacobi sweeps
  mode &= 15;
  if (nu < 0 || nu > 4 || mode < 0 || mode > 3) return fail(MGCMT_ERR_ARG, "bad fused leg mode / nu");
  if (gs_leg && (nu < 1 || nu > ((h->lev[level].dev.five || force_tile) ? 4 : 2)))
    return fail(MGCMT_ERR_ARG, "Gauss-Seidel legs: 1..4 sweeps per pass on the 5-point level and on tiles, 1..2 on streamed 9-point levels");
  if (d_vin == d_vout) return fail(MGCMT_ERR_ARG, "fused legs are out of place");
  NEED_ALIGNED(d_f, d_vout);
  if (mode != FUSED_DOWN_ZERO) NEED_ALIGNED(d_vin);
  if (mode == FUSED_UP) NEED_ALIGNED(d_ecoarse);
  if (mode == FUSED_DOWN || mode == FUSED_DOWN_ZERO) NEED_ALIGNED(d_rcoarse);
  if (gs_leg && force_tile)
    CU(launch_tile_gs_leg(h->lev[level].dev, mode, nu, shift, omega, d_vin, d_f, d_vout, d_ecoarse, d_rcoarse,
                          (cudaStream_t)stream));
  else if (gs_leg)
    CU(launch_fused_gs_leg(h->lev[level].dev, mode, nu, shift, omega, d_vin, d_f, d_vout, d_ecoarse, d_rcoarse,
                           (cudaStream_t)stream));
  else if (force_tile)
    CU(launch_tile_leg(h->lev[level].dev, mode, nu, shift, omega, d_vin, d_f, d_vout, d_ecoarse, d_rcoarse,
                       (cudaStream_t)stream));
  else
    CU(launch_fused_leg(h->lev[level].dev, mode, nu, shift, omega, d_vin, d_f, d_vout, d_ecoarse, d_rcoarse,
                        (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_vcycle(mgcmt_hier_t *h, double shift, int nu1, int nu2, int smoother, double omega, double *d_v,
                 const double *d_f, int v0_is_zero, void *stream) {
  if (!h) return fail(MGCMT_ERR_ARG, "null hierarchy");
  if (!d_v || !d_f || d_v == d_f) return fail(MGCMT_ERR_ARG, "need distinct non-null v and f");
  NEED_ALIGNED(d_v, d_f);
  if (nu1 < 0 || nu2 < 0) return fail(MGCMT_ERR_ARG, "negative sweep count");
  if (h->slab || h->first_work > 0) return fail(MGCMT_ERR_STATE, "this hierarchy has no full finest level (slab piece / coarse part)");
  return vcycle_level(h, 0, shift, nu1, nu2, smoother, omega, d_v, d_f, v0_is_zero != 0, (cudaStream_t)stream);
}

int mgcmt_vcycle_host_block(mgcmt_hier_t *h, int k, const double *shifts, int nu1, int nu2, int smoother, double omega,
                            const double *const *f_host, double *const *v_host, void *stream) {
  if (!h) return fail(MGCMT_ERR_ARG, "null hierarchy");
  if (k < 0 || (k > 0 && (!shifts || !f_host || !v_host))) return fail(MGCMT_ERR_ARG, "bad block arguments");
  if (nu1 < 0 || nu2 < 0) return fail(MGCMT_ERR_ARG, "negative sweep count");
  if (h->slab || h->first_work > 0) return fail(MGCMT_ERR_STATE, "this hierarchy has no full finest level (slab piece / coarse part)");
  for (int c = 0; c < k; ++c)
    if (!f_host[c] || !v_host[c] || f_host[c] == v_host[c]) return fail(MGCMT_ERR_ARG, "need distinct non-null host vectors");
  if (k == 0) return MGCMT_OK;
  const size_t n = (size_t)h->lev[0].dev.nrows * (size_t)h->lev[0].dev.ncols;
  const size_t bytes = n * sizeof(double);
  mgcmt_hier::HostPipe &P = h->pipe;
  const int nslot = k < 3 ? k : 3;
  if (!P.ready) {
    CU(cudaStreamCreateWithFlags(&P.in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&P.out, cudaStreamNonBlocking));
    for (int i = 0; i < 3; ++i) {
      CU(cudaEventCreateWithFlags(&P.uploaded[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&P.cycled[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&P.downloaded[i], cudaEventDisableTiming));
    }
    P.ready = true;
  }
  for (int i = 0; i < nslot; ++i) {
    if (!P.f[i]) CU(cudaMalloc(&P.f[i], bytes));
    if (!P.v[i]) CU(cudaMalloc(&P.v[i], bytes));
  }
  cudaStream_t s = (cudaStream_t)stream;
  // upload of vector c+1 and download of vector c-1 run beside cycle c (PCIe is full duplex); a slot is reused once its
  // previous result has left the device
  for (int c = 0; c < k; ++c) {
    const int i = c % 3;
    if (c >= 3) CU(cudaStreamWaitEvent(P.in, P.downloaded[i], 0));
    CU(staged_upload(P.f[i], f_host[c], bytes, P.in));   // pageable sources: threaded staging through page-locked chunks
    CU(cudaEventRecord(P.uploaded[i], P.in));
    CU(cudaStreamWaitEvent(s, P.uploaded[i], 0));
    int rc = vcycle_level(h, 0, shifts[c], nu1, nu2, smoother, omega, P.v[i], P.f[i], true, s);
    if (rc) {
      cudaStreamSynchronize(P.in);
      cudaStreamSynchronize(P.out);
      return rc;
    }
    CU(cudaEventRecord(P.cycled[i], s));
    CU(cudaStreamWaitEvent(P.out, P.cycled[i], 0));
    CU(cudaMemcpyAsync(v_host[c], P.v[i], bytes, cudaMemcpyDeviceToHost, P.out));
    CU(cudaEventRecord(P.downloaded[i], P.out));
  }
  CU(cudaStreamSynchronize(P.out));   // every result is in host memory (the last download follows all cycles)
  return MGCMT_OK;
}

int mgcmt_vcycle_rq(mgcmt_hier_t *h, double shift, int nu1, int nu2, int smoother, double omega, double *d_v,
                    const double *d_f, int v0_is_zero, double *d_out2, void *stream) {
  if (!h) return fail(MGCMT_ERR_ARG, "null hierarchy");
  if (!d_out2) return fail(MGCMT_ERR_ARG, "null output");
  h->rq_out = d_out2;
  h->rq_done = false;
  int rc = mgcmt_vcycle(h, shift, nu1, nu2, smoother, omega, d_v, d_f, v0_is_zero, stream);
  h->rq_out = nullptr;
  if (rc) return rc;
  if (h->rq_done) {
    // the fused stage summed w^T (A - shift I) w: add shift * w^T w
    CU(launch_rq_unshift(d_out2, shift, (cudaStream_t)stream));
    return MGCMT_OK;
  }
  return mgcmt_rayleigh(h, 0, d_v, d_out2, stream);  // levels / smoothers without the fused stage: one extra pass
}

int mgcmt_vcycle_from(mgcmt_hier_t *h, int level, double shift, int smoother, double omega, double *d_v,
                      const double *d_f, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (h->slab || level < h->first_work) return fail(MGCMT_ERR_STATE, "level has no work vectors in this hierarchy");
  if (!d_v || !d_f || d_v == d_f) return fail(MGCMT_ERR_ARG, "need distinct non-null v and f");
  NEED_ALIGNED(d_v, d_f);
  // the V-cycle restricted to levels level..coarsest, zero initial guess, 4/4 sweeps: what MGCMTSolver.vcycle
  // does at every coarse level (MGCMTSolver.py:316-320)
  return vcycle_level(h, level, shift, 4, 4, smoother, omega, d_v, d_f, true, (cudaStream_t)stream);
}

int mgcmt_slab_up_rq(mgcmt_hier_t *h, int gs, double shift, double omega, const double *d_vin, const double *d_f, double *d_vout,
                     const double *d_ecoarse, double *d_out2, void *stream) {
  int rc = check_level(h, 0);
  if (rc) return rc;
  if (!h->slab) return fail(MGCMT_ERR_STATE, "not a slab hierarchy");
  if (!d_out2) return fail(MGCMT_ERR_ARG, "null output");
  if (d_vin == d_vout) return fail(MGCMT_ERR_ARG, "fused legs are out of place");
  NEED_ALIGNED(d_vin, d_f, d_vout, d_ecoarse);
  Level &L = h->lev[0];
  cudaStream_t s = (cudaStream_t)stream;
  const int slots = fused_rq_slots(L.dev, gs ? 1 : 0);
  if (slots <= 0) return fail(MGCMT_ERR_STATE, "the fused Rayleigh stage is not available for this level");
  if (slots > h->rq_slots) {
    cudaFree(h->rq_partials);
    h->rq_partials = nullptr;
    CU(cudaMalloc(&h->rq_partials, sizeof(double) * 2 * slots));
    h->rq_slots = slots;
  }
  // the halo row next to the owned rows is exact in the output (dependency cone of prolongation + 4 Jacobi sweeps = 5 rows,
  // + 8 colour stages = 9 rows; the slab arrays carry 10 halo rows), so (A w) on the first and last owned row is too:
  // the sums over the owned rows need no second exchange
  if (gs) CU(launch_fused_gs_leg(L.dev, FUSED_UP_RQ, 4, shift, omega, d_vin, d_f, d_vout, d_ecoarse, h->rq_partials, s));
  else CU(launch_fused_leg(L.dev, FUSED_UP_RQ, 4, shift, omega, d_vin, d_f, d_vout, d_ecoarse, h->rq_partials, s));
  CU(launch_finish(2, slots, h->rq_partials, d_out2, s));
  CU(launch_rq_unshift(d_out2, shift, s));
  return MGCMT_OK;
}

int mgcmt_slab_rayleigh(mgcmt_hier_t *h, int level, const double *d_x, double *d_out2, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (!h->slab) return fail(MGCMT_ERR_STATE, "not a slab hierarchy");
  if (!d_out2) return fail(MGCMT_ERR_ARG, "null output");
  NEED_ALIGNED(d_x);
  Level &L = h->lev[level];
  cudaStream_t s = (cudaStream_t)stream;
  // view of the owned rows with the neighbouring halo rows as Dirichlet / halo inputs
  LevelDev own = L.dev;
  const int H = h->halo, nown = L.dev.nrows - 2 * H, begin = L.dev.row0 + H;
  own.nrows = nown;
  own.row0 = begin;
  const double *x0 = d_x + (size_t)H * L.dev.ncols;
  const double *top = begin > 0 ? d_x + (size_t)(H - 1) * L.dev.ncols : nullptr;
  const double *bot = (begin + nown < L.dev.nrows_glob) ? d_x + (size_t)(H + nown) * L.dev.ncols : nullptr;
  const int nb = march_grid_blocks(own);
  if ((size_t)2 * nb > L.n) return fail(MGCMT_ERR_STATE, "scratch too small");
  CU(launch_rayleigh_partials(own, x0, L.tmp, top, bot, s));
  CU(launch_finish(2, nb, L.tmp, d_out2, s));
  return MGCMT_OK;
}

int mgcmt_dot(long long n, const double *d_x, const double *d_y, double *d_out, void *stream) {
  if (n < 0 || !d_x || !d_y || !d_out) return fail(MGCMT_ERR_ARG, "bad dot arguments");
  Scratch *sc;
  int rc = get_scratch(&sc);
  if (rc) return rc;
  CU(launch_dot(n, d_x, d_y, sc->partials, d_out, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_rayleigh(mgcmt_hier_t *h, int level, const double *d_x, double *d_out2, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (!d_out2) return fail(MGCMT_ERR_ARG, "null output");
  NEED_ALIGNED(d_x);
  Scratch *sc;
  rc = get_scratch(&sc);
  if (rc) return rc;
  Level &L = h->lev[level];
  cudaStream_t s = (cudaStream_t)stream;
  if (!L.tmp || h->slab) return fail(MGCMT_ERR_STATE, "level has no scratch vector in this hierarchy");
  // one pass: the operator-apply kernel keeps x^T(Ax) and x^T x partial sums per CTA (L.tmp as scratch),
  // then one ordered finish
  const int nb = march_grid_blocks(L.dev);
  if ((size_t)2 * nb > L.n) {
    // tiny level: L.tmp is too small for the per-CTA partials of the fused pass.  Plain route with partial sums that
    // belong to this hierarchy (allocated on first use), so that concurrent calls on different hierarchies / streams
    // (ShiftMethod runs one per stream) never share reduction scratch.
    if (!h->small_partials) CU(cudaMalloc(&h->small_partials, sizeof(double) * 2 * kReduceBlocks));
    CU(launch_apply(L.dev, 0.0, d_x, L.tmp, nullptr, nullptr, s));
    CU(launch_dot((long long)L.n, L.tmp, d_x, h->small_partials, d_out2, s));
    CU(launch_dot((long long)L.n, d_x, d_x, h->small_partials + kReduceBlocks, d_out2 + 1, s));
    return MGCMT_OK;
  }
  CU(launch_rayleigh_partials(L.dev, d_x, L.tmp, nullptr, nullptr, s));
  CU(launch_finish(2, nb, L.tmp, d_out2, s));
  return MGCMT_OK;
}

int mgcmt_eigen_residual(mgcmt_hier_t *h, int level, const double *d_x, const double *d_rq2, double *d_r,
                         double *d_out_sumsq, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (!d_rq2 || !d_out_sumsq) return fail(MGCMT_ERR_ARG, "null scalar pointer");
  if (h->slab) return fail(MGCMT_ERR_STATE, "not available on slab pieces");
  NEED_ALIGNED(d_x, d_r);
  Level &L = h->lev[level];
  if (!L.tmp) return fail(MGCMT_ERR_STATE, "level has no scratch vector in this hierarchy");
  cudaStream_t s = (cudaStream_t)stream;
  CU(launch_apply(L.dev, 0.0, d_x, d_r, nullptr, nullptr, s));
  CU(launch_axpy_dev((long long)L.n, d_rq2, d_rq2 + 1, -1.0, d_x, d_r, s));  // r -= (num / den) x
  // per-hierarchy scratch (L.tmp) for the partial sums: callable concurrently on different hierarchies / streams
  if ((size_t)kReduceBlocks > L.n) return fail(MGCMT_ERR_STATE, "level too small");
  CU(launch_dot((long long)L.n, d_r, d_r, L.tmp, d_out_sumsq, s));
  return MGCMT_OK;
}

int mgcmt_normalize(long long n, double *d_x, void *stream) {
  if (n < 0 || !d_x) return fail(MGCMT_ERR_ARG, "bad normalize arguments");
  Scratch *sc;
  int rc = get_scratch(&sc);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  CU(launch_dot(n, d_x, d_x, sc->partials, sc->scal, s));
  CU(launch_scale_by_inv_norm(n, d_x, sc->scal, s));
  return MGCMT_OK;
}

int mgcmt_gram(long long n, int k, const double *d_V, long long stride, double *d_out, void *stream) {
  if (n <= 0 || k < 1 || k > 6 || (n & 1) || !d_V || !d_out || !al16(d_V) || (stride & 1))
    return fail(MGCMT_ERR_ARG, "bad gram arguments (k <= 6, even n and stride, aligned vectors)");
  Scratch *sc;
  int rc = get_scratch(&sc);
  if (rc) return rc;
  CU(launch_gram(n, k, d_V, stride, sc->partials, d_out, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_cholqr_apply(long long n, int k, double *d_V, long long stride, const double *d_gram, void *stream) {
  if (n <= 0 || k < 1 || k > 6 || (n & 1) || !d_V || !d_gram || !al16(d_V) || (stride & 1))
    return fail(MGCMT_ERR_ARG, "bad cholqr arguments");
  Scratch *sc;
  int rc = get_scratch(&sc);
  if (rc) return rc;
  CU(launch_chol_inverse(k, d_gram, sc->scal + 24, sc->status, (cudaStream_t)stream));
  CU(launch_cholqr_apply(n, k, d_V, stride, sc->scal + 24, (cudaStream_t)stream));
  return MGCMT_OK;
}

long long mgcmt_rqmin_work_doubles(mgcmt_hier_t *h, int level, int use_mass) {
  if (check_level(h, level)) return -1;
  return (long long)(use_mass ? 7 : 4) * (long long)h->lev[level].n + 32 + 8 * kReduceBlocks;
}

int mgcmt_rqmin(mgcmt_hier_t *h, int level, int use_mass, double *d_x, int nu, double *d_work, long long work_doubles,
                double *d_rq2, void *stream) {
  int rc = check_level(h, level);
  if (rc) return rc;
  if (h->slab) return fail(MGCMT_ERR_STATE, "not available on slab pieces");
  if (!d_work || !d_rq2 || nu < 0) return fail(MGCMT_ERR_ARG, "bad rqmin arguments");
  if (work_doubles < mgcmt_rqmin_work_doubles(h, level, use_mass)) return fail(MGCMT_ERR_ARG, "rqmin work array too small");
  NEED_ALIGNED(d_x, d_work);
  Level &L = h->lev[level];
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)L.n;
  const bool mass = use_mass != 0;
  // carve the work array (vector sizes rounded up to keep 16-byte alignment)
  const long long nn = (n + 1) & ~1LL;
  double *Ax = d_work, *g = Ax + nn, *p = g + nn, *Ap = p + nn;
  double *Mx = mass ? Ap + nn : d_x, *Mp = mass ? Mx + nn : p, *Mg = mass ? Mp + nn : g;
  double *scal = d_work + (mass ? 7 : 4) * nn;
  double *partials = scal + 32;
  if ((mass ? 7 : 4) * nn + 32 + 8 * kReduceBlocks > work_doubles + 8) return fail(MGCMT_ERR_ARG, "rqmin work array too small");
  auto apply_mass = [&](const double *x, double *y) -> int { return mgcmt_apply_mass(h, level, x, y, stream); };
  CU(launch_apply(L.dev, 0.0, d_x, Ax, nullptr, nullptr, s));
  if (mass) { rc = apply_mass(d_x, Mx); if (rc) return rc; }
  CU(launch_rq_sums(n, d_x, Ax, Mx, partials, scal, s));
  CU(cudaMemsetAsync(scal + 10, 0, 2 * sizeof(double), s));
  CU(launch_rq_grad(n, mass, Ax, Mx, g, partials, scal, s));
  if (mass) { rc = apply_mass(g, Mg); if (rc) return rc; CU(launch_dot(n, g, Mg, partials, scal + 10, s)); }
  for (int it = 0; it < nu; ++it) {
    CU(launch_rq_dir(n, it == 0, scal, g, p, s));
    CU(launch_apply(L.dev, 0.0, p, Ap, nullptr, nullptr, s));
    if (mass) { rc = apply_mass(p, Mp); if (rc) return rc; }
    CU(launch_rq_pencil(n, d_x, p, Ax, Ap, Mx, Mp, partials, scal, s));
    CU(launch_rq_update(n, mass, d_x, p, Ax, Ap, Mx, Mp, partials, scal, s));
    CU(launch_rq_grad(n, mass, Ax, Mx, g, partials, scal, s));
    if (mass) { rc = apply_mass(g, Mg); if (rc) return rc; CU(launch_dot(n, g, Mg, partials, scal + 10, s)); }
  }
  CU(cudaMemcpyAsync(d_rq2, scal + 12, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return MGCMT_OK;
}

int mgcmt_ortho_status(int *h_flag, void *stream) {
  if (!h_flag) return fail(MGCMT_ERR_ARG, "null flag pointer");
  Scratch *sc;
  int rc = get_scratch(&sc);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(h_flag, sc->status, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaMemsetAsync(sc->status, 0, sizeof(int), s));
  CU(cudaStreamSynchronize(s));
  return MGCMT_OK;
}

int mgcmt_scale_inv_norm(long long n, double *d_x, const double *d_sumsq, void *stream) {
  if (n < 0 || !d_x || !d_sumsq) return fail(MGCMT_ERR_ARG, "bad scale arguments");
  CU(launch_scale_by_inv_norm(n, d_x, d_sumsq, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_axpy_dev(long long n, const double *d_alpha, double sign, const double *d_x, double *d_y,
                   void *stream) {
  if (n < 0 || !d_alpha || !d_x || !d_y) return fail(MGCMT_ERR_ARG, "bad axpy arguments");
  CU(launch_axpy_dev(n, d_alpha, nullptr, sign, d_x, d_y, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_gramschmidt(long long n, int k, double *d_V, int modified, void *stream) {
  if (n <= 0 || k <= 0 || !d_V) return fail(MGCMT_ERR_ARG, "bad gramschmidt arguments");
  Scratch *sc;
  int rc = get_scratch(&sc);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  double *scal = sc->scal;  // [0] sumsq / <q,q>, [1..16] dots, [32..] <u_i,u_i> (classical)
  if (modified == 2) {
    // Gram-matrix (Cholesky-QR) form: same Q in exact arithmetic, 3k vector passes (see reduce.cu)
    if (k > 6 || (n & 1) || !al16(d_V)) return fail(MGCMT_ERR_ARG, "Gram-matrix orthonormalisation needs k <= 6, even n, aligned block");
    CU(launch_gram(n, k, d_V, n, sc->partials, scal, s));                 // scal[0..k(k+1)/2)
    CU(launch_chol_inverse(k, scal, scal + 24, sc->status, s));           // scal[24..24+k*k)
    CU(launch_cholqr_apply(n, k, d_V, n, scal + 24, s));
    return MGCMT_OK;
  }
  if (modified && (n % 2 == 0) && al16(d_V) && k <= 8) {
    // MGCMTProcessor.py:44-50, two fused passes per column (reduce.cu): 29 instead of 42 vector passes at k = 4
    CU(launch_dot(n, d_V, d_V, sc->partials, scal, s));  // ||w_0||^2
    for (int i = 0; i < k; ++i) {
      double *qi = d_V + (size_t)i * n;
      const int m = k - 1 - i;
      CU(launch_mgs_scale_dots(n, m, qi, scal, qi + n, n, sc->partials, scal + 1, s));  // scal[1] = <q,q>, scal[2..] dots
      if (m > 0) CU(launch_mgs_update(n, m, qi, scal + 1, qi + n, n, sc->partials, scal, s));  // scal[0] = ||w_{i+1}||^2
    }
    return MGCMT_OK;
  }
  if (modified) {
    // MGCMTProcessor.py:44-50
    for (int i = 0; i < k; ++i) {
      double *qi = d_V + (size_t)i * n;
      CU(launch_dot(n, qi, qi, sc->partials, scal, s));
      CU(launch_scale_by_inv_norm(n, qi, scal, s));
      if (i + 1 == k) break;
      CU(launch_dot(n, qi, qi, sc->partials, scal, s));  // <q_i, q_i> (the reference divides by it)
      for (int j0 = i + 1; j0 < k; j0 += 16) {
        const int m = (k - j0 < 16) ? (k - j0) : 16;
        CU(launch_multidot(n, m, d_V + (size_t)j0 * n, n, qi, sc->partials, scal + 1, s));
        for (int j = 0; j < m; ++j)
          CU(launch_axpy_dev(n, scal + 1 + j, scal, -1.0, qi, d_V + (size_t)(j0 + j) * n, s));
      }
    }
    return MGCMT_OK;
  }
  // classical: u_j = v_j - sum_{i<j} (<v_j,u_i>/<u_i,u_i>) u_i with the ORIGINAL v_j in every inner
  // product, then normalise all columns (MGCMTProcessor.py:35-42)
  if (k > 17) return fail(MGCMT_ERR_ARG, "classical Gram-Schmidt supports k <= 17");
  double *uu = scal + 32;
  for (int j = 0; j < k; ++j) {
    double *vj = d_V + (size_t)j * n;
    if (j > 0) {  // all coefficients first (they use the original v_j), then the subtractions in order
      CU(launch_multidot(n, j, d_V, n, vj, sc->partials, scal + 1, s));
      for (int i = 0; i < j; ++i)
        CU(launch_axpy_dev(n, scal + 1 + i, uu + i, -1.0, d_V + (size_t)i * n, vj, s));
    }
    CU(launch_dot(n, vj, vj, sc->partials, uu + j, s));
  }
  for (int j = 0; j < k; ++j) CU(launch_scale_by_inv_norm(n, d_V + (size_t)j * n, uu + j, s));
  return MGCMT_OK;
}

}  // extern "C"
