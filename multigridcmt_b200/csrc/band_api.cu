// band_api.cu -- C ABI of the general banded complex128 path (include/mgcmt_b200.h, "banded operators").
// Hierarchy of diagonal-stored operators A_l = R A_{l-1} P built on the device, the reference's V-cycle
// (MGCMTSolver.py:281-329) over it with wjacobi / gseidel / sor, dense complex coarsest solve cached per shift.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <set>
#include <string>
#include <vector>

#include "../../include/mgcmt_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace mgcmt;

namespace {

struct BandLevel {
  BandDev dev{};
  std::vector<int> offs;   // host copy
  int *d_offs = nullptr;
  double *d_vals = nullptr;  // 2 * ndiag * n doubles
  double *v = nullptr, *f = nullptr, *tmp = nullptr, *y = nullptr, *g = nullptr;  // 2 * n doubles each
};

struct BandInv {
  double shift;
  double *aug;  // n x 2n complex
  uint64_t stamp;
};

constexpr size_t kBandInvCache = 8;

}  // namespace

struct mgcmt_band {
  std::vector<BandLevel> lev;
  std::vector<BandInv> inv;
  uint64_t clock = 0;
  int *status = nullptr;
};

namespace {

#define CU(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return set_error(MGCMT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)

bool al16(const void *p) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int level_ok(const mgcmt_band *h, int level) {
  if (!h) return set_error(MGCMT_ERR_ARG, "null banded hierarchy");
  if (level < 0 || level >= (int)h->lev.size()) return set_error(MGCMT_ERR_ARG, "level out of range");
  return MGCMT_OK;
}

// inverse of (A_coarsest - shift I), cached per shift (the eigen-iteration keeps its shifts fixed)
int band_inverse(mgcmt_band *h, double shift, cudaStream_t s, const double **out) {
  ++h->clock;
  for (BandInv &e : h->inv)
    if (e.shift == shift) {
      e.stamp = h->clock;
      *out = e.aug;
      return MGCMT_OK;
    }
  const BandLevel &L = h->lev.back();
  const size_t bytes = sizeof(double) * 4 * (size_t)L.dev.n * L.dev.n;
  BandInv e{shift, nullptr, h->clock};
  if (h->inv.size() >= kBandInvCache) {
    size_t victim = 0;
    for (size_t i = 1; i < h->inv.size(); ++i)
      if (h->inv[i].stamp < h->inv[victim].stamp) victim = i;
    CU(cudaStreamSynchronize(s));
    e.aug = h->inv[victim].aug;
    h->inv.erase(h->inv.begin() + victim);
  } else {
    CU(cudaMalloc(&e.aug, bytes));
  }
  cudaError_t ce = cudaMemsetAsync(h->status, 0, sizeof(int), s);
  if (ce == cudaSuccess) ce = launch_band_inverse(L.dev, shift, e.aug, h->status, s);
  int st = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&st, h->status, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess || st != 0) {
    cudaFree(e.aug);
    if (ce != cudaSuccess) return set_error(MGCMT_ERR_CUDA, std::string("coarsest inverse: ") + cudaGetErrorString(ce));
    return set_error(MGCMT_ERR_NUMERIC, "coarsest operator minus shift is singular (pivot " + std::to_string(st) + ")");
  }
  h->inv.push_back(e);
  *out = e.aug;
  return MGCMT_OK;
}

int band_smooth(mgcmt_band *h, int level, int smoother, int nu, double shift, double omega, double *v, const double *f,
                cudaStream_t s) {
  BandLevel &L = h->lev[level];
  const size_t bytes = sizeof(double) * 2 * (size_t)L.dev.n;
  if (nu <= 0) return MGCMT_OK;
  if (smoother == MGCMT_SMOOTH_WJACOBI) {
    double *a = v, *b = L.tmp;
    for (int i = 0; i < nu; ++i) {
      CU(launch_band_jacobi(L.dev, shift, omega, a, f, b, s));
      std::swap(a, b);
    }
    if (a != v) CU(cudaMemcpyAsync(v, a, bytes, cudaMemcpyDeviceToDevice, s));
    return MGCMT_OK;
  }
  if (smoother == MGCMT_SMOOTH_GSLEX) {
    if (omega == 1.0) {
      for (int i = 0; i < nu; ++i) CU(launch_band_lower_solve(L.dev, shift, 1.0, 1.0, 0.0, 1.0, 1.0, v, f, v, nullptr, v, L.tmp, s));
      return MGCMT_OK;
    }
    // sor, quirk Q6: v <- (D - wL)^-1 ((1-w) D + w U) v + w (D - L)^-1 f
    CU(launch_band_lower_solve(L.dev, shift, 1.0, 1.0, 0.0, 0.0, omega, v, f, L.y, nullptr, L.g, L.tmp, s));
    for (int i = 0; i < nu; ++i)
      CU(launch_band_lower_solve(L.dev, shift, omega, 0.0, 1.0 - omega, omega, 1.0, v, f, L.y, L.g, v, L.tmp, s));
    return MGCMT_OK;
  }
  return set_error(MGCMT_ERR_ARG, "banded operators take the wjacobi, gseidel and sor smoothers (red-black needs a radius-1 stencil)");
}

int band_cycle(mgcmt_band *h, int level, double shift, int nu1, int nu2, int smoother, double omega, double *v,
               const double *f, cudaStream_t s) {
  const int last = (int)h->lev.size() - 1;
  BandLevel &L = h->lev[level];
  if (level == last) {
    const double *aug = nullptr;
    int rc = band_inverse(h, shift, s, &aug);
    if (rc != MGCMT_OK) return rc;
    CU(launch_band_gemv(L.dev.n, aug, f, v, s));
    return MGCMT_OK;
  }
  int rc = band_smooth(h, level, smoother, nu1, shift, omega, v, f, s);
  if (rc != MGCMT_OK) return rc;
  BandLevel &C = h->lev[level + 1];
  CU(launch_band_residual_restrict(L.dev, shift, v, f, C.f, s));
  CU(cudaMemsetAsync(C.v, 0, sizeof(double) * 2 * (size_t)C.dev.n, s));
  rc = band_cycle(h, level + 1, shift, 4, 4, smoother, omega, C.v, C.f, s);  // quirk Q4: coarse levels run 4/4
  if (rc != MGCMT_OK) return rc;
  CU(launch_band_prolong_correct(L.dev.n, C.v, v, s));
  return band_smooth(h, level, smoother, nu2, shift, omega, v, f, s);
}

}  // namespace

extern "C" {

int mgcmt_band_destroy(mgcmt_band_t *h) {
  if (!h) return MGCMT_OK;
  for (BandLevel &L : h->lev) {
    cudaFree(L.d_offs);
    cudaFree(L.d_vals);
    cudaFree(L.v);
    cudaFree(L.f);
    cudaFree(L.tmp);
    cudaFree(L.y);
    cudaFree(L.g);
  }
  for (BandInv &e : h->inv) cudaFree(e.aug);
  cudaFree(h->status);
  delete h;
  return MGCMT_OK;
}

int mgcmt_band_create(int n, int ndiag, const int *h_offsets, const double *d_vals, int lowest_level, void *stream,
                      mgcmt_band_t **out) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!out || !h_offsets || !d_vals) return set_error(MGCMT_ERR_ARG, "null argument");
  if (n < 2 || lowest_level < 2 || lowest_level > n) return set_error(MGCMT_ERR_ARG, "need 2 <= lowest_level <= n");
  if (ndiag < 1 || ndiag > kBandMaxDiags) return set_error(MGCMT_ERR_ARG, "1..96 diagonals");
  int nlev = 1;
  for (int m = n; m != lowest_level; m >>= 1, ++nlev)
    if (m < lowest_level || (m & 1)) return set_error(MGCMT_ERR_ARG, "n must be lowest_level times a power of two");
  if (lowest_level > kBandMaxCoarse) return set_error(MGCMT_ERR_ARG, "coarsest level of a banded operator is limited to 512 unknowns");
  std::vector<int> offs(h_offsets, h_offsets + ndiag);
  int idiag = -1;
  for (int k = 0; k < ndiag; ++k) {
    if (k && offs[k] <= offs[k - 1]) return set_error(MGCMT_ERR_ARG, "offsets must be strictly ascending");
    if (offs[k] <= -n || offs[k] >= n) return set_error(MGCMT_ERR_ARG, "offset outside the matrix");
    if (offs[k] == 0) idiag = k;
  }
  if (idiag < 0) return set_error(MGCMT_ERR_ARG, "the main diagonal (offset 0) must be stored");
  if (!al16(d_vals)) return set_error(MGCMT_ERR_ARG, "d_vals must be 16-byte aligned");

  mgcmt_band *h = new mgcmt_band();
  h->lev.resize(nlev);
  auto bail = [&](int code, const std::string &msg) {
    mgcmt_band_destroy(h);
    return set_error(code, msg);
  };
#define CUB(expr)                                                                                     \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess) return bail(MGCMT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)
  CUB(cudaMalloc(&h->status, sizeof(int)));
  int m = n;
  for (int l = 0; l < nlev; ++l, m >>= 1) {
    BandLevel &L = h->lev[l];
    if (l == 0) {
      L.offs = offs;
    } else {
      // offset d of the fine operator feeds coarse offsets D with |2D - d| <= 2 (band.cu: band_galerkin_kernel)
      std::set<int> cs;
      auto floor2 = [](int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); };
      for (int d : h->lev[l - 1].offs)
        for (int D = floor2(d - 1); D <= floor2(d + 2); ++D)   // ceil((d-2)/2) .. floor((d+2)/2)
          if (D > -m && D < m) cs.insert(D);
      L.offs.assign(cs.begin(), cs.end());
      if ((int)L.offs.size() > kBandMaxDiags) return bail(MGCMT_ERR_ARG, "coarse operator has more than 96 diagonals");
    }
    const int nd = (int)L.offs.size();
    const size_t vec = sizeof(double) * 2 * (size_t)m;
    CUB(cudaMalloc(&L.d_offs, sizeof(int) * nd));
    CUB(cudaMemcpyAsync(L.d_offs, L.offs.data(), sizeof(int) * nd, cudaMemcpyHostToDevice, s));
    CUB(cudaMalloc(&L.d_vals, vec * nd));
    CUB(cudaMalloc(&L.v, vec));
    CUB(cudaMalloc(&L.f, vec));
    CUB(cudaMalloc(&L.tmp, vec));
    CUB(cudaMalloc(&L.y, vec));
    CUB(cudaMalloc(&L.g, vec));
    L.dev.n = m;
    L.dev.ndiag = nd;
    L.dev.idiag = (int)(std::find(L.offs.begin(), L.offs.end(), 0) - L.offs.begin());
    L.dev.offs = L.d_offs;
    L.dev.vals = (const double2 *)L.d_vals;
    if (l == 0) {
      CUB(cudaMemcpyAsync(L.d_vals, d_vals, vec * nd, cudaMemcpyDeviceToDevice, s));
    } else {
      const BandLevel &F = h->lev[l - 1];
      std::vector<int> lut((size_t)nd * 5, -1);
      for (int kc = 0; kc < nd; ++kc)
        for (int t = 0; t < 5; ++t) {
          const int d = 2 * L.offs[kc] + t - 2;
          auto it = std::lower_bound(F.offs.begin(), F.offs.end(), d);
          if (it != F.offs.end() && *it == d) lut[(size_t)kc * 5 + t] = (int)(it - F.offs.begin());
        }
      int *d_lut = nullptr;
      CUB(cudaMalloc(&d_lut, sizeof(int) * lut.size()));
      cudaError_t ce = cudaMemcpyAsync(d_lut, lut.data(), sizeof(int) * lut.size(), cudaMemcpyHostToDevice, s);
      if (ce == cudaSuccess) ce = launch_band_galerkin(F.dev, m, nd, L.d_offs, d_lut, L.d_vals, s);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);  // lut (host + device) is released below
      cudaFree(d_lut);
      CUB(ce);
    }
  }
  CUB(cudaStreamSynchronize(s));  // L.offs host vectors were the source of async copies
#undef CUB
  *out = h;
  return MGCMT_OK;
}

int mgcmt_band_num_levels(const mgcmt_band_t *h, int *out) {
  if (!h || !out) return set_error(MGCMT_ERR_ARG, "null argument");
  *out = (int)h->lev.size();
  return MGCMT_OK;
}

int mgcmt_band_level_shape(const mgcmt_band_t *h, int level, int *n, int *ndiag) {
  int rc = level_ok(h, level);
  if (rc != MGCMT_OK) return rc;
  if (n) *n = h->lev[level].dev.n;
  if (ndiag) *ndiag = h->lev[level].dev.ndiag;
  return MGCMT_OK;
}

int mgcmt_band_level_diags(const mgcmt_band_t *h, int level, int *h_offsets, double *h_vals) {
  int rc = level_ok(h, level);
  if (rc != MGCMT_OK) return rc;
  const BandLevel &L = h->lev[level];
  if (h_offsets) std::copy(L.offs.begin(), L.offs.end(), h_offsets);
  if (h_vals) {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h_vals, L.d_vals, sizeof(double) * 2 * (size_t)L.dev.n * L.dev.ndiag, cudaMemcpyDeviceToHost));
  }
  return MGCMT_OK;
}

#define NEED_VEC(...)                                                                                          \
  do {                                                                                                         \
    const void *ps__[] = {__VA_ARGS__};                                                                        \
    for (const void *p__ : ps__)                                                                               \
      if (!al16(p__)) return set_error(MGCMT_ERR_ARG, "complex vectors must be non-null and 16-byte aligned"); \
  } while (0)

int mgcmt_band_apply(mgcmt_band_t *h, int level, double shift, const double *d_x, double *d_y, void *stream) {
  int rc = level_ok(h, level);
  if (rc != MGCMT_OK) return rc;
  NEED_VEC(d_x, d_y);
  if (d_x == d_y) return set_error(MGCMT_ERR_ARG, "apply cannot run in place");
  CU(launch_band_apply(h->lev[level].dev, shift, d_x, d_y, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_band_smooth(mgcmt_band_t *h, int level, int smoother, int nu, double shift, double omega, double *d_v,
                      const double *d_f, void *stream) {
  int rc = level_ok(h, level);
  if (rc != MGCMT_OK) return rc;
  NEED_VEC(d_v, d_f);
  if (nu < 0) return set_error(MGCMT_ERR_ARG, "nu must be >= 0");
  return band_smooth(h, level, smoother, nu, shift, omega, d_v, d_f, (cudaStream_t)stream);
}

int mgcmt_band_residual_restrict(mgcmt_band_t *h, int level, double shift, const double *d_v, const double *d_f,
                                 double *d_rc, void *stream) {
  int rc = level_ok(h, level);
  if (rc != MGCMT_OK) return rc;
  if (level + 1 >= (int)h->lev.size()) return set_error(MGCMT_ERR_ARG, "no coarser level");
  NEED_VEC(d_v, d_f, d_rc);
  CU(launch_band_residual_restrict(h->lev[level].dev, shift, d_v, d_f, d_rc, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_band_prolong_correct(mgcmt_band_t *h, int level, const double *d_ec, double *d_v, void *stream) {
  int rc = level_ok(h, level);
  if (rc != MGCMT_OK) return rc;
  if (level + 1 >= (int)h->lev.size()) return set_error(MGCMT_ERR_ARG, "no coarser level");
  NEED_VEC(d_ec, d_v);
  CU(launch_band_prolong_correct(h->lev[level].dev.n, d_ec, d_v, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_band_coarse_solve(mgcmt_band_t *h, double shift, const double *d_f, double *d_v, void *stream) {
  if (!h) return set_error(MGCMT_ERR_ARG, "null banded hierarchy");
  NEED_VEC(d_f, d_v);
  if (d_f == d_v) return set_error(MGCMT_ERR_ARG, "coarse solve cannot run in place");
  const double *aug = nullptr;
  int rc = band_inverse(h, shift, (cudaStream_t)stream, &aug);
  if (rc != MGCMT_OK) return rc;
  CU(launch_band_gemv(h->lev.back().dev.n, aug, d_f, d_v, (cudaStream_t)stream));
  return MGCMT_OK;
}

int mgcmt_band_vcycle(mgcmt_band_t *h, double shift, int nu1, int nu2, int smoother, double omega, double *d_v,
                      const double *d_f, void *stream) {
  if (!h) return set_error(MGCMT_ERR_ARG, "null banded hierarchy");
  NEED_VEC(d_v, d_f);
  if (d_v == d_f) return set_error(MGCMT_ERR_ARG, "v and f must be different buffers");
  if (nu1 < 0 || nu2 < 0) return set_error(MGCMT_ERR_ARG, "nu1, nu2 must be >= 0");
  if (smoother != MGCMT_SMOOTH_WJACOBI && smoother != MGCMT_SMOOTH_GSLEX)
    return set_error(MGCMT_ERR_ARG, "banded operators take the wjacobi, gseidel and sor smoothers (red-black needs a radius-1 stencil)");
  return band_cycle(h, 0, shift, nu1, nu2, smoother, omega, d_v, d_f, (cudaStream_t)stream);
}

}  // extern "C"
