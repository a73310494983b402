/*
 * mgcmt_oracle.c -- TEST INFRASTRUCTURE ONLY.  Matrix-free C / OpenMP restatement of the 2-D V-cycle of
 * MultigridCMT (MGCMTSolver.vcycle, MGCMTSolver.py:281-329, with the weighted-Jacobi smoother :182-208, the
 * transfer operators of MGCMTStencilMaker.py:27-78 and Galerkin coarse operators :318) for separable operators
 * A = I (x) Kb + Ka (x) I, i.e. the wells of the reference's 2-D drivers.
 *
 * Why it exists: the numpy/scipy oracle (oracle/mgcmt_oracle.py, pinned to the real reference through
 * tests/golden) is single-threaded and matrix-based, so as the *timed CPU baseline* it undersells the host.  This file
 * is the same arithmetic without matrices, threaded over grid rows, so bench.py's cpu_baseline / --impl reference leg
 * can use all host cores.  tests/test_c_oracle.py holds it to the numpy oracle (1e-12).  Nothing under
 * multigridcmt_b200/ links or loads it.
 *
 * Level operators are kept as the tridiagonal factors of  A_l = Ma (x) Kb + Ka (x) Mb  (M_0 = I), coarsened with
 * T -> R T P  (R = [1/4 1/2 1/4] with a truncated last row, P = 2 R^T; coarse j sits on fine 2j+1).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int n;                                   /* n x n grid */
  double *ka[3], *ma[3], *kb[3], *mb[3];   /* lo, di, up; length n */
  double *v, *f, *t;                       /* work vectors (levels >= 1: v, f; all: t) */
  double *lu;                              /* coarsest: banded LU of (A - shift I), n^2 rows x lu_w window */
  int *piv;
  int lu_kl, lu_ku, lu_w;
  double lu_shift;
  int lu_valid;
} level_t;

typedef struct {
  int nlev;
  level_t *lev;
} hier_t;

static void galerkin(int nf, double *const t[3], double *c[3]) {
  const int nc = nf / 2;
  const double wr[3] = {0.25, 0.5, 0.25}, wp[3] = {0.5, 1.0, 0.5};
  for (int j = 0; j < nc; ++j) {
    double res[3];
    for (int dj = -1; dj <= 1; ++dj) {
      const int jp = j + dj;
      double acc = 0.0;
      if (jp >= 0 && jp < nc) {
        for (int s = 0; s < 3; ++s) {
          const int b = 2 * jp + s;
          if (b >= nf) continue;
          double rt = 0.0;
          for (int q = 0; q < 3; ++q) {
            const int a = 2 * j + q;
            if (a >= nf) continue;
            double tv;
            if (b == a - 1) tv = t[0][a];
            else if (b == a) tv = t[1][a];
            else if (b == a + 1) tv = t[2][a];
            else continue;
            rt += wr[q] * tv;
          }
          acc += rt * wp[s];
        }
      }
      res[dj + 1] = acc;
    }
    c[0][j] = j > 0 ? res[0] : 0.0;
    c[1][j] = res[1];
    c[2][j] = j + 1 < nc ? res[2] : 0.0;
  }
}

static double *dalloc(size_t n) { return (double *)calloc(n, sizeof(double)); }

hier_t *orc_create(int n, const double *row_lo, const double *row_di, const double *row_up, const double *col_lo,
                   const double *col_di, const double *col_up, int lowest) {
  hier_t *h = (hier_t *)calloc(1, sizeof(hier_t));
  int nlev = 1;
  for (int c = n; c > lowest; c >>= 1) ++nlev;
  h->nlev = nlev;
  h->lev = (level_t *)calloc(nlev, sizeof(level_t));
  int m = n;
  for (int l = 0; l < nlev; ++l) {
    level_t *L = &h->lev[l];
    L->n = m;
    for (int k = 0; k < 3; ++k) { L->ka[k] = dalloc(m); L->ma[k] = dalloc(m); L->kb[k] = dalloc(m); L->mb[k] = dalloc(m); }
    L->t = dalloc((size_t)m * m);
    if (l > 0) { L->v = dalloc((size_t)m * m); L->f = dalloc((size_t)m * m); }
    if (l == 0) {
      for (int i = 0; i < m; ++i) {
        L->ka[0][i] = i > 0 ? row_lo[i] : 0.0; L->ka[1][i] = row_di[i]; L->ka[2][i] = i + 1 < m ? row_up[i] : 0.0;
        L->kb[0][i] = i > 0 ? col_lo[i] : 0.0; L->kb[1][i] = col_di[i]; L->kb[2][i] = i + 1 < m ? col_up[i] : 0.0;
        L->ma[1][i] = 1.0; L->mb[1][i] = 1.0;
      }
    } else {
      level_t *F = &h->lev[l - 1];
      galerkin(F->n, F->ka, L->ka); galerkin(F->n, F->ma, L->ma);
      galerkin(F->n, F->kb, L->kb); galerkin(F->n, F->mb, L->mb);
    }
    m >>= 1;
  }
  return h;
}

void orc_destroy(hier_t *h) {
  if (!h) return;
  for (int l = 0; l < h->nlev; ++l) {
    level_t *L = &h->lev[l];
    for (int k = 0; k < 3; ++k) { free(L->ka[k]); free(L->ma[k]); free(L->kb[k]); free(L->mb[k]); }
    free(L->t); free(L->v); free(L->f); free(L->lu); free(L->piv);
  }
  free(h->lev);
  free(h);
}

/* (A_s x)(i,j) */
static inline double apply_pt(const level_t *L, double shift, const double *x, int i, int j) {
  const int n = L->n;
#define X(ii, jj) (((ii) >= 0 && (ii) < n && (jj) >= 0 && (jj) < n) ? x[(size_t)(ii) * n + (jj)] : 0.0)
  double acc = 0.0;
  for (int di = -1; di <= 1; ++di) {
    const double ma = L->ma[di + 1][i], ka = L->ka[di + 1][i];
    if (ma == 0.0 && ka == 0.0) continue;
    const double xm = X(i + di, j - 1), x0 = X(i + di, j), xp = X(i + di, j + 1);
    const double t = L->kb[0][j] * xm + L->kb[1][j] * x0 + L->kb[2][j] * xp;
    const double s = L->mb[0][j] * xm + L->mb[1][j] * x0 + L->mb[2][j] * xp;
    acc += ma * t + ka * s;
  }
#undef X
  return acc - shift * x[(size_t)i * n + j];
}

/* out[j] = (A_s x)(i, j) for a whole row; interior columns in a branch-free (vectorisable) loop */
static void apply_row(const level_t *L, double shift, const double *x, int i, const double *zero, double *out) {
  const int n = L->n;
  const double *xm = i > 0 ? x + (size_t)(i - 1) * n : zero, *x0 = x + (size_t)i * n,
               *xp = i + 1 < n ? x + (size_t)(i + 1) * n : zero;
  const double mal = L->ma[0][i], mad = L->ma[1][i], mau = L->ma[2][i];
  const double kal = L->ka[0][i], kad = L->ka[1][i], kau = L->ka[2][i];
  const double *kbl = L->kb[0], *kbd = L->kb[1], *kbu = L->kb[2], *mbl = L->mb[0], *mbd = L->mb[1], *mbu = L->mb[2];
  out[0] = apply_pt(L, shift, x, i, 0);
  if (n > 1) out[n - 1] = apply_pt(L, shift, x, i, n - 1);
  for (int j = 1; j < n - 1; ++j) {
    const double tm = kbl[j] * xm[j - 1] + kbd[j] * xm[j] + kbu[j] * xm[j + 1];
    const double t0 = kbl[j] * x0[j - 1] + kbd[j] * x0[j] + kbu[j] * x0[j + 1];
    const double tp = kbl[j] * xp[j - 1] + kbd[j] * xp[j] + kbu[j] * xp[j + 1];
    const double sm = mbl[j] * xm[j - 1] + mbd[j] * xm[j] + mbu[j] * xm[j + 1];
    const double s0 = mbl[j] * x0[j - 1] + mbd[j] * x0[j] + mbu[j] * x0[j + 1];
    const double sp = mbl[j] * xp[j - 1] + mbd[j] * xp[j] + mbu[j] * xp[j + 1];
    out[j] = (mal * tm + kal * sm) + (mad * t0 + kad * s0) + (mau * tp + kau * sp) - shift * x0[j];
  }
}

static void jacobi(const level_t *L, double shift, double omega, int nu, double *v, const double *f, double *tmp) {
  const int n = L->n;
  double *a = v, *b = tmp;
  double *zero = dalloc(n);
  for (int it = 0; it < nu; ++it) {
#pragma omp parallel
    {
      double *av = (double *)malloc(sizeof(double) * n);
#pragma omp for schedule(static)
      for (int i = 0; i < n; ++i) {
        apply_row(L, shift, a, i, zero, av);
        for (int j = 0; j < n; ++j) {
          const double d = (L->ma[1][i] * L->kb[1][j] + L->ka[1][i] * L->mb[1][j]) - shift;
          b[(size_t)i * n + j] = a[(size_t)i * n + j] + omega * (f[(size_t)i * n + j] - av[j]) / d;
        }
      }
      free(av);
    }
    double *s = a; a = b; b = s;
  }
  free(zero);
  if (a != v) memcpy(v, a, sizeof(double) * (size_t)n * n);
}

/* Red-black (four-colour) Gauss-Seidel / SOR, the CPU twin of the product's rbgs (NOT in the reference: its gseidelrb is
 * dead code, MGCMTSolver.py:248-279; the ordering is the one oracle/mgcmt_oracle.py: Solver.rbgs defines): colours
 * (i%2, j%2) in the order (0,0), (1,1), (0,1), (1,0); within a colour every unknown is relaxed with the latest values of
 * the other colours -- points of one colour never neighbour each other, so a colour is one parallel pass, in place. */
static void rbgs(const level_t *L, double shift, double omega, int nu, double *v, const double *f) {
  const int n = L->n;
  static const int col[4][2] = {{0, 0}, {1, 1}, {0, 1}, {1, 0}};
  for (int it = 0; it < nu; ++it)
    for (int c = 0; c < 4; ++c) {
      const int pa = col[c][0], pb = col[c][1];
#pragma omp parallel for schedule(static)
      for (int i = pa; i < n; i += 2)
        for (int j = pb; j < n; j += 2) {
          const double d = (L->ma[1][i] * L->kb[1][j] + L->ka[1][i] * L->mb[1][j]) - shift;
          v[(size_t)i * n + j] += omega * (f[(size_t)i * n + j] - apply_pt(L, shift, v, i, j)) / d;
        }
    }
}

static void residual_restrict(const level_t *L, double shift, const double *v, const double *f, double *r, double *rc) {
  const int n = L->n, nc = n / 2;
  double *zero = dalloc(n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    double *ri = r + (size_t)i * n;
    apply_row(L, shift, v, i, zero, ri);
    for (int j = 0; j < n; ++j) ri[j] = f[(size_t)i * n + j] - ri[j];
  }
  free(zero);
  const double w[3] = {0.25, 0.5, 0.25};
#pragma omp parallel for schedule(static)
  for (int I = 0; I < nc; ++I)
    for (int J = 0; J < nc; ++J) {
      double acc = 0.0;
      for (int a = 0; a < 3; ++a) {
        const int i = 2 * I + a;
        if (i >= n) continue;
        double ra = 0.0;
        for (int b = 0; b < 3; ++b) {
          const int j = 2 * J + b;
          if (j >= n) continue;
          ra += w[b] * r[(size_t)i * n + j];
        }
        acc += w[a] * ra;
      }
      rc[(size_t)I * nc + J] = acc;
    }
}

static void prolong_add(const level_t *L, const double *e, double *v) {
  const int n = L->n, nc = n / 2;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const int I = i >> 1, J = j >> 1;
#define E(ii, jj) (((ii) >= 0 && (ii) < nc && (jj) >= 0 && (jj) < nc) ? e[(size_t)(ii) * nc + (jj)] : 0.0)
      double pe;
      if (i & 1) {
        pe = (j & 1) ? E(I, J) : 0.5 * (E(I, J - 1) + E(I, J));
      } else {
        const double top = (j & 1) ? E(I - 1, J) : 0.5 * (E(I - 1, J - 1) + E(I - 1, J));
        const double bot = (j & 1) ? E(I, J) : 0.5 * (E(I, J - 1) + E(I, J));
        pe = 0.5 * (top + bot);
      }
#undef E
      v[(size_t)i * n + j] += pe;
    }
}

/* Coarsest level: exact solve (`spsolve(shifted_matrix, f)`, MGCMTSolver.py:305-308) by LU with partial pivoting.  The
 * 9-point operator on an n x n grid has half-bandwidth n + 1, so the factorisation works on the band only (window
 * storage: row i keeps columns i - kl .. i + ku + kl, the extra kl columns take the fill of the row interchanges) --
 * O(N kl (kl + ku)) work instead of O(N^3), which is what makes lowest_level = 64 (BASELINE config 3: "7 levels" at
 * 4096^2, N = 4096 unknowns) usable.  Multipliers are not permuted retroactively, so the solve interleaves the
 * interchanges with the eliminations (the LAPACK gbtrf/gbtrs convention). */
#define BW(L, i, c) (L)->lu[(size_t)(i) * (L)->lu_w + ((c) - ((i) - (L)->lu_kl))]
static int coarse_factor(level_t *L, double shift) {
  const int n = L->n, N = n * n;
  const int kl = (n + 1 < N - 1) ? n + 1 : N - 1, ku = kl;
  L->lu_kl = kl;
  L->lu_ku = ku;
  L->lu_w = 2 * kl + ku + 1;
  if (!L->lu) { L->lu = dalloc((size_t)N * L->lu_w); L->piv = (int *)calloc(N, sizeof(int)); }
  memset(L->lu, 0, sizeof(double) * (size_t)N * L->lu_w);
  double *unit = dalloc(N);
  for (int c = 0; c < N; ++c) {      /* column c of A_s = A_s e_c, rows within the band */
    unit[c] = 1.0;
    const int rlo = c - ku < 0 ? 0 : c - ku, rhi = c + kl > N - 1 ? N - 1 : c + kl;
    for (int r = rlo; r <= rhi; ++r) BW(L, r, c) = apply_pt(L, shift, unit, r / n, r % n);
    unit[c] = 0.0;
  }
  free(unit);
  for (int k = 0; k < N; ++k) {
    const int rhi = k + kl > N - 1 ? N - 1 : k + kl;
    const int chi = k + ku + kl > N - 1 ? N - 1 : k + ku + kl;
    int p = k;
    double best = fabs(BW(L, k, k));
    for (int r = k + 1; r <= rhi; ++r)
      if (fabs(BW(L, r, k)) > best) { best = fabs(BW(L, r, k)); p = r; }
    L->piv[k] = p;
    if (best == 0.0) return 1;
    if (p != k)
      for (int c = k; c <= chi; ++c) { double t = BW(L, k, c); BW(L, k, c) = BW(L, p, c); BW(L, p, c) = t; }
    const double piv = BW(L, k, k);
    for (int r = k + 1; r <= rhi; ++r) {
      const double m = BW(L, r, k) / piv;
      BW(L, r, k) = m;
      if (m != 0.0)
        for (int c = k + 1; c <= chi; ++c) BW(L, r, c) -= m * BW(L, k, c);
    }
  }
  L->lu_shift = shift;
  L->lu_valid = 1;
  return 0;
}

static void coarse_solve(level_t *L, const double *f, double *v) {
  const int N = L->n * L->n, kl = L->lu_kl, ku = L->lu_ku;
  memcpy(v, f, sizeof(double) * N);
  for (int k = 0; k < N; ++k) {
    const int p = L->piv[k];
    if (p != k) { double t = v[k]; v[k] = v[p]; v[p] = t; }
    const int rhi = k + kl > N - 1 ? N - 1 : k + kl;
    for (int r = k + 1; r <= rhi; ++r) v[r] -= BW(L, r, k) * v[k];
  }
  for (int k = N - 1; k >= 0; --k) {
    const int chi = k + ku + kl > N - 1 ? N - 1 : k + ku + kl;
    double t = v[k];
    for (int c = k + 1; c <= chi; ++c) t -= BW(L, k, c) * v[c];
    v[k] = t / BW(L, k, k);
  }
}
#undef BW

static void smooth(const level_t *L, int smoother, double shift, double omega, int nu, double *v, const double *f, double *tmp) {
  if (smoother == 1) rbgs(L, shift, omega, nu, v, f);
  else jacobi(L, shift, omega, nu, v, f, tmp);
}

static int cycle(hier_t *h, int l, int smoother, double shift, double omega, int nu1, int nu2, double *v, const double *f) {
  level_t *L = &h->lev[l];
  if (l == h->nlev - 1) {
    if (!L->lu_valid || L->lu_shift != shift)
      if (coarse_factor(L, shift)) return 1;
    coarse_solve(L, f, v);
    return 0;
  }
  level_t *C = &h->lev[l + 1];
  smooth(L, smoother, shift, omega, nu1, v, f, L->t);
  residual_restrict(L, shift, v, f, L->t, C->f);
  memset(C->v, 0, sizeof(double) * (size_t)C->n * C->n);
  if (cycle(h, l + 1, smoother, shift, omega, 4, 4, C->v, C->f)) return 1;   /* coarse levels always 4/4 (MGCMTSolver.py:320) */
  prolong_add(L, C->v, v);
  smooth(L, smoother, shift, omega, nu2, v, f, L->t);
  return 0;
}

int orc_vcycle(hier_t *h, double shift, double omega, int nu1, int nu2, double *v, const double *f) {
  return cycle(h, 0, 0, shift, omega, nu1, nu2, v, f);
}

/* smoother: 0 = weighted Jacobi (MGCMTSolver.py:182-208), 1 = red-black Gauss-Seidel / SOR (see rbgs above) */
int orc_vcycle_smoother(hier_t *h, int smoother, double shift, double omega, int nu1, int nu2, double *v, const double *f) {
  return cycle(h, 0, smoother, shift, omega, nu1, nu2, v, f);
}

/* ---- one outer iteration of the shift method on a block of k vectors (2DPotGS.py:91-105), all host threads ----------
 * for each c:  w_c = vcycle(0, v_c, H, shift = mu_c)            (2DPotGS.py:95; hs[c]: the hierarchy that caches mu_c's LU)
 *              lam[2c] = w_c^T H w_c, lam[2c+1] = w_c^T w_c     (the Rayleigh quotient of :103 on the un-normalised w)
 * then         W = gramschmidt(W)                               (MGCMTProcessor.py:44-50, modified: normalise u_i, subtract
 *                                                                its projection from all later columns)
 * V, W: k vectors of n*n doubles, vector-major.  Used by bench.py's reference arm so that it times the same step as
 * the GPU arm (cycles + Rayleigh sums + block orthonormalisation), not the cycles alone. */
static double dot_omp(const double *x, const double *y, size_t n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long long i = 0; i < (long long)n; ++i) s += x[i] * y[i];
  return s;
}

int orc_block_step(hier_t **hs, int k, int smoother, const double *shifts, double omega, int nu1, int nu2, const double *V,
                   double *W, double *lam) {
  const level_t *L0 = &hs[0]->lev[0];
  const int n = L0->n;
  const size_t nn = (size_t)n * n;
  for (int c = 0; c < k; ++c) {
    double *w = W + (size_t)c * nn;
    memset(w, 0, sizeof(double) * nn);
    if (cycle(hs[c], 0, smoother, shifts[c], omega, nu1, nu2, w, V + (size_t)c * nn)) return 1;
    double num = 0.0, den = 0.0;
    double *zero = dalloc(n);
#pragma omp parallel
    {
      double *row = (double *)malloc(sizeof(double) * n);
#pragma omp for reduction(+ : num, den) schedule(static)
      for (int i = 0; i < n; ++i) {
        apply_row(L0, 0.0, w, i, zero, row);
        const double *wi = w + (size_t)i * n;
        for (int j = 0; j < n; ++j) { num += wi[j] * row[j]; den += wi[j] * wi[j]; }
      }
      free(row);
    }
    free(zero);
    lam[2 * c] = num;
    lam[2 * c + 1] = den;
  }
  for (int i = 0; i < k; ++i) {
    double *ui = W + (size_t)i * nn;
    const double nrm = sqrt(dot_omp(ui, ui, nn));
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < (long long)nn; ++t) ui[t] /= nrm;
    if (i + 1 == k) break;
    const double uu = dot_omp(ui, ui, nn);
    for (int j = i + 1; j < k; ++j) {
      double *wj = W + (size_t)j * nn;
      const double a = dot_omp(wj, ui, nn) / uu;
#pragma omp parallel for schedule(static)
      for (long long t = 0; t < (long long)nn; ++t) wj[t] -= a * ui[t];
    }
  }
  return 0;
}
