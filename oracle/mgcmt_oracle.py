"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy) of MultigridCMT's V-cycle path.

This file is the *oracle*: a from-scratch restatement of the arithmetic in the reference's
`MGCMTStencilMaker.py`, `MGCMTSolver.py` and `MGCMTProcessor.py` (file:line cited per function,
paths relative to /root/reference).  It exists so the CUDA path can be checked against something
that runs where the reference cannot (the GPU box has no /root/reference, and the reference is
Python 2).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  Nothing under `multigridcmt_b200/` imports it; the product path
raises if its CUDA library is missing instead of falling back to this file.

Parity status: PINNED.  `oracle/make_golden.py` executes the real reference (through the
syntax-only in-memory translation in `oracle/ref_loader.py`) in the build container and stores its
outputs on seeded inputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py` holds this file
to those outputs and to the known-answer scalars of the reference's `UnitTests/*.py`
(SURVEY.md section 4).  Un-pinned corners are listed in DESIGN.md ("Oracle").

Differences from the reference that are deliberate and arithmetic-neutral (<= a few ulp):
  * smoothers iterate `v <- v + w D^-1 (f - A v)` style sweeps / triangular substitutions instead
    of materialising `(D-L)^-1 U` (which is O(n^2) memory);
  * the grid hierarchy (R, P, R A P) can be cached per operator instead of rebuilt every call.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.linalg import eig as _dense_eig


def _is_pow2(x) -> bool:
    p = math.log(x) / math.log(2)
    return float(p).is_integer()


# ----------------------------------------------------------------------------------------------
# operators  (MGCMTStencilMaker.py)
# ----------------------------------------------------------------------------------------------
class StencilMaker:
    """Restates MGCMTStencilMaker (MGCMTStencilMaker.py:5-78)."""

    def laplacian(self, n, dimension="1d"):
        # MGCMTStencilMaker.py:15-25 -- n unknowns, h = 1/n, tridiag(1,-2,1)/h^2; 2D = kronsum.
        n = int(n)
        h = 1.0 / n
        one_d = sp.diags([np.ones(n - 1), -2.0 * np.ones(n), np.ones(n - 1)], [-1, 0, 1],
                         shape=(n, n), format="csc", dtype=float)
        one_d = one_d * (1 / h ** 2)
        if dimension == "1d":
            return one_d
        if dimension == "2d":
            return sp.kronsum(one_d, one_d, format="csc")
        return None

    def interpolation(self, old_gridsize, new_gridsize, dimension="1d"):
        # MGCMTStencilMaker.py:27-54.  Coarse point j sits on fine point m*(j+1)-1 (m = new/old);
        # column j holds the hat function [1..m..1]*(1/2)^p centred there, truncated at the ends.
        new_gridsize = int(new_gridsize)
        if dimension == "2d":
            s = self.interpolation(old_gridsize, new_gridsize, dimension="1d")
            if s is None:
                return None
            return sp.kron(s, s, format="csc")
        if dimension != "1d":
            return None
        p_old = math.log(old_gridsize) / math.log(2)
        p_new = math.log(new_gridsize) / math.log(2)
        if not p_new > p_old:
            print("New gridsize isn't bigger than old gridsize !")
            return None
        if not float(p_old).is_integer():
            print("Old gridsize isn't a power of 2 !")
            return None
        if not float(p_new).is_integer():
            print("New gridsize isn't a power of 2 !")
            return None
        m = int(new_gridsize / old_gridsize)
        n_coarse = len(range(m - 1, new_gridsize, m))
        prefactor = (1.0 / 2) ** (p_new - p_old)
        rows, cols, vals = [], [], []
        for j in range(n_coarse):
            centre = m * (j + 1) - 1
            for d in range(-(m - 1), m):
                i = centre + d
                if 0 <= i < new_gridsize:
                    rows.append(i)
                    cols.append(j)
                    vals.append(prefactor * float(m - abs(d)))
        return sp.csc_matrix((vals, (rows, cols)), shape=(new_gridsize, n_coarse), dtype=float)

    def restriction(self, old_gridsize, new_gridsize, dimension="1d"):
        # MGCMTStencilMaker.py:57-78.  1D: (1/2)^p P^T.  2D: fixed 1/4 (P (x) P)^T (quirk Q3).
        if dimension == "2d":
            p2 = self.interpolation(new_gridsize, old_gridsize, dimension="2d")
            if p2 is None:
                return None
            return (1.0 / 4.0) * p2.T
        if dimension != "1d":
            return None
        p_old = math.log(old_gridsize) / math.log(2)
        p_new = math.log(new_gridsize) / math.log(2)
        if not p_new < p_old:
            print("New gridsize is bigger (more elements) than old gridsize !")
            return None
        if not float(p_old).is_integer():
            print("Old gridsize isn't a power of 2 !")
            return None
        if not float(p_new).is_integer():
            print("New gridsize isn't a power of 2 !")
            return None
        prefactor = (1.0 / 2) ** (p_old - p_new)
        p1 = self.interpolation(new_gridsize, old_gridsize)
        return sp.csc_matrix(prefactor * p1.T)


# ----------------------------------------------------------------------------------------------
# Gram-Schmidt etc.  (MGCMTProcessor.py)
# ----------------------------------------------------------------------------------------------
class Processor:
    """Restates MGCMTProcessor (MGCMTProcessor.py:4-72)."""

    def projection(self, v, u):
        # MGCMTProcessor.py:10-20
        return (float(np.inner(v, u)) / float(np.inner(u, u))) * u

    def normalize(self, vectors):
        # MGCMTProcessor.py:52-63
        out = np.zeros((vectors.shape[0], vectors.shape[1]))
        for j in range(vectors.shape[1]):
            out[:, j] = vectors[:, j] / np.linalg.norm(vectors[:, j])
        return out

    def gramschmidt(self, vectors, modified=1):
        # MGCMTProcessor.py:22-50
        rows, cols = vectors.shape
        q = np.zeros((rows, cols))
        if not modified:
            for j in range(cols):
                col = vectors[:, j]
                q[:, j] = col
                for i in range(j):
                    q[:, j] = q[:, j] - self.projection(col, q[:, i])
            return self.normalize(q)
        w = np.zeros((rows, cols))
        for j in range(cols):
            w[:, j] = vectors[:, j]
        for i in range(cols):
            q[:, i] = w[:, i] / np.linalg.norm(w[:, i])
            for j in range(i + 1, cols):
                w[:, j] = w[:, j] - self.projection(w[:, j], q[:, i])
        return q

    def orthogonality_check(self, vectors):
        # MGCMTProcessor.py:65-72
        return np.array([[np.inner(vectors[:, i], vectors[:, j]) for j in range(vectors.shape[1])]
                         for i in range(vectors.shape[1])])


# ----------------------------------------------------------------------------------------------
# solver  (MGCMTSolver.py)
# ----------------------------------------------------------------------------------------------
def _col(x):
    # complex stays complex (the multiband Hamiltonians of ThesisProblem.py are complex128)
    x = np.asarray(x)
    return x.astype(complex if np.iscomplexobj(x) else float, copy=False).reshape(-1, 1)


class Solver:
    """Restates MGCMTSolver (MGCMTSolver.py:8-436)."""

    def __init__(self, cache_hierarchy=True):
        self.stencil_maker = StencilMaker()
        self.processor = Processor()
        self._cache = {} if cache_hierarchy else None

    # ---- smoothers -------------------------------------------------------------------------
    def wjacobi(self, v0, f, A, nu=4, omega=2. / 3.):
        # MGCMTSolver.py:182-208:  Rwj = I - w D^-1 A (entries a_ij/d_i), v <- Rwj v + w (f/D)
        A = sp.csc_matrix(A)
        n = A.shape[0]
        d = A.diagonal()
        dinv_a = A.tocoo()
        dinv_a = sp.csc_matrix((dinv_a.data / d[dinv_a.row], (dinv_a.row, dinv_a.col)), shape=A.shape)
        rwj = sp.eye(n, format="csc") - omega * dinv_a
        v = _col(v0)
        rhs = omega * (_col(f) / d.reshape(-1, 1))
        for _ in range(nu):
            v = rwj * v
            v = v + rhs
        return v

    def _split(self, A):
        A = sp.csr_matrix(A)
        n = A.shape[0]
        D = sp.diags(A.diagonal(), 0, shape=(n, n), format="csr")
        L = -sp.tril(A, -1, format="csr")
        U = -sp.triu(A, 1, format="csr")
        return D, L, U

    def gseidel(self, v0, f, A, nu=4):
        # MGCMTSolver.py:210-227:  v <- (D-L)^-1 U v + (D-L)^-1 f   (lexicographic forward sweep)
        D, L, U = self._split(A)
        low = sp.csr_matrix(D - L)
        v = _col(v0)
        f = _col(f)
        cf = spla.spsolve_triangular(low, f, lower=True)
        for _ in range(nu):
            v = spla.spsolve_triangular(low, U * v, lower=True)
            v = v + cf
        return v

    def sor(self, v0, f, A, nu=4, omega=1):
        # MGCMTSolver.py:229-246:  v <- (D-wL)^-1((1-w)D + wU) v + w (D-L)^-1 f   (quirk Q6)
        D, L, U = self._split(A)
        low_w = sp.csr_matrix(D - omega * L)
        low = sp.csr_matrix(D - L)
        rhs_m = sp.csr_matrix((1 - omega) * D + omega * U)
        v = _col(v0)
        f = _col(f)
        cf = omega * spla.spsolve_triangular(low, f, lower=True)
        for _ in range(nu):
            v = spla.spsolve_triangular(low_w, rhs_m * v, lower=True)
            v = v + cf
        return v

    def rbgs(self, v0, f, A, nu=4, omega=1.0, dimension="2d"):
        """Red-black Gauss-Seidel/SOR -- NOT in the reference (its gseidelrb is dead code,
        MGCMTSolver.py:248-279).  Defined here so the CUDA kernel has a CPU twin: four-colour
        (i%2, j%2) ordering in 2D [(0,0),(1,1),(0,1),(1,0)], two-colour in 1D [even, odd]; within a
        colour every unknown is relaxed with the latest values of all other colours.  For a 5-point
        stencil the four-colour order is the classical red-black sweep; it stays a valid
        Gauss-Seidel for the 9-point Galerkin stencils on coarse levels."""
        A = sp.csr_matrix(A)
        n = A.shape[0]
        d = A.diagonal()
        v = np.array(_col(v0)[:, 0])
        f = np.asarray(f, dtype=float).reshape(-1)
        if dimension == "2d":
            N = int(round(math.sqrt(n)))
            ii, jj = np.divmod(np.arange(n), N)
            colours = [(ii % 2 == a) & (jj % 2 == b) for (a, b) in ((0, 0), (1, 1), (0, 1), (1, 0))]
        else:
            idx = np.arange(n)
            colours = [idx % 2 == 0, idx % 2 == 1]
        subs = [(np.nonzero(c)[0], A[np.nonzero(c)[0], :]) for c in colours]
        for _ in range(nu):
            for rows, a_rows in subs:
                v[rows] = v[rows] + omega * (f[rows] - a_rows @ v) / d[rows]
        return v.reshape(-1, 1)

    # ---- grid hierarchy ----------------------------------------------------------------------
    def _transfer(self, A, g, dimension, stencil_maker):
        """R, P and the Galerkin product R A P for grid dimension g (MGCMTSolver.py:310-311,318)."""
        key = None
        if self._cache is not None:
            key = (id(A), A.shape, int(g), dimension, id(stencil_maker))
            hit = self._cache.get(key)
            if hit is not None and hit[0] is A:
                return hit[1:]
        R = stencil_maker.restriction(g, g / 2, dimension=dimension)
        P = stencil_maker.interpolation(g / 2, g, dimension=dimension)
        Ac = R * A * P
        if key is not None:
            self._cache[key] = (A, R, P, Ac)
        return R, P, Ac

    # ---- cycles ------------------------------------------------------------------------------
    def vcycle(self, v0, f, A, stencil_maker, nu1=4, nu2=4, smoother=None, shift=0, lowest_level=2,
               dimension="1d"):
        # MGCMTSolver.py:281-329
        if smoother is None:
            smoother = self.wjacobi
        n = len(v0)
        shifted = A - sp.eye(n) * shift  # shift kept apart from A: A is coarsened unshifted (:288)
        g = n if dimension == "1d" else np.sqrt(n)
        f = _col(f)
        v0 = _col(v0)
        if g < 2:
            print("Length of start vector is not a power of 2")
            return None
        if g == lowest_level:
            return _col(spla.spsolve(sp.csc_matrix(shifted), f))
        R, P, Ac = self._transfer(A, g, dimension, stencil_maker)
        v = smoother(v0, f, shifted, nu=nu1)
        r = R * (f - shifted * v)
        e2h = np.zeros(np.shape(r))
        # nu1/nu2 are NOT forwarded: coarse levels always run 4/4 (MGCMTSolver.py:320, quirk Q4)
        e2h = self.vcycle(e2h, r, Ac, stencil_maker, shift=shift, smoother=smoother,
                          lowest_level=lowest_level, dimension=dimension)
        e2h = _col(e2h)
        v = v + P * e2h
        v = smoother(v, f, shifted, nu=nu2)
        return v[:, 0]

    def twogrid(self, v0, f, A, stencil_maker, nu1=4, nu2=4, smoother=None, shift=0, dimension="1d"):
        # MGCMTSolver.py:331-371 (1-D only in effect: coarse shift matrix is eye(n//2), quirk Q10)
        if smoother is None:
            smoother = self.wjacobi
        n = len(v0)
        g = n if dimension == "1d" else np.sqrt(n)
        f = _col(f)
        v0 = _col(v0)
        shifted = A - sp.eye(n) * shift
        R, P, Ac = self._transfer(A, g, dimension, stencil_maker)
        coarse = Ac - sp.eye(n // 2) * shift
        v = smoother(v0, f, shifted, nu1)
        r = R * (f - shifted * v)
        e = _col(spla.spsolve(sp.csc_matrix(coarse), r))
        v = v + P * e
        v = smoother(v, f, shifted, nu2)
        return v[:, 0]

    def vcycle_matrix(self, v0_matrix, f_matrix, A, stencil_maker, nu1=4, nu2=4, smoother=None,
                      shifts=None, lowest_level=2, dimension="1d"):
        # MGCMTSolver.py:375-436; `shifts` must be 1-D of length k (the default None path is
        # broken in the reference on current scipy, SURVEY.md section 4 correction 3).
        n = v0_matrix.shape[0]
        k = f_matrix.shape[1]
        if smoother is None:
            smoother = self.wjacobi
        if shifts is None:
            shifts = np.zeros(k)
        shifts = np.asarray(shifts, dtype=float).reshape(-1)
        shifted = [A - sp.eye(n) * s for s in shifts]
        g = n if dimension == "1d" else np.sqrt(n)
        v = np.zeros((n, k))
        if g < 2:
            print("Length of start vector is not a power of 2")
            return None
        if g == lowest_level:
            for i in range(k):
                v[:, i] = spla.spsolve(sp.csc_matrix(shifted[i]), f_matrix[:, i])
            return v
        R, P, Ac = self._transfer(A, g, dimension, stencil_maker)
        r = np.zeros((R.shape[0], k))
        for i in range(k):
            v[:, i] = smoother(np.array(v0_matrix[:, i]), np.array(f_matrix[:, i]), shifted[i], nu=nu1)[:, 0]
        for i in range(k):
            r[:, i] = R * (f_matrix[:, i] - shifted[i] * v[:, i])
        e2h = self.vcycle_matrix(np.zeros(r.shape), r, Ac, stencil_maker, shifts=shifts,
                                 smoother=smoother, lowest_level=lowest_level, dimension=dimension)
        for i in range(k):
            v[:, i] = v[:, i] + P * e2h[:, i]
            v[:, i] = smoother(np.array(v[:, i]), np.array(f_matrix[:, i]), shifted[i], nu=nu2)[:, 0]
        return self.processor.gramschmidt(v)

    # ---- Rayleigh-quotient minimisation ------------------------------------------------------
    def rqmin(self, A, v0, M=None, nu=4):
        # MGCMTSolver.py:17-57.  (M=None is unusable in the reference, quirk Q9; here it means I.)
        x = np.array(v0, dtype=float)
        if M is None:
            M = sp.eye(len(x), format="csr")
        rho = np.dot(x, A.dot(x)) / np.dot(x, M.dot(x))
        gold = np.array(x)
        g = 2 * (A.dot(x) - rho * M.dot(x))
        p = np.array(x)
        Rm = np.zeros((2, 2))
        RM = np.zeros((2, 2))
        for it in range(nu):
            if it == 0:
                p = -g
            else:
                p = -g + (np.dot(g, M.dot(g)) / np.dot(gold, M.dot(gold))) * p
            Rm[0, 0] = np.dot(x, A.dot(x)); Rm[0, 1] = np.dot(x, A.dot(p))
            Rm[1, 0] = np.dot(p, A.dot(x)); Rm[1, 1] = np.dot(p, A.dot(p))
            RM[0, 0] = np.dot(x, M.dot(x)); RM[0, 1] = np.dot(x, M.dot(p))
            RM[1, 0] = np.dot(p, M.dot(x)); RM[1, 1] = np.dot(p, M.dot(p))
            w, vecs = _dense_eig(Rm, b=RM)
            rx = np.array(vecs[:, np.argmin(w)])
            delta = rx[1] / rx[0]
            x = np.array(x + delta * p)
            rho = np.dot(x, A.dot(x)) / np.dot(x, M.dot(x))
            gold = np.array(g)
            g = np.array(2 * (A.dot(x) - rho * M.dot(x)))
        return x, rho

    def vcycle_rqmg(self, x, A, M, nu1=4, nu2=4, nmin=2, dimension="1d"):
        # MGCMTSolver.py:99-122 (1-D transfer operators only, D6).  dimension="2d" is NOT in the reference: the same
        # recursion with the 2-D transfer operators (CPU twin of the product's extension; parity unpinned for 2-D).
        k = np.array(x)
        n = len(k)
        k, rho = self.rqmin(A, k, M, nu=nu1)
        g = n if dimension == "1d" else int(round(math.sqrt(n)))
        if n > nmin and g > 2:
            P = self.stencil_maker.interpolation(g // 2, g, dimension=dimension)
            R = self.stencil_maker.restriction(g, g // 2, dimension=dimension)
            Ac = R * A * P
            Mc = R * M * P
            c, rho = self.vcycle_rqmg(R * k, Ac, Mc, nu1=nu1, nu2=nu2, nmin=nmin, dimension=dimension)
            k = k + P * c
            k, rho = self.rqmin(A, k, M, nu=nu2)
        return k, rho

    def vcycle_rqmg2(self, x_matrix, A, M, nu1=4, nu2=4, nmin=2, level=0):
        # MGCMTSolver.py:59-94
        k = np.array(x_matrix)
        n, nv = k.shape
        for i in range(nv):
            k[:, i], rho = self.rqmin(A, k[:, i], M, nu=nu1)
        if level == 0:
            for _ in range(4):
                k = self.processor.gramschmidt(k)
        if n > nmin:
            P = self.stencil_maker.interpolation(n // 2, n)
            R = self.stencil_maker.restriction(n, n // 2)
            Ac = R * A * P
            Mc = R * M * P
            kc = np.zeros((n // 2, nv))
            for i in range(nv):
                kc[:, i] = R * k[:, i]
            c = self.vcycle_rqmg2(kc, Ac, Mc, nu1=nu1, nu2=nu2, nmin=nmin, level=level + 1)
            for i in range(nv):
                k[:, i] = k[:, i] + P * c[:, i]
                k[:, i], rho = self.rqmin(A, k[:, i], M, nu=nu2)
        return k


# ----------------------------------------------------------------------------------------------
# closed-form spectrum of the reference's operator ("extra oracles")
# ----------------------------------------------------------------------------------------------
def well_eigenvalue_1d(n, k):
    """k-th eigenvalue (k >= 1) of (-1/pi^2) * laplacian(n, '1d')."""
    return (4.0 * n * n / math.pi ** 2) * math.sin(k * math.pi / (2.0 * (n + 1))) ** 2


def well_eigenvector_1d(n, k):
    v = np.sin(k * math.pi * (np.arange(n) + 1.0) / (n + 1.0))
    return v / np.linalg.norm(v)


def well_eigenvalue_2d(n, kx, ky):
    return well_eigenvalue_1d(n, kx) + well_eigenvalue_1d(n, ky)


def well_eigenvector_2d(n, kx, ky):
    """Row-major (index i*n + j) eigenvector v_kx (x) v_ky of (-1/pi^2) * laplacian(n, '2d')."""
    return np.kron(well_eigenvector_1d(n, kx), well_eigenvector_1d(n, ky))
