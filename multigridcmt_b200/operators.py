"""Host-side operator recognition: scipy.sparse matrix  ->  separable tridiagonal factors.

The reference hands `vcycle` an explicit scipy.sparse matrix (e.g. 2DPotGS.py:26-27:
`hamiltonian = (-1/pi**2) * stencil_maker.laplacian(N, "2d")`).  The device path keeps operators
in the separable form  A = I (x) Kb + Ka (x) I  (row-major index i*N + j, Ka acts on i, Kb on j;
include/mgcmt_b200.h), so the first thing the drop-in does is check that the matrix it was given IS
of that form and pull out the two tridiagonals.  Anything else is refused loudly -- there is no
general-sparse or CPU path behind this one.

`SeparableOperator` can also be built directly (matrix-free) for grids where materialising the
scipy matrix is the bottleneck (4096^2: 84 M non-zeros; 16384^2: 1.3 G).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class UnsupportedOperator(ValueError):
    """The matrix is not a separable radius-1 stencil this path can run."""


def _tri(n, lo, di, up):
    return sp.diags([np.asarray(lo)[1:], np.asarray(di), np.asarray(up)[:-1]], [-1, 0, 1], shape=(n, n),
                    format="csc", dtype=float)


class SeparableOperator:
    """A = I (x) Kb + Ka (x) I on an nrows x ncols grid; 1-D when nrows == 1 (then A = Kb + ka_di I).

    row = (lo, di, up) of Ka (length nrows), col = (lo, di, up) of Kb (length ncols); lo[0] and up[-1]
    are ignored.  Mimics the slice of the scipy.sparse surface the reference's drivers touch: `.shape`,
    scalar `*`, unary `-`, `.dot(v)`, `.diagonal()`, `.tocsc()`.
    """

    def __init__(self, nrows, ncols, row, col, dimension):
        self.nrows = int(nrows)
        self.ncols = int(ncols)
        self.row = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in row)
        self.col = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in col)
        self.dimension = dimension
        n = self.nrows * self.ncols
        self.shape = (n, n)
        self._device = {}   # device -> (hierarchy cache); filled by hierarchy.get_hierarchy

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def laplacian(cls, n, dimension="1d", scale=1.0):
        """scale * MGCMTStencilMaker.laplacian(n, dimension) (MGCMTStencilMaker.py:15-25), matrix-free."""
        n = int(n)
        h = 1.0 / n
        c = 1.0 / h ** 2
        lo = np.full(n, 1.0 * c) * scale
        di = np.full(n, -2.0 * c) * scale
        up = np.full(n, 1.0 * c) * scale
        if dimension == "1d":
            z = np.zeros(1)
            return cls(1, n, (z, z, z), (lo, di, up), "1d")
        return cls(n, n, (lo.copy(), di.copy(), up.copy()), (lo, di, up), "2d")

    @classmethod
    def from_sparse(cls, A, dimension):
        """Recognise a scipy.sparse matrix; raises UnsupportedOperator if it is not separable."""
        if isinstance(A, SeparableOperator):
            return A
        if not sp.issparse(A):
            A = sp.csc_matrix(np.asarray(A))
        if np.iscomplexobj(A.data if hasattr(A, "data") else A):
            raise UnsupportedOperator("complex operators are outside this path (SURVEY.md section 8(f) row 3)")
        n = A.shape[0]
        if A.shape[0] != A.shape[1]:
            raise UnsupportedOperator("operator must be square")
        A = A.tocsc() if A.format not in ("csc", "csr", "dia") else A
        nnz = A.count_nonzero()
        if dimension == "1d":
            lo = np.zeros(n); up = np.zeros(n)
            di = np.asarray(A.diagonal(0), dtype=float)
            if n > 1:
                lo[1:] = A.diagonal(-1)
                up[:-1] = A.diagonal(1)
            if np.count_nonzero(lo) + np.count_nonzero(di) + np.count_nonzero(up) != nnz:
                raise UnsupportedOperator("1-D operator is not tridiagonal")
            z = np.zeros(1)
            return cls(1, n, (z, z, z), (lo, di, up), "1d")
        if dimension != "2d":
            raise UnsupportedOperator("dimension must be '1d' or '2d'")
        N = int(round(np.sqrt(n)))
        if N * N != n:
            raise UnsupportedOperator("2-D operator size is not a square number")
        d0 = np.asarray(A.diagonal(0), dtype=float).reshape(N, N)
        dE = np.zeros(n); dW = np.zeros(n); dS = np.zeros(n); dNn = np.zeros(n)
        dE[:-1] = A.diagonal(1)       # (i,j) -> (i,j+1)
        dW[1:] = A.diagonal(-1)       # (i,j) -> (i,j-1)
        dS[:-N] = A.diagonal(N)       # (i,j) -> (i+1,j)
        dNn[N:] = A.diagonal(-N)      # (i,j) -> (i-1,j)
        if (np.count_nonzero(d0) + np.count_nonzero(dE) + np.count_nonzero(dW) + np.count_nonzero(dS)
                + np.count_nonzero(dNn)) != nnz:
            raise UnsupportedOperator("2-D operator has entries outside the 5-point stencil")
        dE = dE.reshape(N, N); dW = dW.reshape(N, N); dS = dS.reshape(N, N); dNn = dNn.reshape(N, N)
        if np.any(dE[:, -1] != 0) or np.any(dW[:, 0] != 0):
            raise UnsupportedOperator("2-D operator couples the end of one grid row to the next")
        col_up = dE[0].copy(); col_lo = dW[0].copy()
        row_up = dS[:, 0].copy(); row_lo = dNn[:, 0].copy()
        if not (np.array_equal(dE, np.broadcast_to(col_up, (N, N))) and np.array_equal(dW, np.broadcast_to(col_lo, (N, N)))
                and np.array_equal(dS, np.broadcast_to(row_up[:, None], (N, N)))
                and np.array_equal(dNn, np.broadcast_to(row_lo[:, None], (N, N)))):
            raise UnsupportedOperator("off-diagonal stencil coefficients are not separable")
        # diagonal must split as a[i] + b[j]
        a = d0[:, 0] - 0.5 * d0[0, 0]
        b = d0[0, :] - 0.5 * d0[0, 0]
        recon = a[:, None] + b[None, :]
        tol = 4 * np.finfo(float).eps * max(1.0, float(np.max(np.abs(d0))))
        if np.max(np.abs(recon - d0)) > tol:
            raise UnsupportedOperator("diagonal is not of the form a[i] + b[j] (non-separable potential)")
        return cls(N, N, (row_lo, a, row_up), (col_lo, b, col_up), "2d")

    # ---- scipy-like surface -----------------------------------------------------------------
    def _scaled(self, s):
        s = float(s)
        return SeparableOperator(self.nrows, self.ncols, tuple(a * s for a in self.row),
                                 tuple(a * s for a in self.col), self.dimension)

    def __mul__(self, other):
        if np.isscalar(other):
            return self._scaled(other)
        return self.dot(other)

    def __rmul__(self, other):
        if np.isscalar(other):
            return self._scaled(other)
        return NotImplemented

    def __neg__(self):
        return self._scaled(-1.0)

    def __truediv__(self, other):
        return self._scaled(1.0 / float(other))

    def diagonal(self):
        return (self.row[1][:, None] + self.col[1][None, :]).reshape(-1)

    def tocsc(self):
        """Materialise as scipy CSC (small grids only)."""
        kb = _tri(self.ncols, *self.col)
        if self.nrows == 1:
            return sp.csc_matrix(kb + sp.eye(self.ncols) * float(self.row[1][0]))
        ka = _tri(self.nrows, *self.row)
        return sp.csc_matrix(sp.kron(sp.eye(self.nrows), kb) + sp.kron(ka, sp.eye(self.ncols)))

    def dot(self, x):
        """A x on the GPU.  numpy in -> numpy out; torch cuda tensor in -> torch cuda tensor out."""
        from . import hierarchy
        return hierarchy.apply_operator(self, x)

    __matmul__ = dot

    # operator fingerprint used for caching recognised scipy matrices
    def key(self):
        return (self.nrows, self.ncols, self.dimension)


_RECOGNISED = {}   # id(A) -> (weakref to A, fingerprint, operator)


def data_fingerprint(A):
    """Position-sensitive fingerprint of a scipy matrix: CRC-32 of its value / index / pointer arrays -- whole arrays up
    to 64 K entries, 64 evenly spaced blocks of 1024 entries (plus the ends) beyond that, so that the check stays well
    under a millisecond on the 84 M-entry matrices the drivers build at 4096^2.  An in-place edit that keeps sums
    (a potential moved along the diagonal with `setdiag`) changes the CRC; one confined to entries no block samples can
    slip through on very large matrices -- call `invalidate(A)` after editing a matrix in place."""
    import zlib
    crc = 0
    for name in ("data", "indices", "indptr", "offsets"):
        arr = getattr(A, name, None)
        if not isinstance(arr, np.ndarray) or arr.size == 0:
            continue
        flat = np.ascontiguousarray(arr).reshape(-1)
        if flat.size <= (1 << 16):
            crc = zlib.crc32(flat.view(np.uint8), crc)
        else:
            step = flat.size // 64
            for b in range(64):
                crc = zlib.crc32(flat[b * step:b * step + 1024].view(np.uint8), crc)
            crc = zlib.crc32(flat[-1024:].view(np.uint8), crc)
        crc = zlib.crc32(np.int64(flat.size).tobytes(), crc)
    data = getattr(A, "data", None)
    addr = data.ctypes.data if isinstance(data, np.ndarray) else None
    return (addr, crc)


def _cache_lookup(cache, A, fp):
    hit = cache.get(id(A))
    if hit is not None and hit[0]() is A and hit[1] == fp:
        return hit[2]
    return None


def _cache_store(cache, A, fp, op):
    import weakref
    key = id(A)
    if len(cache) > 64:
        cache.clear()
    try:
        ref = weakref.ref(A, lambda _r, k=key, c=cache: c.pop(k, None))   # a recycled id can never hit a dead entry
    except TypeError:
        return
    cache[key] = (ref, fp, op)


def invalidate(A=None):
    """Forget the operator recognised for the scipy matrix A (or all of them): call after editing a matrix in place."""
    from . import banded
    for cache in (_RECOGNISED, banded._RECOGNISED):
        if A is None:
            cache.clear()
        else:
            cache.pop(id(A), None)


def recognise(A, dimension):
    """from_sparse with a small identity cache (drivers pass the same matrix object every call)."""
    if isinstance(A, SeparableOperator):
        return A
    fp = (A.shape, dimension, getattr(A, "nnz", None)) + data_fingerprint(A)
    op = _cache_lookup(_RECOGNISED, A, fp)
    if op is None:
        op = SeparableOperator.from_sparse(A, dimension)
        _cache_store(_RECOGNISED, A, fp, op)
    return op
