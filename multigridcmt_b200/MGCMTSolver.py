"""Drop-in for the reference's MGCMTSolver (MGCMTSolver.py:8-436) running on the B200.

Same class name, method names, positional/keyword signatures and return shapes as the reference;
inputs are numpy arrays + scipy.sparse matrices (or a matrix-free SeparableOperator), outputs are
fresh numpy arrays.  Passing torch cuda tensors instead of numpy arrays keeps the data on the device
(no host copies; a torch tensor comes back) -- that is the extension the large grids need.

Reference conventions kept on purpose (SURVEY.md section 5 and the quirk table):
  * bad sizes PRINT a message and return None (MGCMTSolver.py:303-304, :404-405);
  * `vcycle`/`twogrid`/`wjacobi` reshape the caller's numpy v0 and f to (n, 1) in place
    (MGCMTSolver.py:187-191, :297-300, :344-347);
  * `vcycle` returns (n,) -- but (n, 1) when called directly at the coarsest size (quirk Q7);
  * coarse levels ignore the caller's nu1/nu2 and run 4/4 (quirk Q4);
  * the shift is re-applied as -shift*I on every level, A is coarsened unshifted (quirk Q5);
  * `sor` adds w (D-L)^-1 f (quirk Q6).
What is NOT kept: the O(n^2) iteration matrices and the per-call rebuild of R, P, R*A*P.

Operators: the real separable radius-1 stencils of the well problems take the fused path (hierarchy.py);
any other 1-D operator -- complex, or more than three diagonals, e.g. the multiband Hamiltonians of
ThesisProblem.py:26-104 -- takes the general banded complex path (banded.py).  There is no CPU fallback: a 2-D
operator that is not separable, a foreign smoother callable or a foreign stencil_maker raises (loudly) instead
of silently running somewhere else.
"""
from __future__ import annotations

import functools

import numpy as np

from . import _lib
from .MGCMTProcessor import MGCMTProcessor
from .MGCMTStencilMaker import MGCMTStencilMaker
from .hierarchy import (_ptr, _stream_ptr, get_hierarchy, is_device_tensor, pinned_result, to_device, to_host)
from .operators import SeparableOperator, UnsupportedOperator, recognise
from .banded import (BandedOperator, get_banded_hierarchy, recognise_banded, to_device_complex)


def _is_complex(x):
    if is_device_tensor(x) or hasattr(x, "is_complex"):
        return bool(x.is_complex())
    return np.iscomplexobj(x)


class ZeroVector:
    """Stands for `np.zeros(n)` as the start vector of `vcycle` / `vcycle_matrix`: every driver of the reference starts
    its cycles from zero (`w0 = np.zeros(n)`, 2DPotGS.py:94, 2DPot.py:88), and shipping 8 n bytes of zeros over PCIe per
    call -- to have the first smoothing pass read them back from HBM -- is a third of the host traffic of a cycle.
    `solver.vcycle(ZeroVector(n), f, H, ...)` uploads nothing and takes the zero-start legs (the down leg never reads v)."""

    def __init__(self, n):
        self.n = int(n)
        self.shape = (self.n,)

    def __len__(self):
        return self.n


def _inplace_column(x, n):
    """The reference does `x.shape = (n, 1)` on the caller's array; mimic it when possible."""
    if isinstance(x, np.ndarray):
        try:
            x.shape = (n, 1)
        except (AttributeError, ValueError):
            pass


class MGCMTSolver:

    def __init__(self):
        self.stencil_maker = MGCMTStencilMaker()
        self.processor = MGCMTProcessor()

    # ------------------------------------------------------------------------------------------
    # smoother seam: which device smoother does a `smoother=` argument mean?
    # ------------------------------------------------------------------------------------------
    def _smoother_code(self, smoother):
        """-> (code, omega).  Only this class's own smoothers (optionally wrapped in functools.partial
        to pin omega) are accepted; the device cycle cannot call back into arbitrary Python."""
        if smoother is None:
            return _lib.SMOOTH_WJACOBI, 2. / 3.
        omega = None
        fn = smoother
        if isinstance(fn, functools.partial):
            omega = fn.keywords.get("omega")
            fn = fn.func
        owner = getattr(fn, "__self__", None)
        name = getattr(fn, "__name__", "")
        if isinstance(owner, MGCMTSolver):
            if name == "wjacobi":
                return _lib.SMOOTH_WJACOBI, (2. / 3. if omega is None else float(omega))
            if name == "gseidel":
                return _lib.SMOOTH_GSLEX, 1.0
            if name == "sor":
                return _lib.SMOOTH_GSLEX, (1.0 if omega is None else float(omega))
            if name in ("rbgs", "gseidelrb"):
                return _lib.SMOOTH_RBGS, (1.0 if omega is None else float(omega))
        raise NotImplementedError(
            "smoother must be one of MGCMTSolver.wjacobi/gseidel/sor/rbgs (optionally functools.partial "
            "with omega=...); arbitrary Python smoothers cannot run inside the device V-cycle")

    @staticmethod
    def _check_stencil_maker(stencil_maker):
        if not isinstance(stencil_maker, MGCMTStencilMaker):
            raise NotImplementedError(
                "stencil_maker must be a multigridcmt_b200 MGCMTStencilMaker: the device path implements "
                "exactly that class's full-weighting restriction / linear interpolation")

    # ------------------------------------------------------------------------------------------
    # operator routing: real separable stencil -> fused path; any other 1-D operator -> banded complex path
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _route(A, dimension, *vectors):
        if isinstance(A, BandedOperator):
            return A
        if dimension == "1d" and not isinstance(A, SeparableOperator):
            if any(_is_complex(x) for x in vectors) or np.iscomplexobj(getattr(A, "data", np.zeros(0))):
                return recognise_banded(A)
            try:
                return recognise(A, dimension)
            except UnsupportedOperator:
                return recognise_banded(A)
        if any(_is_complex(x) for x in vectors):
            if dimension == "1d":
                return recognise_banded(A.tocsc())
            raise UnsupportedOperator("complex vectors are supported on 1-D operators only (the reference's "
                                      "multiband problems are 1-D, ThesisProblem.py:80-101)")
        return recognise(A, dimension)

    @staticmethod
    def _banded_result(op, out, dev_in, *inputs):
        """complex128 device vector -> what the reference would hand back: complex if anything was complex."""
        real = op.is_real and not any(_is_complex(x) for x in inputs)
        if real:
            out = out.real.contiguous()
        return out if dev_in else out.cpu().numpy()

    def _smooth_banded(self, op, code, v0, f, nu, omega):
        if code == _lib.SMOOTH_RBGS:
            raise NotImplementedError("red-black Gauss-Seidel needs a radius-1 stencil; banded operators take "
                                      "wjacobi, gseidel and sor")
        n = len(v0)
        low = n            # any hierarchy exposes level 0; build the shallowest legal one
        while low > 512:
            if low % 2:
                raise UnsupportedOperator("operator size %d cannot be coarsened to <= 512 unknowns" % n)
            low //= 2
        h = get_banded_hierarchy(op, low)
        v = to_device_complex(v0)
        fd = to_device_complex(f)
        h.smooth(0, code, nu, 0.0, omega, v, fd)
        return self._banded_result(op, v, is_device_tensor(v0), v0, f).reshape(n, 1)

    def _vcycle_banded(self, op, v0, f, nu1, nu2, code, omega, shift, low):
        if code == _lib.SMOOTH_RBGS:
            raise NotImplementedError("red-black Gauss-Seidel needs a radius-1 stencil; banded operators take "
                                      "wjacobi, gseidel and sor")
        n = len(v0)
        h = get_banded_hierarchy(op, low)
        v = to_device_complex(v0)
        fd = to_device_complex(f)
        h.vcycle(shift, nu1, nu2, code, omega, v, fd)
        out = self._banded_result(op, v, is_device_tensor(v0), v0, f)
        return out.reshape(n, 1) if h.num_levels == 1 else out

    # ------------------------------------------------------------------------------------------
    # single-level smoothers (MGCMTSolver.py:182-246): return an (n, 1) array like the reference
    # ------------------------------------------------------------------------------------------
    def _smooth(self, code, v0, f, A, nu, omega, dimension=None):
        n = len(v0)
        dev_in = is_device_tensor(v0)
        if dimension is None:
            dimension = self._guess_dimension(A, n)
        op = self._route(A, dimension, v0, f)
        if isinstance(op, BandedOperator):
            return self._smooth_banded(op, code, v0, f, nu, omega)
        h = get_hierarchy(op, self._any_lowest(op))
        v = to_device(v0)
        if dev_in:
            v = v.clone()
        fd = to_device(f)
        h.smooth(0, code, 0.0, omega, nu, v, fd)
        if dev_in:
            return v.reshape(n, 1)
        return to_host(v).reshape(n, 1)

    @staticmethod
    def _guess_dimension(A, n):
        """The reference's smoothers take only the matrix; decide 1-D vs 2-D from its bandwidth."""
        if isinstance(A, (SeparableOperator, BandedOperator)):
            return A.dimension
        if np.iscomplexobj(getattr(A, "data", np.zeros(0))):
            return "1d"     # complex operators exist in 1-D only (ThesisProblem.py)
        N = int(round(np.sqrt(n)))
        if N * N == n and N > 2:
            try:
                far = A.diagonal(N)
                if np.any(far != 0):
                    return "2d"
            except Exception:
                pass
        return "1d"

    @staticmethod
    def _any_lowest(op):
        return min(op.ncols, 64 if op.nrows > 1 else 4096)

    def wjacobi(self, v0, f, A, nu=4, omega=2. / 3.):
        n = len(v0)
        out = self._smooth(_lib.SMOOTH_WJACOBI, v0, f, A, nu, omega)
        _inplace_column(f, n)   # side effect of the reference (MGCMTSolver.py:187-191)
        _inplace_column(v0, n)
        return out

    def gseidel(self, v0, f, A, nu=4):
        return self._smooth(_lib.SMOOTH_GSLEX, v0, f, A, nu, 1.0)

    def sor(self, v0, f, A, nu=4, omega=1):
        return self._smooth(_lib.SMOOTH_GSLEX, v0, f, A, nu, float(omega))

    def rbgs(self, v0, f, A, nu=4, omega=1.0):
        """Red-black (four-colour) Gauss-Seidel/SOR -- the working version of the reference's dead
        `gseidelrb` (MGCMTSolver.py:248-279)."""
        return self._smooth(_lib.SMOOTH_RBGS, v0, f, A, nu, float(omega))

    gseidelrb = rbgs

    # ------------------------------------------------------------------------------------------
    # V-cycle (MGCMTSolver.py:281-329)
    # ------------------------------------------------------------------------------------------
    def vcycle(self, v0, f, A, stencil_maker, nu1=4, nu2=4, smoother=None, shift=0, lowest_level=2,
               dimension="1d"):
        code, omega = self._smoother_code(smoother)
        self._check_stencil_maker(stencil_maker)
        n = len(v0)
        grid_dimension = 0
        if dimension == "1d":
            grid_dimension = n
        elif dimension == "2d":
            grid_dimension = np.sqrt(n)
        zero_start = isinstance(v0, ZeroVector)
        dev_in = is_device_tensor(f if zero_start else v0)
        if not dev_in:
            _inplace_column(f, n)
            if not zero_start:
                _inplace_column(v0, n)
        if grid_dimension < 2:
            print("Length of start vector is not a power of 2")
            return None
        g = int(round(grid_dimension))
        low = int(lowest_level)
        if (dimension == "2d" and g * g != n) or (g & (g - 1)) or (low & (low - 1)) or low > g or low < 2:
            # the reference would recurse until the stencil maker prints its power-of-two message
            print("Length of start vector is not a power of 2")
            return None
        op = self._route(A, dimension, f if zero_start else v0, f)
        if isinstance(op, BandedOperator):
            return self._vcycle_banded(op, np.zeros(n) if zero_start else v0, f, nu1, nu2, code, omega, shift, low)
        h = get_hierarchy(op, low)
        if zero_start and not dev_in and h.num_levels > 1 and n >= (1 << 16):
            # host vector in, host vector out: one-vector form of the pipelined block call -- a pageable f (a plain numpy
            # array) is staged to the device by several host threads instead of the driver's single-threaded bounce copy
            f1 = np.ascontiguousarray(np.asarray(f, dtype=np.float64).reshape(-1))
            buf, out = pinned_result(n)
            h.vcycle_host_block([shift], nu1, nu2, code, omega, [f1], [out])
            return out
        fd = to_device(f)
        if zero_start:
            v = _lib.require_cuda().empty(n, dtype=fd.dtype, device=fd.device)
        else:
            v = to_device(v0)
            if dev_in:
                v = v.clone()
            if fd.data_ptr() == v.data_ptr():
                fd = fd.clone()
        h.vcycle(shift, nu1, nu2, code, omega, v, fd, v0_is_zero=zero_start)
        coarsest_direct = (h.num_levels == 1)
        if dev_in:
            return v.reshape(n, 1) if coarsest_direct else v
        out = to_host(v)
        return out.reshape(n, 1) if coarsest_direct else out

    # ------------------------------------------------------------------------------------------
    # the loop body of the shift-method drivers as one call (2DPotGS.py:93-95: `w0 = zeros; w = vcycle(w0, v_i, H, ...,
    # shift=shifts[i])` for every eigenvector i).  The cycles are independent, so with host arrays the upload of vector
    # i+1 and the download of vector i-1 ride beside cycle i (mgcmt_vcycle_host_block) instead of 2k serial PCIe copies.
    # ------------------------------------------------------------------------------------------
    def vcycle_many(self, fs, A, stencil_maker, shifts, nu1=4, nu2=4, smoother=None, lowest_level=2, dimension="1d"):
        """[vcycle(ZeroVector(n), fs[i], A, stencil_maker, shift=shifts[i], ...) for i in range(len(fs))], same results
        bit for bit.  fs: sequence of vectors (or the columns of an n x k array); returns a list of arrays."""
        if isinstance(fs, np.ndarray) and fs.ndim == 2:
            fs = [np.ascontiguousarray(fs[:, i]) for i in range(fs.shape[1])]
        fs = list(fs)
        shifts = [float(x) for x in np.asarray(shifts, dtype=float).reshape(-1)]
        if len(shifts) != len(fs):
            raise ValueError("vcycle_many: %d right-hand sides but %d shifts" % (len(fs), len(shifts)))
        if not fs:
            return []
        n = int(np.prod(fs[0].shape))
        code, omega = self._smoother_code(smoother)
        one_by_one = lambda: [self.vcycle(ZeroVector(n), f, A, stencil_maker, nu1=nu1, nu2=nu2, smoother=smoother, shift=sh,
                                          lowest_level=lowest_level, dimension=dimension) for f, sh in zip(fs, shifts)]
        g = int(round(np.sqrt(n))) if dimension == "2d" else n
        low = int(lowest_level)
        if (any(is_device_tensor(f) for f in fs) or any(int(np.prod(f.shape)) != n for f in fs) or n < (1 << 16)
                or (dimension == "2d" and g * g != n) or g < 2 or (g & (g - 1)) or (low & (low - 1)) or low > g or low < 2):
            return one_by_one()      # device tensors, small or ill-sized input: nothing to pipeline / the checks of vcycle
        self._check_stencil_maker(stencil_maker)
        op = self._route(A, dimension, fs[0], fs[0])
        if isinstance(op, BandedOperator):
            return one_by_one()
        h = get_hierarchy(op, low)
        if h.num_levels == 1:
            return one_by_one()
        f_host = [np.ascontiguousarray(np.asarray(f, dtype=np.float64).reshape(-1)) for f in fs]
        outs = [pinned_result(n) for _ in fs]
        h.vcycle_host_block(shifts, nu1, nu2, code, omega, f_host, [o[1] for o in outs])
        return [o[1] for o in outs]

    # ------------------------------------------------------------------------------------------
    # two-grid cycle (MGCMTSolver.py:331-371) == a 2-level V-cycle with exact coarse solve
    # ------------------------------------------------------------------------------------------
    def twogrid(self, v0, f, A, stencil_maker, nu1=4, nu2=4, smoother=None, shift=0, dimension="1d"):
        if dimension != "1d":
            # quirk Q10: the reference's coarse shift matrix is eye(n/2), which mismatches the 2-D coarse
            # size n/4 -- it cannot run in 2-D there either.
            raise NotImplementedError("twogrid is 1-D only, as in the reference (MGCMTSolver.py:350)")
        n = len(v0)
        if n < 4 or n & (n - 1):
            print("Length of start vector is not a power of 2")
            return None
        return self.vcycle(v0, f, A, stencil_maker, nu1=nu1, nu2=nu2, smoother=smoother, shift=shift,
                           lowest_level=n // 2, dimension="1d")

    # ------------------------------------------------------------------------------------------
    # block V-cycle with Gram-Schmidt on the way up (MGCMTSolver.py:375-436)
    # ------------------------------------------------------------------------------------------
    def vcycle_matrix(self, v0_matrix, f_matrix, A, stencil_maker, nu1=4, nu2=4, smoother=None, shifts=None,
                      lowest_level=2, dimension="1d"):
        torch = _lib.require_cuda()
        code, omega = self._smoother_code(smoother)
        self._check_stencil_maker(stencil_maker)
        n = v0_matrix.shape[0]
        k = f_matrix.shape[1]
        if shifts is None:
            shifts = np.zeros(k)
        shifts = np.asarray(shifts.cpu() if is_device_tensor(shifts) else shifts, dtype=float).reshape(-1)
        grid_dimension = n if dimension == "1d" else np.sqrt(n)
        if grid_dimension < 2:
            print("Length of start vector is not a power of 2")
            return None
        g = int(round(grid_dimension))
        low = int(lowest_level)
        if (dimension == "2d" and g * g != n) or g & (g - 1) or low & (low - 1) or low > g or low < 2:
            print("Length of start vector is not a power of 2")
            return None
        op = recognise(A, dimension)
        h = get_hierarchy(op, low)
        dev_in = is_device_tensor(v0_matrix)

        def block(m):
            if is_device_tensor(m):
                return m.t().to(torch.float64).contiguous().clone()
            return torch.from_numpy(np.ascontiguousarray(np.asarray(m, dtype=np.float64).T)).cuda()
        V = block(v0_matrix)
        F = block(f_matrix)
        lib = _lib.load()

        def cycle(level, V, F, n1, n2, zero_start=False):
            nl = h.level_size(level)
            if level == h.num_levels - 1:
                for i in range(k):
                    h.coarse_solve(shifts[i], F[i], V[i])
                return V
            nr_, nc_ = h.level_shape(level)
            if code == _lib.SMOOTH_WJACOBI and nr_ >= 32 and nc_ >= 32 and 0 < n1 <= 4 and 0 < n2 <= 4:
                # one fused pass per leg and column (csrc/fused.cu): sweeps + residual/restriction, then
                # interpolation/correction + sweeps; Gram-Schmidt of the block on the way up (MGCMTSolver.py:434)
                ncs = h.level_size(level + 1)
                Tm = torch.empty_like(V)
                Rc = torch.empty(k, ncs, dtype=torch.float64, device="cuda")
                for i in range(k):
                    h.fused_leg(level, 2 if zero_start else 1, n1, shifts[i], omega, None if zero_start else V[i], F[i],
                                Tm[i], None, Rc[i])
                E = cycle(level + 1, torch.zeros(k, ncs, dtype=torch.float64, device="cuda"), Rc, 4, 4, True)
                for i in range(k):
                    h.fused_leg(level, 3, n2, shifts[i], omega, Tm[i], F[i], V[i], E[i], None)
                _lib.check(lib.mgcmt_gramschmidt(nl, k, _ptr(V), 1, _stream_ptr(torch)))
                return V
            for i in range(k):
                h.smooth(level, code, shifts[i], omega, n1, V[i], F[i])
            nc = h.level_size(level + 1)
            Rc = torch.empty(k, nc, dtype=torch.float64, device="cuda")
            for i in range(k):
                h.residual_restrict(level, shifts[i], V[i], F[i], Rc[i])
            E = cycle(level + 1, torch.zeros(k, nc, dtype=torch.float64, device="cuda"), Rc, 4, 4, True)
            for i in range(k):
                h.prolong_correct(level, E[i], V[i])
                h.smooth(level, code, shifts[i], omega, n2, V[i], F[i])
            _lib.check(lib.mgcmt_gramschmidt(nl, k, _ptr(V), 1, _stream_ptr(torch)))
            return V
        V = cycle(0, V, F, nu1, nu2)
        if dev_in:
            return V.t()
        return np.ascontiguousarray(V.cpu().numpy().T)

    # ------------------------------------------------------------------------------------------
    # Rayleigh-quotient minimisation family (MGCMTSolver.py:17-122).  1-D only, as in the reference (D6).
    # Vector work (operator / mass applies, dots, updates, transfers) runs on the device through the C ABI;
    # the 2x2 generalised eigenproblem of each CG step is solved on the host with the same
    # scipy.linalg.eig call the reference makes (MGCMTSolver.py:48).
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _is_identity(M, n):
        import scipy.sparse as sp
        if M is None:
            return True
        if sp.issparse(M):
            return M.shape == (n, n) and (M - sp.eye(n)).count_nonzero() == 0
        M = np.asarray(M)
        return M.shape == (n, n) and np.array_equal(M, np.eye(n))

    def _rq_hierarchy(self, A, M, n, lowest, dimension="1d"):
        if not self._is_identity(M, n):
            raise NotImplementedError("the device RQ path needs M = identity on the finest level (as in RQMin.py:17); "
                                      "coarse mass matrices R*M*P are built from it")
        op = recognise(A, dimension)
        return get_hierarchy(op, lowest)

    def _rqmin_level(self, h, level, x, nu):
        """rqmin (MGCMTSolver.py:17-57) for the level operators A_l, M_l of hierarchy h; x: device vector (updated in
        place).  One C call (mgcmt_rqmin): the whole iteration -- mat-vecs, the 8 pencil sums in one pass, the 2 x 2
        generalised eigenproblem, updates -- stays on the device; rho comes back as a device pair (x^T A x, x^T M x)."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        use_mass = 1 if level > 0 else 0          # M_0 = I (checked by _rq_hierarchy); coarser levels: M_l = R M P
        need = int(lib.mgcmt_rqmin_work_doubles(h.handle, level, use_mass))
        cache = self.__dict__.setdefault("_rq_work", {})
        key = (id(h), level)
        work = cache.get(key)
        if work is None or work.numel() < need:
            if len(cache) > 32:
                cache.clear()
            work = torch.empty(need + 2, dtype=torch.float64, device="cuda")
            cache[key] = work
        rq2 = torch.empty(2, dtype=torch.float64, device="cuda")
        _lib.check(lib.mgcmt_rqmin(h.handle, level, use_mass, _ptr(x), int(nu), _ptr(work), work.numel(), _ptr(rq2), _stream_ptr(torch)))
        return x, rq2

    @staticmethod
    def _rho(rq2):
        r = rq2.cpu().tolist()
        return r[0] / r[1]

    def rqmin(self, A, v0, M=None, nu=4):
        # MGCMTSolver.py:17-57 (M=None is unusable in the reference, quirk Q9; here it means the identity)
        shape = np.shape(v0) if not is_device_tensor(v0) else None
        x = to_device(v0)
        if is_device_tensor(v0):
            x = x.clone()
        n = x.numel()
        dimension = self._guess_dimension(A, n)
        if dimension == "2d":
            op = recognise(A, "2d")
            h = self._rq_hierarchy(op, M, n, self._any_lowest(op), "2d")
        else:
            h = self._rq_hierarchy(A, M, n, min(n, 4096))
        x, rq2 = self._rqmin_level(h, 0, x, nu)
        rho = self._rho(rq2)
        if shape is None:
            return x, rho
        return to_host(x).reshape(shape), rho

    def _rqmg_level(self, h, level, k, nu1, nu2, nmin):
        torch = _lib.require_cuda()
        n = k.numel()
        k, rho = self._rqmin_level(h, level, k, nu1)
        if n > nmin and level + 1 < h.num_levels:
            kc = torch.empty(h.level_size(level + 1), dtype=torch.float64, device="cuda")
            h.restrict(level, k, kc)                      # k_coarse = R k: the restricted ITERATE (MGCMTSolver.py:113)
            c, rho = self._rqmg_level(h, level + 1, kc, nu1, nu2, nmin)
            h.prolong_correct(level, c, k)                # k = k + P c (:116-118)
            k, rho = self._rqmin_level(h, level, k, nu2)
        return k, rho

    def vcycle_rqmg(self, x, A, M, nu1=4, nu2=4, nmin=2, dimension="1d"):
        # MGCMTSolver.py:99-122.  dimension="2d" is an extension: the reference's RQMG is
        # 1-D only (its transfer operators are built without `dimension=`, D6); the 2-D form uses the same recursion with
        # the 2-D restriction / interpolation and stops coarsening when the VECTOR length is <= nmin, like the original.
        shape = np.shape(x) if not is_device_tensor(x) else None
        k = to_device(x)
        if is_device_tensor(x):
            k = k.clone()
        n = k.numel()
        if dimension == "2d":
            N = int(round(np.sqrt(n)))
            if N * N != n or N < 2 or N & (N - 1):
                print("New gridsize isn't a power of 2 !")
                return None
            low = N
            while low > 2 and low * low > max(int(nmin), 4):
                low //= 2
            h = self._rq_hierarchy(A, M, n, low, "2d")
            k, rho = self._rqmg_level(h, 0, k, nu1, nu2, max(int(nmin), low * low))
            rho = self._rho(rho)
            if shape is None:
                return k, rho
            return to_host(k).reshape(shape), rho
        if n < 2 or n & (n - 1):
            print("New gridsize isn't a power of 2 !")
            return None
        # the reference recurses while n > nmin (MGCMTSolver.py:105): the coarsest size is the largest power of two <= nmin
        low = 1 << (max(2, int(nmin)).bit_length() - 1)
        h = self._rq_hierarchy(A, M, n, min(low, n))
        k, rho = self._rqmg_level(h, 0, k, nu1, nu2, nmin)
        rho = self._rho(rho)
        if shape is None:
            return k, rho
        return to_host(k).reshape(shape), rho

    def vcycle_rqmg2(self, x_matrix, A, M, nu1=4, nu2=4, nmin=2, level=0):
        # MGCMTSolver.py:59-94 (block variant; Gram-Schmidt x4 on the finest level only)
        torch = _lib.require_cuda()
        dev_in = is_device_tensor(x_matrix)
        xm = x_matrix if dev_in else np.asarray(x_matrix, dtype=np.float64)
        n, nv = xm.shape
        low = 1 << (max(2, int(nmin)).bit_length() - 1)   # largest power of two <= nmin, as in vcycle_rqmg
        h = self._rq_hierarchy(A, M, n, min(low, n))
        K = (xm.t().contiguous().clone() if dev_in
             else torch.from_numpy(np.ascontiguousarray(xm.T)).cuda())
        lib = _lib.load()

        def rec(lv, K):
            nl = K.shape[1]
            for i in range(nv):
                xi, _ = self._rqmin_level(h, lv, K[i].clone(), nu1)
                K[i].copy_(xi)
            if lv == 0:
                for _ in range(4):
                    _lib.check(lib.mgcmt_gramschmidt(nl, nv, _ptr(K), 1, _stream_ptr(torch)))
            if nl > nmin and lv + 1 < h.num_levels:
                Kc = torch.empty(nv, h.level_size(lv + 1), dtype=torch.float64, device="cuda")
                for i in range(nv):
                    h.restrict(lv, K[i], Kc[i])
                Cc = rec(lv + 1, Kc)
                for i in range(nv):
                    h.prolong_correct(lv, Cc[i], K[i])
                    xi, _ = self._rqmin_level(h, lv, K[i].clone(), nu2)
                    K[i].copy_(xi)
            return K
        K = rec(0, K)
        return K.t() if dev_in else np.ascontiguousarray(K.cpu().numpy().T)

    # ------------------------------------------------------------------------------------------
    # Rayleigh quotient helper (the `v^T H v` the drivers compute inline, e.g. 2DPotGS.py:103)
    # ------------------------------------------------------------------------------------------
    def rayleigh_quotient(self, A, v, dimension="1d"):
        """(v^T A v) / (v^T v) on the device; returns a Python float."""
        torch = _lib.require_cuda()
        op = recognise(A, dimension)
        h = get_hierarchy(op, self._any_lowest(op)) if not op._device else next(iter(op._device.values()))
        x = to_device(v)
        out = torch.zeros(2, dtype=torch.float64, device="cuda")
        h.rayleigh(0, x, out)
        num, den = out.cpu().tolist()
        return num / den
