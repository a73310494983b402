"""The shift-method eigen-iteration, device-resident.

The reference's drivers wrap the V-cycle in this loop (2DPotGS.py:84-108, 1DPotGS.py, main.py:86-104):

    for it in range(max_iters):
        for i in range(k):
            w = solver.vcycle(w0, V[:, i], H, stencil_maker, shift=mu[i], ...)   # w0 = 0: one step of (H - mu_i)^-1
            V[:, i] = w / norm(w)
            eigenvalues[it, i] = V[:, i].T (H V[:, i])
        V = processor.gramschmidt(V)

with `mu` = eigenvalues of a coarse-grid problem and V = their interpolated eigenvectors.  Each pass through the
reference's classes costs a host round trip per call; `ShiftMethod` keeps the block in HBM and issues the same
arithmetic through the C ABI: k V-cycles (independent until the orthonormalisation, so each runs on its own CUDA
stream with its own level buffers), the Rayleigh-quotient sums taken inside the finest up leg (`mgcmt_vcycle_rq`),
then `mgcmt_gramschmidt` on the block.  bench.py times exactly this object.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .hierarchy import Hierarchy, _ptr, _stream_ptr, get_hierarchy, is_device_tensor
from .operators import recognise

_SMOOTHERS = {"wjacobi": (_lib.SMOOTH_WJACOBI, 2.0 / 3.0), "rbgs": (_lib.SMOOTH_RBGS, 1.0),
              "gseidel": (_lib.SMOOTH_GSLEX, 1.0), "sor": (_lib.SMOOTH_GSLEX, 1.0)}
_ORTHO = {"cgs": 0, "mgs": 1, "gram": 2}


def well_eigenvalue_1d(n, k):
    """k-th eigenvalue of (-1/pi^2) * laplacian(n) (MGCMTStencilMaker.py:15-25: h = 1/n, Dirichlet ends)."""
    return (4.0 * n * n / np.pi ** 2) * np.sin(k * np.pi / (2.0 * (n + 1))) ** 2


def well_eigenvector_1d(n, k):
    v = np.sin(k * np.pi * (np.arange(n) + 1.0) / (n + 1.0))
    return v / np.linalg.norm(v)


def well_start_block(N, modes, N0=16, dimension="2d"):
    """Start block and shifts the way the drivers get them (2DPotGS.py:56-77), with the coarse `eigsh` replaced by
    the closed-form spectrum of the N0 grid: (k, n) vectors P(N0 -> N) * eigvec_N0, normalised, and mu = eig_N0."""
    from .MGCMTStencilMaker import MGCMTStencilMaker
    P = MGCMTStencilMaker().interpolation(N0, N).toarray()
    if dimension == "1d":
        V = np.stack([P @ well_eigenvector_1d(N0, a) for a in modes])
        shifts = [well_eigenvalue_1d(N0, a) for a in modes]
    else:
        V = np.stack([np.kron(P @ well_eigenvector_1d(N0, a), P @ well_eigenvector_1d(N0, b)) for a, b in modes])
        shifts = [well_eigenvalue_1d(N0, a) + well_eigenvalue_1d(N0, b) for a, b in modes]
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    return V, shifts


class ShiftMethod:
    """k eigenpairs of H near the fixed shifts `shifts`, iterated on the device.

    H: scipy.sparse matrix or SeparableOperator (real separable radius-1 stencil);  V0: start block, (k, n)
    vector-major (numpy or cuda tensor; it is copied).  smoother: "wjacobi" | "rbgs" | "gseidel" | "sor";
    ortho: "mgs" (MGCMTProcessor.gramschmidt, modified=1), "cgs" (modified=0) or "gram" (Gram-matrix form, k <= 6).
    """

    def __init__(self, H, shifts, V0, dimension="2d", lowest_level=8, nu1=4, nu2=4, smoother="wjacobi", omega=None,
                 ortho="mgs", streams=None):
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        self.op = recognise(H, dimension)
        self.shifts = [float(s) for s in shifts]
        self.k = len(self.shifts)
        self.n = self.op.nrows * self.op.ncols
        self.nu1, self.nu2 = int(nu1), int(nu2)
        self.set_smoother(smoother, omega)
        self.ortho = _ORTHO[ortho]
        if is_device_tensor(V0):
            blk = V0.to(dtype=torch.float64).reshape(self.k, self.n).clone()
        else:
            blk = torch.from_numpy(np.ascontiguousarray(V0, dtype=np.float64).reshape(self.k, self.n)).cuda()
        self.blocks = [blk, torch.zeros_like(blk)]      # blocks[cur] = V (input), blocks[1 - cur] = W (output)
        self.cur = 0
        self.rq = torch.zeros(self.k, 2, dtype=torch.float64, device="cuda")   # (w^T H w, w^T w) of the last step
        nstreams = self.k if streams is None else max(1, min(int(streams), self.k))
        # every stream gets its own hierarchy (level work vectors + the cached coarsest inverse of its shifts)
        self.hier = [get_hierarchy(self.op, lowest_level)] + [Hierarchy(self.op, lowest_level) for _ in range(nstreams - 1)]
        self.streams = [torch.cuda.Stream() for _ in range(nstreams)]
        self.iterations = 0

    def set_smoother(self, smoother, omega=None):
        code, default = _SMOOTHERS[smoother]
        self.smoother = smoother
        self.code = code
        self.omega = default if omega is None else float(omega)

    @property
    def block(self):
        """the current (k, n) block of eigenvector estimates, on the device"""
        return self.blocks[self.cur]

    def step(self, serial=False):
        """one outer iteration: k V-cycles (zero start, right-hand side = current estimate) with their Rayleigh
        quotients, then the block orthonormalisation.  Asynchronous; serial=True keeps everything on the current
        stream (per-kernel timing)."""
        torch, lib = self._torch, self._lib
        V, W = self.blocks[self.cur], self.blocks[1 - self.cur]
        main = torch.cuda.current_stream()
        ns = len(self.streams)
        if not serial:
            for s in self.streams:
                s.wait_stream(main)
        for c in range(self.k):
            with torch.cuda.stream(main if serial else self.streams[c % ns]):
                # w0 = 0 as in the drivers (2DPotGS.py:94): flagged, so the zero vector is never read.  The
                # normalisation w / ||w|| (2DPotGS.py:96) is what the orthonormalisation does to every column anyway;
                # the Rayleigh quotient is formed from the two sums.
                _lib.check(lib.mgcmt_vcycle_rq(self.hier[c % ns].handle, self.shifts[c], self.nu1, self.nu2, self.code,
                                               self.omega, _ptr(W[c]), _ptr(V[c]), 1, _ptr(self.rq[c]), _stream_ptr(torch)))
        if not serial:
            for s in self.streams:
                main.wait_stream(s)
        _lib.check(lib.mgcmt_gramschmidt(self.n, self.k, _ptr(W), self.ortho, _stream_ptr(torch)))
        self.cur = 1 - self.cur
        self.iterations += 1

    def iterate(self, iters):
        """`iters` steps; returns the (iters, k) history of Rayleigh quotients taken right after each V-cycle
        (`eigenvalues[iters, i]`, 2DPotGS.py:103) with a single device->host copy at the end."""
        torch = self._torch
        hist = torch.zeros(iters, self.k, 2, dtype=torch.float64, device="cuda")
        for it in range(iters):
            self.step()
            hist[it].copy_(self.rq)
        h = hist.cpu().numpy()
        return h[:, :, 0] / h[:, :, 1]

    # ---- iteration to a tolerance (the reference has no convergence criterion) -----------
    def _rayleigh_block(self, out):
        """out[c] = (v_c^T H v_c, v_c^T v_c) of the current block, one fused pass per vector (no host sync)"""
        for c in range(self.k):
            self.hier[0].rayleigh(0, self.block[c], out[c])

    def _residual_block(self, rq, R, rr):
        """R[c] = H v_c - rho_c v_c with rho_c = rq[c,0] / rq[c,1] read on the device; rr[c] = ||R[c]||^2"""
        torch, lib = self._torch, self._lib
        for c in range(self.k):
            _lib.check(lib.mgcmt_eigen_residual(self.hier[0].handle, 0, _ptr(self.block[c]), _ptr(rq[c]), _ptr(R[c]),
                                                _ptr(rr[c]), _stream_ptr(torch)))

    def _check_ortho(self):
        """Gram-matrix orthonormalisation broke down since the last check (mgcmt_ortho_status)?  Then the block is redone
        column by column (MGCMTProcessor.gramschmidt, modified=1) -- the ordering the Gram form reproduces."""
        import ctypes as C
        torch, lib = self._torch, self._lib
        flag = C.c_int(0)
        _lib.check(lib.mgcmt_ortho_status(C.byref(flag), _stream_ptr(torch)))
        if flag.value:
            self.ortho_breakdowns = getattr(self, "ortho_breakdowns", 0) + 1
            _lib.check(lib.mgcmt_gramschmidt(self.n, self.k, _ptr(self.block), 1, _stream_ptr(torch)))
        return bool(flag.value)

    def correction_step(self, rq, R, rr):
        """One iteration in correction form (fixed shifts):  r_i = H v_i - rho_i v_i,  v_i <- v_i - Vcycle_{mu_i}(0, r_i),
        orthonormalise.  With an exact solve in place of the V-cycle this is the reference's inverse iteration
        ((H - mu)^-1 r = v - (mu - rho)... = v + (mu - rho)(H - mu)^-1 v, so v - that is parallel to (H - mu)^-1 v,
        2DPotGS.py:95-96); with the V-cycle as approximate inverse its fixed points are EXACT eigenvectors (r = 0),
        whereas the plain form converges to the dominant eigenvector of the V-cycle operator itself, which differs
        from H's at the level of the cycle's mode mixing.  A labelled departure from the reference's loop."""
        torch, lib = self._torch, self._lib
        self._rayleigh_block(rq)
        self._residual_block(rq, R, rr)
        W = self.blocks[1 - self.cur]
        main = torch.cuda.current_stream()
        ns = len(self.streams)
        for s in self.streams:
            s.wait_stream(main)
        for c in range(self.k):
            with torch.cuda.stream(self.streams[c % ns]):
                _lib.check(lib.mgcmt_vcycle(self.hier[c % ns].handle, self.shifts[c], self.nu1, self.nu2, self.code, self.omega,
                                            _ptr(W[c]), _ptr(R[c]), 1, _stream_ptr(torch)))
                _lib.check(lib.mgcmt_axpby(self.n, 1.0, _ptr(self.block[c]), -1.0, _ptr(W[c]), _ptr(W[c]), _stream_ptr(torch)))
        for s in self.streams:
            main.wait_stream(s)
        _lib.check(lib.mgcmt_gramschmidt(self.n, self.k, _ptr(W), self.ortho, _stream_ptr(torch)))
        self.cur = 1 - self.cur
        self.iterations += 1

    def solve(self, tol=1e-10, max_iters=200, form="reference", update_shift=False, check_every=1, exact=None):
        """Iterate until every eigenpair satisfies ||H v - rho v||_2 <= tol (||v|| = 1, rho = v^T H v) and, when the exact
        eigenvalues are given, |rho - exact| <= tol.  form="reference": the drivers' loop (step(): w = Vcycle(0, v), normalise,
        Gram-Schmidt -- 2DPotGS.py:91-105), which has no stopping rule of its own; form="correction": correction_step().
        update_shift=True replaces mu_i by the current rho_i at every check (Rayleigh-quotient iteration; departs from the
        reference, which keeps the coarse-grid eigenvalues as shifts; each new shift costs one coarsest-level inverse).
        Returns a dict: converged, iterations, eigenvalues, residual_norms, history [(iteration, residual norms, rho)]."""
        torch = self._torch
        rq = torch.zeros(self.k, 2, dtype=torch.float64, device="cuda")
        rr = torch.zeros(self.k, 1, dtype=torch.float64, device="cuda")
        R = torch.empty(self.k, self.n, dtype=torch.float64, device="cuda")
        history = []
        converged = False
        it0 = self.iterations
        res = rho = None
        while True:
            # convergence check on the current (orthonormal) block
            self._rayleigh_block(rq)
            self._residual_block(rq, R, rr)
            self._check_ortho()
            rq_h, rr_h = rq.cpu().numpy(), rr.cpu().numpy()[:, 0]
            rho = rq_h[:, 0] / rq_h[:, 1]
            res = np.sqrt(rr_h / rq_h[:, 1])
            history.append((self.iterations - it0, res.copy(), rho.copy()))
            ok = bool(np.all(res <= tol))
            if exact is not None:
                ok = ok and bool(np.all(np.abs(rho - np.asarray(exact)) <= tol))
            if ok:
                converged = True
                break
            if self.iterations - it0 >= max_iters:
                break
            if update_shift:
                self.shifts = [float(x) for x in rho]
            for _ in range(max(1, int(check_every))):
                if form == "correction":
                    self.correction_step(rq, R, rr)
                else:
                    self.step()
        return {"converged": converged, "iterations": self.iterations - it0, "eigenvalues": rho, "residual_norms": res,
                "history": history, "form": form, "update_shift": bool(update_shift)}

    def last_rayleigh(self):
        """Rayleigh quotients of the last step's V-cycle outputs (before the orthonormalisation)"""
        r = self.rq.cpu().numpy()
        return r[:, 0] / r[:, 1]

    def eigenvalues(self):
        """v_i^T H v_i of the current (orthonormalised) block -- `eigenvalues_MG`, 2DPotGS.py:107-108"""
        torch = self._torch
        out = torch.zeros(self.k, 2, dtype=torch.float64, device="cuda")
        for c in range(self.k):
            self.hier[0].rayleigh(0, self.block[c], out[c])
        r = out.cpu().numpy()
        return r[:, 0] / r[:, 1]

    def residual_norms(self):
        """||H v_i - rho_i v_i||_2 with rho_i = v_i^T H v_i / v_i^T v_i (the convergence measure of SURVEY.md 8(d))"""
        torch = self._torch
        rho = self.eigenvalues()
        y = torch.empty(self.n, dtype=torch.float64, device="cuda")
        out = []
        for c in range(self.k):
            self.hier[0].apply(0, float(rho[c]), self.block[c], y)
            out.append(float(y.norm() / self.block[c].norm()))
        return np.array(out)

    def vectors(self):
        """(n, k) numpy array, columns = eigenvector estimates (the reference's `eigenvectors_MG` layout)"""
        return self.block.t().contiguous().cpu().numpy()
