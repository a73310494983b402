"""Device-side grid hierarchy handle + host<->device plumbing (PyTorch is used for device memory
and streams only; every computation goes through the C ABI in include/mgcmt_b200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .operators import SeparableOperator


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _hptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Hierarchy:
    """Owns one mgcmt_hier_t: all levels of one operator for one `lowest_level`."""

    def __init__(self, op: SeparableOperator, lowest_level: int):
        torch = _lib.require_cuda()
        lib = _lib.load()
        self.op = op
        self.lowest_level = int(lowest_level)
        self.device = torch.cuda.current_device()
        handle = C.c_void_p()
        coarsen_rows = 0 if op.nrows == 1 else 1
        _lib.check(lib.mgcmt_hier_create(C.byref(handle), op.nrows, op.ncols, coarsen_rows,
                                         _hptr(op.row[0]), _hptr(op.row[1]), _hptr(op.row[2]),
                                         _hptr(op.col[0]), _hptr(op.col[1]), _hptr(op.col[2]),
                                         self.lowest_level, _stream_ptr(torch)))
        self.handle = handle
        self.num_levels = lib.mgcmt_hier_num_levels(handle)
        self.n = op.nrows * op.ncols

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().mgcmt_hier_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def level_shape(self, level):
        nr, nc = C.c_int(), C.c_int()
        _lib.check(_lib.load().mgcmt_hier_level_shape(self.handle, level, C.byref(nr), C.byref(nc)))
        return nr.value, nc.value

    def level_size(self, level):
        nr, nc = self.level_shape(level)
        return nr * nc

    def level_coefs(self, level):
        """dict of the 12 tridiagonal factor arrays of a level (host copies; for tests)."""
        nr, nc = self.level_shape(level)
        rows = np.zeros(6 * nr)
        cols = np.zeros(6 * nc)
        _lib.check(_lib.load().mgcmt_hier_level_coefs(self.handle, level, _hptr(rows), _hptr(cols)))
        names_r = ["ka_lo", "ka_di", "ka_up", "ma_lo", "ma_di", "ma_up"]
        names_c = ["kb_lo", "kb_di", "kb_up", "mb_lo", "mb_di", "mb_up"]
        out = {k: rows[i * nr:(i + 1) * nr].copy() for i, k in enumerate(names_r)}
        out.update({k: cols[i * nc:(i + 1) * nc].copy() for i, k in enumerate(names_c)})
        return out

    # ---- single-level operators on torch cuda tensors (float64, contiguous) -----------------
    def apply(self, level, shift, x, y):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_apply(self.handle, level, float(shift), _ptr(x), _ptr(y), _stream_ptr(torch)))
        return y

    def apply_mass(self, level, x, y):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_apply_mass(self.handle, level, _ptr(x), _ptr(y), _stream_ptr(torch)))
        return y

    def residual(self, level, shift, v, f, r):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_residual(self.handle, level, float(shift), _ptr(v), _ptr(f), _ptr(r),
                                              _stream_ptr(torch)))
        return r

    def smooth(self, level, smoother, shift, omega, nu, v, f):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_smooth(self.handle, level, int(smoother), float(shift), float(omega), int(nu),
                                            _ptr(v), _ptr(f), None, _stream_ptr(torch)))
        return v

    def restrict(self, level, fine, coarse):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_restrict(self.handle, level, _ptr(fine), _ptr(coarse), _stream_ptr(torch)))
        return coarse

    def residual_restrict(self, level, shift, v, f, rc):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_residual_restrict(self.handle, level, float(shift), _ptr(v), _ptr(f), _ptr(rc),
                                                       _stream_ptr(torch)))
        return rc

    def prolong(self, level, coarse, fine):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_prolong(self.handle, level, _ptr(coarse), _ptr(fine), _stream_ptr(torch)))
        return fine

    def prolong_correct(self, level, ec, v):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_prolong_correct(self.handle, level, _ptr(ec), _ptr(v), _stream_ptr(torch)))
        return v

    def coarse_solve(self, shift, f, v):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_coarse_solve(self.handle, float(shift), _ptr(f), _ptr(v), _stream_ptr(torch)))
        return v

    def vcycle(self, shift, nu1, nu2, smoother, omega, v, f, v0_is_zero=False):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_vcycle(self.handle, float(shift), int(nu1), int(nu2), int(smoother),
                                            float(omega), _ptr(v), _ptr(f), 1 if v0_is_zero else 0,
                                            _stream_ptr(torch)))
        return v

    def vcycle_host_block(self, shifts, nu1, nu2, smoother, omega, f_host, v_host):
        """zero-start cycles for k HOST vectors with their PCIe copies pipelined against each other
        (mgcmt_vcycle_host_block); f_host / v_host: lists of contiguous float64 numpy arrays of the finest level's length"""
        import ctypes as C
        torch = _lib.require_cuda()
        k = len(f_host)
        for a in list(f_host) + list(v_host):
            if a.dtype != np.float64 or a.size != self.n or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("vcycle_host_block: vectors must be contiguous float64 of length %d" % self.n)
        fp = (C.c_void_p * k)(*[a.ctypes.data for a in f_host])
        vp = (C.c_void_p * k)(*[a.ctypes.data for a in v_host])
        sh = (C.c_double * k)(*[float(x) for x in shifts])
        _lib.check(_lib.load().mgcmt_vcycle_host_block(self.handle, k, sh, int(nu1), int(nu2), int(smoother), float(omega),
                                                       fp, vp, _stream_ptr(torch)))

    def fused_leg(self, level, mode, nu, shift, omega, v_in, f, v_out, e_coarse=None, r_coarse=None):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_fused_leg(self.handle, level, int(mode), int(nu), float(shift), float(omega),
                                               _ptr(v_in) if v_in is not None else None, _ptr(f), _ptr(v_out),
                                               _ptr(e_coarse) if e_coarse is not None else None,
                                               _ptr(r_coarse) if r_coarse is not None else None,
                                               _stream_ptr(torch)))
        return v_out

    def rayleigh(self, level, x, out2):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_rayleigh(self.handle, level, _ptr(x), _ptr(out2), _stream_ptr(torch)))
        return out2


def get_hierarchy(op: SeparableOperator, lowest_level: int) -> Hierarchy:
    """Hierarchies are cached on the operator object: the reference rebuilds R, P and R*A*P on every
    call at every level (MGCMTSolver.py:310-318); here that happens once per (operator, lowest_level)."""
    torch = _lib.require_cuda()
    key = (torch.cuda.current_device(), int(lowest_level))
    h = op._device.get(key)
    if h is None:
        h = Hierarchy(op, lowest_level)
        op._device[key] = h
    return h


# ---- host <-> device -------------------------------------------------------------------------------
def is_device_tensor(x):
    try:
        import torch
    except Exception:
        return False
    return isinstance(x, torch.Tensor) and x.is_cuda


def to_device(x, n=None):
    """1-D float64 contiguous cuda tensor from numpy / torch input (copy unless already suitable)."""
    torch = _lib.require_cuda()
    if isinstance(x, torch.Tensor):
        t = x.reshape(-1)
        if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
            t = t.to(device="cuda", dtype=torch.float64).contiguous()
        if t.data_ptr() % 16:
            t = t.clone()  # the kernels read grid vectors as double2
        return t
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    src = torch.from_numpy(a)
    out = torch.empty(a.shape[0], dtype=torch.float64, device="cuda")
    out.copy_(src, non_blocking=src.is_pinned())
    return out


_PINNED_FREE = {}   # numel -> pinned float64 tensors not referenced by any live numpy array


class _PinnedOwner:
    """Owns one pinned tensor on behalf of the numpy arrays that view it.  numpy collapses `.base` chains
    to the object that exposes the memory, so every view keeps this owner alive; when the last one is
    collected the tensor goes back to the free list."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.__array_interface__ = {"shape": (tensor.numel(),), "typestr": "<f8",
                                    "data": (tensor.data_ptr(), False), "version": 3}

    def __del__(self):
        try:
            free = _PINNED_FREE.setdefault(self.tensor.numel(), [])
            if len(free) < 8:
                free.append(self.tensor)
        except Exception:
            pass


def to_host(t):
    """Device vector -> fresh numpy array the caller owns, backed by pinned host memory (one async D2H
    copy at PCIe speed instead of a pageable staging copy)."""
    torch = _lib.require_cuda()
    t = t.detach().reshape(-1)
    n = t.numel()
    if n < (1 << 16):
        return t.cpu().numpy()
    free = _PINNED_FREE.get(n)
    buf = free.pop() if free else torch.empty(n, dtype=torch.float64, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return np.asarray(_PinnedOwner(buf))


def pinned_result(n):
    """(tensor, numpy view the caller will own) of n float64 in page-locked memory, recycled like to_host's buffers"""
    torch = _lib.require_cuda()
    free = _PINNED_FREE.get(n)
    buf = free.pop() if free else torch.empty(n, dtype=torch.float64, pin_memory=True)
    return buf, np.asarray(_PinnedOwner(buf))


def lowest_for_apply(op):
    """Any hierarchy exposes level 0; reuse one if present, else build the shallowest legal one
    (the coarsest level must stay <= 4096 unknowns for the dense coarse solve)."""
    for h in op._device.values():
        return h
    low = min(op.ncols, 64 if op.nrows > 1 else 4096)
    return get_hierarchy(op, low)


def apply_operator(op, x):
    torch = _lib.require_cuda()
    h = lowest_for_apply(op)
    dev_in = is_device_tensor(x)
    xv = to_device(x)
    if xv.numel() != h.n:
        raise ValueError("vector length %d does not match operator size %d" % (xv.numel(), h.n))
    y = torch.empty_like(xv)
    h.apply(0, 0.0, xv, y)
    if dev_in:
        return y.reshape(x.shape)
    return to_host(y).reshape(np.asarray(x).shape)
