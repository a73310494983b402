"""Drop-in for the reference's MGCMTProcessor (MGCMTProcessor.py:4-72) on the GPU.

numpy (n, k) in -> numpy (n, k) out, as in the reference.  A torch cuda tensor of shape (n, k) is
processed on the device and returned as a torch tensor (zero-copy when its transpose is contiguous,
i.e. when each column is stored contiguously -- the layout the kernels use).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .hierarchy import _ptr, _stream_ptr, is_device_tensor


def _to_block(vectors):
    """-> (torch cuda tensor of shape (k, n), contiguous; was_device)"""
    torch = _lib.require_cuda()
    if is_device_tensor(vectors):
        blk = vectors.t()
        if not blk.is_contiguous() or blk.dtype != torch.float64:
            blk = blk.to(torch.float64).contiguous()
        return blk, True
    a = np.asarray(vectors)
    if np.iscomplexobj(a):
        a = a.real  # quirk Q8: the reference silently drops imaginary parts (MGCMTProcessor.py:17-18,31-32)
    host = np.ascontiguousarray(a.T, dtype=np.float64)
    return torch.from_numpy(host).cuda(), False


def _from_block(blk, was_device, copy_into=None):
    if was_device:
        return blk.t()
    return np.ascontiguousarray(blk.cpu().numpy().T)


class MGCMTProcessor:

    def __init__(self):
        pass

    def projection(self, v, u):
        # MGCMTProcessor.py:10-20: (<v,u>/<u,u>) u
        torch = _lib.require_cuda()
        lib = _lib.load()
        dev = is_device_tensor(v) and is_device_tensor(u)
        from .hierarchy import to_device, to_host
        dv, du = to_device(v), to_device(u)
        n = du.numel()
        sc = torch.zeros(2, dtype=torch.float64, device="cuda")
        _lib.check(lib.mgcmt_dot(n, _ptr(dv), _ptr(du), _ptr(sc[0:1]), _stream_ptr(torch)))
        _lib.check(lib.mgcmt_dot(n, _ptr(du), _ptr(du), _ptr(sc[1:2]), _stream_ptr(torch)))
        out = (sc[0] / sc[1]) * du
        return out if dev else to_host(out).reshape(np.asarray(u).shape)

    def gramschmidt(self, vectors, modified=1):
        # MGCMTProcessor.py:22-50
        torch = _lib.require_cuda()
        blk, was_dev = _to_block(vectors)
        if was_dev and blk.data_ptr() == vectors.data_ptr():
            blk = blk.clone()  # the reference returns a new array and leaves its input alone
        k, n = blk.shape
        # modified=2 (extension): Gram-matrix / Cholesky-QR form, same Q in exact arithmetic, 3k vector passes
        mode = 2 if modified == 2 else (1 if modified else 0)
        _lib.check(_lib.load().mgcmt_gramschmidt(n, k, _ptr(blk), mode, _stream_ptr(torch)))
        return _from_block(blk, was_dev)

    def normalize(self, vectors):
        # MGCMTProcessor.py:52-63
        torch = _lib.require_cuda()
        blk, was_dev = _to_block(vectors)
        if was_dev and blk.data_ptr() == vectors.data_ptr():
            blk = blk.clone()
        k, n = blk.shape
        lib = _lib.load()
        for j in range(k):
            _lib.check(lib.mgcmt_normalize(n, _ptr(blk[j]), _stream_ptr(torch)))
        return _from_block(blk, was_dev)

    def orthogonality_check(self, vectors):
        # MGCMTProcessor.py:65-72: matrix of inner products
        torch = _lib.require_cuda()
        blk, was_dev = _to_block(vectors)
        k, n = blk.shape
        lib = _lib.load()
        out = torch.zeros(k, k, dtype=torch.float64, device="cuda")
        for i in range(k):
            for j in range(k):
                _lib.check(lib.mgcmt_dot(n, _ptr(blk[i]), _ptr(blk[j]), _ptr(out[i, j:j + 1]), _stream_ptr(torch)))
        return out if was_dev else out.cpu().numpy()
