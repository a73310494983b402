"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/mgcmt_oracle.c (matrix-free, OpenMP CPU V-cycle).

Built on demand with `gcc -O3 -march=native -fopenmp` into oracle/_build/ (git-ignored).  Used by
tests/test_c_oracle.py (held to the numpy oracle) and by the cpu_baseline / --impl reference legs of bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mgcmt_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")


def _host_tag():
    """-march=native code must not travel between machines (the repo snapshot, built files included, is copied
    to the GPU box): one library per CPU feature set."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:12]


LIB = os.path.join(OUT_DIR, "libmgcmt_oracle_%s.so" % _host_tag())
_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:   # e.g. a host without -march=native support for this gcc: retry portable
            cmd.remove("-march=native")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("gcc failed for the C oracle:\n" + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, C.c_int]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_vcycle.restype = C.c_int
        lib.orc_vcycle.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
        lib.orc_vcycle_smoother.restype = C.c_int
        lib.orc_vcycle_smoother.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
        lib.orc_block_step.restype = C.c_int
        lib.orc_block_step.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, dp, C.c_double, C.c_int, C.c_int, dp, dp, dp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class WellHierarchy:
    """(scale * laplacian(n, '2d')) hierarchy down to `lowest` (MGCMTStencilMaker.py:15-25 operator)."""

    def __init__(self, n, lowest, scale=-1.0 / np.pi ** 2):
        self.n = n
        c = scale * float(n) ** 2
        lo = np.full(n, c); di = np.full(n, -2.0 * c); up = np.full(n, c)
        self._keep = (lo, di, up)
        self.h = load().orc_create(n, _p(lo), _p(di), _p(up), _p(lo), _p(di), _p(up), int(lowest))

    def vcycle(self, v0, f, shift, nu1=4, nu2=4, omega=None, smoother="wjacobi"):
        """smoother: "wjacobi" (omega default 2/3) or "rbgs" (four-colour Gauss-Seidel / SOR, omega default 1)"""
        v = np.ascontiguousarray(v0, dtype=np.float64).copy().reshape(-1)
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        code = {"wjacobi": 0, "rbgs": 1}[smoother]
        if omega is None:
            omega = 2.0 / 3.0 if code == 0 else 1.0
        rc = load().orc_vcycle_smoother(self.h, code, float(shift), float(omega), int(nu1), int(nu2), _p(v), _p(f))
        if rc:
            raise RuntimeError("C oracle: singular coarsest operator")
        return v

    def __del__(self):
        try:
            if self.h:
                load().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass


class ShiftBlock:
    """The shift-method step of the drivers (2DPotGS.py:91-105) on a block of k vectors, all host threads: k V-cycles
    (one hierarchy per shift, so every shift keeps its coarsest LU), Rayleigh sums, modified Gram-Schmidt."""

    def __init__(self, n, lowest, shifts, smoother="wjacobi", omega=None):
        self.n, self.k = n, len(shifts)
        # a small coarsest operator is re-factored per cycle in microseconds: one hierarchy (its level vectors are
        # 4/3 x 3 grids: 4.3 GB at 16384^2) serves every shift; a 64^2 coarsest level keeps one LU per shift
        if lowest * lowest <= 1024:
            self.hs = [WellHierarchy(n, lowest)] * len(shifts)
        else:
            self.hs = [WellHierarchy(n, lowest) for _ in shifts]
        self.shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        self.code = {"wjacobi": 0, "rbgs": 1}[smoother]
        self.omega = (2.0 / 3.0 if self.code == 0 else 1.0) if omega is None else float(omega)
        self._handles = (C.c_void_p * self.k)(*[h.h for h in self.hs])
        self.lam = np.zeros(2 * self.k)

    def step(self, V, W, nu1=4, nu2=4):
        """W <- orthonormalised V-cycle outputs of V (both (k, n*n) float64 C-contiguous); returns Rayleigh quotients"""
        rc = load().orc_block_step(self._handles, self.k, self.code, _p(self.shifts), self.omega, int(nu1), int(nu2), _p(V), _p(W),
                                   _p(self.lam))
        if rc:
            raise RuntimeError("C oracle: singular coarsest operator")
        return self.lam[0::2] / self.lam[1::2]
