"""Drop-in for the reference's MGCMTStencilMaker (MGCMTStencilMaker.py:5-78): same class name,
methods, argument meaning, return types (scipy.sparse CSC) and print-and-return-None error
behaviour.  These are host-side *descriptions* of the operators; the device path never multiplies by
them -- it recognises them (operators.py) and runs the stencil kernels instead.

Extension: `laplacian(n, dimension, matrix_free=True)` returns a SeparableOperator (no n^2-row scipy
matrix), for the grid sizes the reference could not reach.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as spsparse

from .operators import SeparableOperator


class MGCMTStencilMaker:

    def __init__(self):
        pass

    def laplacian(self, n, dimension="1d", matrix_free=False):
        # MGCMTStencilMaker.py:15-25
        if matrix_free:
            return SeparableOperator.laplacian(n, dimension)
        n = int(n)
        h = 1. / n
        laplacian = None
        if dimension == "1d":
            laplacian = spsparse.diags([1., -2., 1.], [-1, 0, 1], shape=(n, n), format="csc", dtype=float)
            laplacian *= (1 / h ** 2)
        elif dimension == "2d":
            one_d = self.laplacian(n, dimension="1d")
            laplacian = spsparse.kronsum(one_d, one_d, format="csc")
        return laplacian

    @staticmethod
    def _log2(x):
        return math.log(x) / math.log(2)

    def interpolation(self, old_gridsize, new_gridsize, dimension="1d"):
        # MGCMTStencilMaker.py:27-54: coarse j sits on fine m(j+1)-1, hat weights (m-|d|)/m
        new_gridsize = int(new_gridsize)
        if dimension == "1d":
            p_old, p_new = self._log2(old_gridsize), self._log2(new_gridsize)
            if p_new > p_old:
                if float(p_old).is_integer():
                    if float(p_new).is_integer():
                        m = int(new_gridsize / old_gridsize)
                        prefactor = (1. / 2) ** (p_new - p_old)
                        centres = np.arange(m - 1, new_gridsize, m)
                        offs = np.arange(-(m - 1), m)
                        rows = centres[None, :] + offs[:, None]
                        vals = np.broadcast_to((prefactor * (m - np.abs(offs)).astype(float))[:, None], rows.shape)
                        cols = np.broadcast_to(np.arange(len(centres))[None, :], rows.shape)
                        keep = (rows >= 0) & (rows < new_gridsize)
                        return spsparse.csc_matrix((vals[keep], (rows[keep], cols[keep])),
                                                   shape=(new_gridsize, len(centres)), dtype=float)
                    else:
                        print("New gridsize isn't a power of 2 !")
                else:
                    print("Old gridsize isn't a power of 2 !")
            else:
                print("New gridsize isn't bigger than old gridsize !")
        elif dimension == "2d":
            S = self.interpolation(old_gridsize, new_gridsize, dimension="1d")
            if S is None:
                return None
            return spsparse.kron(S, S, format="csc")

    def restriction(self, old_gridsize, new_gridsize, dimension="1d"):
        # MGCMTStencilMaker.py:57-78: 1-D (1/2)^p P^T; 2-D fixed 1/4 (P (x) P)^T (quirk Q3)
        if dimension == "1d":
            p_old, p_new = self._log2(old_gridsize), self._log2(new_gridsize)
            if p_new < p_old:
                if float(p_old).is_integer():
                    if float(p_new).is_integer():
                        prefactor = (1. / 2) ** (p_old - p_new)
                        P = self.interpolation(new_gridsize, old_gridsize)
                        return spsparse.csc_matrix(prefactor * P.T)
                    else:
                        print("New gridsize isn't a power of 2 !")
                else:
                    print("Old gridsize isn't a power of 2 !")
            else:
                print("New gridsize is bigger (more elements) than old gridsize !")
        elif dimension == "2d":
            P = self.interpolation(new_gridsize, old_gridsize, dimension="2d")
            if P is None:
                return None
            return 1. / 4. * P.T
