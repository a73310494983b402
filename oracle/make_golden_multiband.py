"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/multiband_golden.npz from the REAL reference.

Run in the build container (needs /root/reference):   python oracle/make_golden_multiband.py
Builds the reference's own multiband (4-band Luttinger-Kohn) quantum-well Hamiltonians with
PotWellSolver.makeMatrix (PotWellSolver.py:54-233) the way ThesisProblem.py:26-40 does, and records what
the reference's MGCMTSolver (through oracle/ref_loader.py's syntax-only translation) makes of them:
single-level smoothers, the Galerkin coarse operator, V-cycles with every smoother and the fixed-shift
inverse iteration of ThesisProblem.py:84-104.  Every case stores its inputs (the matrix as dense complex:
256 x 256), so the tests never regenerate them and never need /root/reference.
"""
from __future__ import annotations

import functools
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
GRID = 64          # points per band -> 256 unknowns
BAD = 32           # coarse grid the shift guesses come from (ThesisProblem.py:62)


def crand(n, seed):
    r = np.random.RandomState(seed)
    return r.random_sample(n) - 0.5 + 1j * (r.random_sample(n) - 0.5)


def main():
    warnings.simplefilter("ignore")
    SM, S, P = ref_loader.load_reference()
    bc, pw, ps = ref_loader.load_potwell()
    sm, solver = SM(), S()
    g = {}
    cases = {"z0": ("z", 0.0), "z7": ("z", 0.7), "x7": ("x", 0.7)}
    for tag, (direction, k) in cases.items():
        pws = ps.PotWellSolver(bc.Compound(bc.GaAsValues), pw.PotentialWell(direction), 4)
        pws.setGridPoints(GRID); pws.setXRange(-1, 1); pws.setDense(0)
        H = pws.makeMatrix(k)
        pws.setGridPoints(BAD); pws.setXRange(-1, 1)
        Hbad = pws.makeMatrix(k)
        n = H.shape[0]
        g[tag + "_H"] = H.toarray()
        mus = np.linalg.eigvalsh(Hbad.toarray())
        g[tag + "_mus"] = mus
        shift = float(mus[2])
        g[tag + "_shift"] = shift
        # transfer + Galerkin (MGCMTSolver.py:310-311,318)
        R = sm.restriction(n, n // 2); Pm = sm.interpolation(n // 2, n)
        g[tag + "_RAP"] = (R * H * Pm).toarray()
        x = crand(n, 1); f = crand(n, 2)
        g[tag + "_x"] = x; g[tag + "_f"] = f
        g[tag + "_Hx"] = H.dot(x)
        # single-level smoothers on the shifted matrix (MGCMTSolver.py:182-246)
        As = (H - sp.eye(n) * shift).tocsc()
        g[tag + "_wj"] = np.asarray(solver.wjacobi(x.copy(), f.copy(), As, nu=3)).reshape(-1)
        g[tag + "_gs"] = np.asarray(solver.gseidel(x.copy().reshape(n, 1), f.copy().reshape(n, 1), As, nu=3)).reshape(-1)
        g[tag + "_sor"] = np.asarray(solver.sor(x.copy().reshape(n, 1), f.copy().reshape(n, 1), As, nu=3, omega=1.3)).reshape(-1)
        # V-cycles, zero start like the drivers (ThesisProblem.py:97-101)
        for sname, smo in (("wj", None), ("gs", solver.gseidel), ("sor", functools.partial(solver.sor, omega=1.3))):
            for low in (32, 8):
                w = solver.vcycle(np.zeros((n, 1)), f.copy(), H, sm, shift=shift, lowest_level=low, smoother=smo)
                g["%s_vc_%s_%d" % (tag, sname, low)] = np.asarray(w).reshape(-1)
        w = solver.vcycle(x.copy(), f.copy(), H, sm, nu1=2, nu2=3, shift=shift, lowest_level=16, smoother=solver.gseidel)
        g[tag + "_vc_gs_x0"] = np.asarray(w).reshape(-1)
        # fixed-shift inverse iteration (ThesisProblem.py:84-104), 4 cycles, start = f normalised
        v = f / np.linalg.norm(f)
        lam = []
        for _ in range(4):
            w = solver.vcycle(np.zeros((n, 1)), v.copy(), H, sm, shift=shift, lowest_level=32, smoother=solver.gseidel)
            v = np.asarray(w).reshape(-1)
            v = v / np.linalg.norm(v)
            lam.append(np.dot(v.conj().T, H.dot(v)))
        g[tag + "_it_v"] = v
        g[tag + "_it_lam"] = np.array(lam)
    np.savez_compressed(os.path.join(OUT, "multiband_golden.npz"), **g)
    print("wrote", len(g), "arrays;", {k: (np.round(g[k + "_it_lam"].real, 6).tolist(), g[k + "_shift"]) for k in cases})


if __name__ == "__main__":
    main()
