"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/config0_golden.npz from the REAL reference.

Run in the build container (needs /root/reference):   python oracle/make_golden_config0.py
BASELINE config 0: the 1-D infinite well on a 1024-point grid, lowest eigenpairs by the shift method with V-cycles
and the Gauss-Seidel smoother (the loop of 1DPotMGS.py:60-75 / 1DPotGS.py:66-80 at the size BASELINE names), run through
the unmodified reference classes (oracle/ref_loader.py).  The start vectors replace the drivers' coarse-grid `eigsh` by the
closed-form coarse eigenvectors, interpolated with the reference's own interpolation matrix.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    warnings.simplefilter("ignore")
    SM, S, P = ref_loader.load_reference()
    sm, solver, proc = SM(), S(), P()
    n, n0, low, iters, k = 1024, 16, 8, 3, 2
    H = (-1. / np.pi ** 2) * sm.laplacian(n)

    def ev1(m, j):
        return (4. * m * m / np.pi ** 2) * np.sin(j * np.pi / (2. * (m + 1))) ** 2

    def vec1(m, j):
        v = np.sin(j * np.pi * (np.arange(m) + 1.) / (m + 1.))
        return v / np.linalg.norm(v)
    shifts = np.array([ev1(n0, j + 1) for j in range(k)])
    Pm = sm.interpolation(n0, n)
    V = np.zeros((n, k))
    for j in range(k):
        V[:, j] = Pm * vec1(n0, j + 1)
        V[:, j] /= np.linalg.norm(V[:, j])
    g = {"meta": np.array([n, n0, low, iters, k], dtype=float), "shifts": shifts, "V0": V.copy()}
    lam = np.zeros((iters, k))
    for it in range(iters):
        for j in range(k):
            w = solver.vcycle(np.zeros((n, 1)), np.array(V[:, j]), H, sm, shift=shifts[j], lowest_level=low,
                              smoother=solver.gseidel)
            V[:, j] = w / np.linalg.norm(w)
            lam[it, j] = float(np.dot(V[:, j], H.dot(V[:, j])))
        V = proc.gramschmidt(V)
    g["V"] = V
    g["lam"] = lam
    # one weighted-Jacobi and one SOR cycle on a seeded right-hand side at the same size
    f = np.random.RandomState(0).random_sample(n)
    g["f"] = f
    g["vc_wj"] = np.asarray(solver.vcycle(np.zeros((n, 1)), f.copy(), H, sm, shift=shifts[0], lowest_level=low)).reshape(-1)
    import functools
    g["vc_sor"] = np.asarray(solver.vcycle(np.zeros((n, 1)), f.copy(), H, sm, shift=shifts[0], lowest_level=low,
                                           smoother=functools.partial(solver.sor, omega=1.2))).reshape(-1)
    np.savez_compressed(os.path.join(OUT, "config0_golden.npz"), **g)
    print("wrote", len(g), "arrays; eigenvalues per iteration:", lam.tolist(), "exact:", [ev1(n, j + 1) for j in range(k)])


if __name__ == "__main__":
    main()
