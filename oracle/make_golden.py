"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the REAL reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
It executes the unmodified reference classes (through oracle/ref_loader.py's in-memory,
syntax-only Python-3 translation) on seeded inputs and records inputs + outputs.  The fixtures
travel to the GPU box; this script and /root/reference do not need to.

Every case stores its inputs, so a test never has to regenerate them.
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


def rand(n, seed):
    return np.random.RandomState(seed).random_sample(n)


def main():
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    SM, S, P = ref_loader.load_reference()
    sm, solver, proc = SM(), S(), P()
    g = {}

    # ---- operators (UnitTests/operatorTest.py:17-37) --------------------------------------
    g["op_R_16_8"] = sm.restriction(16, 8).toarray()
    g["op_P_8_16"] = sm.interpolation(8, 16).toarray()
    g["op_P_4_16"] = sm.interpolation(4, 16).toarray()          # multi-level jump
    g["op_R_16_4"] = sm.restriction(16, 4).toarray()
    L16 = sm.laplacian(16)
    g["op_L_16"] = L16.toarray()
    g["op_RAP_16"] = (sm.restriction(16, 8) * L16 * sm.interpolation(8, 16)).toarray()
    L8_2d = sm.laplacian(8, dimension="2d")
    g["op_L2d_8"] = L8_2d.toarray()
    g["op_P2d_4_8"] = sm.interpolation(4, 8, dimension="2d").toarray()
    g["op_R2d_8_4"] = sm.restriction(8, 4, dimension="2d").toarray()
    g["op_RAP2d_8"] = (sm.restriction(8, 4, dimension="2d") * L8_2d
                       * sm.interpolation(4, 8, dimension="2d")).toarray()
    g["op_P2d_4_16"] = sm.interpolation(4, 16, dimension="2d").toarray()

    # ---- scalar known answers (UnitTests/{wjacobi,gseidel,sor,vcycle,twogrid}Test.py) -------
    def five(fn):
        x = np.ones(16); f = np.zeros(16)
        for _ in range(5):
            x = fn(x, f, L16)
        return float(np.linalg.norm(x))
    g["ka_wjacobi"] = five(lambda x, f, A: solver.wjacobi(x, f, A, nu=4))
    g["ka_gseidel"] = five(lambda x, f, A: solver.gseidel(np.reshape(x, (16, 1)), np.reshape(f, (16, 1)), A, nu=4))
    g["ka_sor"] = five(lambda x, f, A: solver.sor(np.reshape(x, (16, 1)), np.reshape(f, (16, 1)), A, nu=4, omega=2. / 3.))
    g["ka_vcycle"] = float(np.linalg.norm(solver.vcycle(np.ones(16), np.zeros(16), L16, sm, nu1=4, nu2=4)))
    g["ka_twogrid"] = float(np.linalg.norm(solver.twogrid(np.ones(16), np.zeros(16), L16, sm, 4, 4)))
    L4 = sm.laplacian(4)
    trip_v, trip_t = [], []
    for i in range(3):
        x = solver.vcycle(np.ones(4) * 4, np.ones(4) * i, L4, sm)
        trip_v.append(float(np.dot(x, L4.dot(x))))
        x = solver.twogrid(np.ones(4) * 4, np.ones(4) * i, L4, sm)
        trip_t.append(float(np.dot(x, L4.dot(x))))
    g["ka_vcycle_triple"] = np.array(trip_v)
    g["ka_twogrid_triple"] = np.array(trip_t)
    fm = np.zeros((4, 3))
    for i in range(3):
        fm[:, i] = i
    xm = solver.vcycle_matrix(np.ones((4, 3)) * 4, fm, L4, sm, shifts=np.zeros(3))
    g["ka_vcycle_matrix_triple"] = np.array([float(np.dot(xm[:, j], L4.dot(xm[:, j]))) for j in range(3)])

    # ---- smoothers on seeded inputs -----------------------------------------------------------
    for tag, n, dim, shift in (("1d64", 64, "1d", 0.0), ("1d64s", 64, "1d", 3.3),
                               ("2d16", 16, "2d", 0.0), ("2d16s", 16, "2d", 4.38639582)):
        nn = n if dim == "1d" else n * n
        H = (-1. / np.pi ** 2) * sm.laplacian(n, dimension=dim)
        from scipy import sparse
        A = H - sparse.eye(nn) * shift
        v0 = rand(nn, 1); f = rand(nn, 2)
        g["sm_%s_v0" % tag] = v0; g["sm_%s_f" % tag] = f; g["sm_%s_shift" % tag] = shift
        g["sm_%s_wjacobi" % tag] = solver.wjacobi(np.array(v0), np.array(f), A, nu=3)[:, 0]
        g["sm_%s_wjacobi_w08" % tag] = solver.wjacobi(np.array(v0), np.array(f), A, nu=2, omega=0.8)[:, 0]
        g["sm_%s_gseidel" % tag] = np.asarray(solver.gseidel(v0.reshape(-1, 1), f.reshape(-1, 1), A, nu=3))[:, 0]
        g["sm_%s_sor" % tag] = np.asarray(solver.sor(v0.reshape(-1, 1), f.reshape(-1, 1), A, nu=3, omega=1.3))[:, 0]

    # ---- V-cycles --------------------------------------------------------------------------
    def vc_case(tag, n, dim, shift, lowest, nu1=4, nu2=4, smoother=None, zero_v0=False):
        nn = n if dim == "1d" else n * n
        H = (-1. / np.pi ** 2) * sm.laplacian(n, dimension=dim)
        v0 = np.zeros(nn) if zero_v0 else rand(nn, 3)
        f = rand(nn, 4)
        g["vc_%s_v0" % tag] = np.array(v0); g["vc_%s_f" % tag] = np.array(f)
        g["vc_%s_meta" % tag] = np.array([n, 1 if dim == "1d" else 2, shift, lowest, nu1, nu2], dtype=float)
        kw = {}
        if smoother is not None:
            kw["smoother"] = getattr(solver, smoother)
        out = solver.vcycle(np.array(v0), np.array(f), H, sm, nu1=nu1, nu2=nu2, shift=shift,
                            lowest_level=lowest, dimension=dim, **kw)
        g["vc_%s_out" % tag] = np.asarray(out).reshape(-1)

    vc_case("1d64", 64, "1d", 0.0, 2)
    vc_case("1d64s", 64, "1d", 3.9, 2)
    vc_case("1d256s_l8", 256, "1d", 8.9, 8, nu1=2, nu2=3)
    vc_case("1d64_gs", 64, "1d", 3.9, 4, smoother="gseidel")
    vc_case("2d16_l8", 16, "2d", 1.76659015, 8, zero_v0=True)
    vc_case("2d32_l8", 32, "2d", 4.38639582, 8, zero_v0=True)
    vc_case("2d32_l2", 32, "2d", 0.0, 2)
    vc_case("2d32_l8_nu", 32, "2d", 7.00620149, 8, nu1=2, nu2=1)
    vc_case("2d64_l8", 64, "2d", 4.38639582, 8, zero_v0=True)
    vc_case("2d16_l8_gs", 16, "2d", 1.76659015, 8, smoother="gseidel", zero_v0=True)

    # twogrid 1-D
    H = (-1. / np.pi ** 2) * sm.laplacian(64)
    v0 = rand(64, 5); f = rand(64, 6)
    g["tg_1d64_v0"] = np.array(v0); g["tg_1d64_f"] = np.array(f)
    g["tg_1d64_out"] = solver.twogrid(np.array(v0), np.array(f), H, sm, nu1=3, nu2=2, shift=3.9)

    # vcycle_matrix 1-D and 2-D with per-column shifts
    for tag, n, dim, lowest, shifts in (("1d64", 64, "1d", 2, np.array([0.9, 3.9, 8.8])),
                                        ("2d16", 16, "2d", 8, np.array([1.76659015, 4.38639582, 4.38639582, 7.00620149]))):
        nn = n if dim == "1d" else n * n
        H = (-1. / np.pi ** 2) * sm.laplacian(n, dimension=dim)
        k = len(shifts)
        V0 = rand(nn * k, 7).reshape(nn, k); F = rand(nn * k, 8).reshape(nn, k)
        g["vm_%s_v0" % tag] = np.array(V0); g["vm_%s_f" % tag] = np.array(F); g["vm_%s_shifts" % tag] = shifts
        g["vm_%s_meta" % tag] = np.array([n, 1 if dim == "1d" else 2, lowest], dtype=float)
        g["vm_%s_out" % tag] = solver.vcycle_matrix(np.array(V0), np.array(F), H, sm, shifts=shifts,
                                                    lowest_level=lowest, dimension=dim)

    # ---- Rayleigh-quotient minimisation (RQMin.py:28-49 pattern) -----------------------------
    from scipy import sparse
    n = 32
    H = sparse.csr_matrix((-1. / np.pi ** 2) * sm.laplacian(n))
    M = sparse.eye(n, format="csr")
    x0 = rand(n, 0)
    g["rq_x0"] = np.array(x0)
    x, rho = solver.rqmin(H, np.array(x0), M, nu=4)
    g["rq_rqmin_x"] = x; g["rq_rqmin_rho"] = float(rho)
    x, rho = solver.vcycle_rqmg(np.array(x0), H, M)
    g["rq_rqmg_x"] = x; g["rq_rqmg_rho"] = float(rho)

    # ---- Gram-Schmidt (UnitTests/GramSchmidt.py) ---------------------------------------------
    eps = 1e-8
    ill = np.array([[1, 1, 1], [eps, eps, 0], [eps, 0, eps]], dtype=float)
    well = np.array([[1, 1, 1], [2, 1, 0], [5, 1, 3]], dtype=float)
    rnd = rand(40 * 5, 9).reshape(40, 5)
    for tag, mat in (("ill", ill), ("well", well), ("rnd", rnd)):
        g["gs_%s_in" % tag] = mat
        g["gs_%s_mgs" % tag] = proc.gramschmidt(np.array(mat))
        g["gs_%s_cgs" % tag] = proc.gramschmidt(np.array(mat), modified=0)
        g["gs_%s_norm" % tag] = proc.normalize(np.array(mat))

    # ---- shift-method outer loop, 2D (2DPotGS.py:79-105 pattern), closed-form start ----------
    N, N0, iters = 32, 16, 3
    H = (-1. / np.pi ** 2) * sm.laplacian(N, dimension="2d")
    modes = [(1, 1), (1, 2), (2, 1), (2, 2)]

    def ev1(n, k):
        return (4. * n * n / np.pi ** 2) * np.sin(k * np.pi / (2. * (n + 1))) ** 2

    def vec1(n, k):
        v = np.sin(k * np.pi * (np.arange(n) + 1.) / (n + 1.))
        return v / np.linalg.norm(v)
    shifts = np.array([ev1(N0, a) + ev1(N0, b) for a, b in modes])
    P2 = sm.interpolation(N0, N, dimension="2d")
    V = np.zeros((N * N, 4))
    for c, (a, b) in enumerate(modes):
        V[:, c] = P2 * np.kron(vec1(N0, a), vec1(N0, b))
        V[:, c] /= np.linalg.norm(V[:, c])
    g["sh_V0"] = np.array(V); g["sh_shifts"] = shifts
    g["sh_meta"] = np.array([N, N0, iters, 8], dtype=float)
    lam = np.zeros((iters, 4))
    for it in range(iters):
        for c in range(4):
            w = solver.vcycle(np.zeros((N * N, 1)), np.array(V[:, c]), H, sm, shift=shifts[c],
                              dimension="2d", lowest_level=8)
            V[:, c] = w / np.linalg.norm(w)
            lam[it, c] = float(np.dot(V[:, c], H.dot(V[:, c])))
        V = proc.gramschmidt(V)
    g["sh_V"] = V; g["sh_lambda"] = lam

    np.savez_compressed(os.path.join(OUT, "reference_golden.npz"), **g)
    print("wrote %d arrays to %s" % (len(g), os.path.join(OUT, "reference_golden.npz")))
    for k in sorted(g):
        if k.startswith("ka_"):
            print(k, g[k])


if __name__ == "__main__":
    main()
