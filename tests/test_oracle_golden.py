"""CPU: the oracle (oracle/mgcmt_oracle.py) against outputs of the REAL reference
(tests/golden/reference_golden.npz, written by oracle/make_golden.py) and against the
known-answer scalars in the reference's UnitTests."""
import numpy as np
import pytest
import scipy.sparse as sp

import mgcmt_oracle as orc

RTOL = 1e-12


def rel(a, b):
    a = np.asarray(a, dtype=float).reshape(-1)
    b = np.asarray(b, dtype=float).reshape(-1)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def o():
    return orc.StencilMaker(), orc.Solver(), orc.Processor()


# ---- known-answer scalars printed in the reference's own tests --------------------------------
def test_known_answers_unit_tests(o, golden):
    sm, s, _ = o
    L = sm.laplacian(16)

    def five(fn):
        x = np.ones(16)
        for _ in range(5):
            x = fn(x, np.zeros(16), L)
        return np.linalg.norm(x)
    # UnitTests/wjacobiTest.py:25, gseidelTest.py:25, sorTest.py:25, vcycleTest.py:25, twogridTest.py:25
    assert abs(five(lambda x, f, A: s.wjacobi(x, f, A, nu=4)) - 2.94959) < 5e-6
    assert abs(five(lambda x, f, A: s.gseidel(x, f, A, nu=4)) - 1.88358) < 5e-6
    assert abs(five(lambda x, f, A: s.sor(x, f, A, nu=4, omega=2. / 3.)) - 2.63327) < 5e-6
    assert abs(np.linalg.norm(s.vcycle(np.ones(16), np.zeros(16), L, sm, nu1=4, nu2=4)) - 0.17756) < 5e-6
    assert abs(np.linalg.norm(s.twogrid(np.ones(16), np.zeros(16), L, sm, 4, 4)) - 0.04979) < 5e-6
    # and to full precision against the reference run in the build container
    assert abs(five(lambda x, f, A: s.wjacobi(x, f, A, nu=4)) - float(golden["ka_wjacobi"])) < 1e-13
    assert abs(five(lambda x, f, A: s.gseidel(x, f, A, nu=4)) - float(golden["ka_gseidel"])) < 1e-13
    assert abs(five(lambda x, f, A: s.sor(x, f, A, nu=4, omega=2. / 3.)) - float(golden["ka_sor"])) < 1e-13


def test_known_answer_triple(o, golden):
    # UnitTests/vcycle_matrixTest.py:21-39; middle value corrected
    sm, s, _ = o
    L4 = sm.laplacian(4)
    want = np.array([-0.382301639189, -0.0257586075778, -0.840838733687])
    for fn in (lambda x, f: s.vcycle(x, f, L4, sm), lambda x, f: s.twogrid(x, f, L4, sm)):
        got = []
        for i in range(3):
            x = fn(np.ones(4) * 4, np.ones(4) * i)
            got.append(np.dot(x, L4.dot(x)))
        assert np.allclose(got, want, rtol=0, atol=5e-12)
        assert np.allclose(got, golden["ka_vcycle_triple"], rtol=1e-12)
    fm = np.zeros((4, 3))
    fm[:, 1] = 1
    fm[:, 2] = 2
    xm = s.vcycle_matrix(np.ones((4, 3)) * 4, fm, L4, sm, shifts=np.zeros(3))
    got = [np.dot(xm[:, j], L4.dot(xm[:, j])) for j in range(3)]
    assert np.allclose(got, golden["ka_vcycle_matrix_triple"], rtol=1e-11)


# ---- operators ----------------------------------------------------------------------------------
def test_operators_match_reference(o, golden):
    sm, _, _ = o
    assert np.array_equal(sm.restriction(16, 8).toarray(), golden["op_R_16_8"])
    assert np.array_equal(sm.interpolation(8, 16).toarray(), golden["op_P_8_16"])
    assert np.array_equal(sm.interpolation(4, 16).toarray(), golden["op_P_4_16"])
    assert np.array_equal(sm.restriction(16, 4).toarray(), golden["op_R_16_4"])
    assert np.array_equal(sm.laplacian(16).toarray(), golden["op_L_16"])
    assert np.array_equal(sm.laplacian(8, "2d").toarray(), golden["op_L2d_8"])
    assert np.array_equal(sm.interpolation(4, 8, "2d").toarray(), golden["op_P2d_4_8"])
    assert np.array_equal(sm.restriction(8, 4, "2d").toarray(), golden["op_R2d_8_4"])
    assert np.array_equal(sm.interpolation(4, 16, "2d").toarray(), golden["op_P2d_4_16"])
    rap = (sm.restriction(16, 8) * sm.laplacian(16) * sm.interpolation(8, 16)).toarray()
    assert np.array_equal(rap, golden["op_RAP_16"])
    rap2 = (sm.restriction(8, 4, "2d") * sm.laplacian(8, "2d") * sm.interpolation(4, 8, "2d")).toarray()
    assert np.array_equal(rap2, golden["op_RAP2d_8"])


def test_operator_error_convention(o, capsys):
    # prints and returns None, never raises (MGCMTStencilMaker.py:44-49,69-74)
    sm, _, _ = o
    assert sm.interpolation(16, 8) is None
    assert sm.interpolation(6, 16) is None
    assert sm.restriction(8, 16) is None
    assert sm.restriction(16, 6) is None
    out = capsys.readouterr().out
    assert "isn't" in out and "bigger" in out


def test_galerkin_structure(o):
    # RAP = rediscretised Laplacian except last diagonal -3 instead of -2
    sm, _, _ = o
    n = 8
    rap = (sm.restriction(n, n // 2) * sm.laplacian(n) * sm.interpolation(n // 2, n)).toarray()
    hc = 1.0 / (n // 2)
    assert np.allclose(np.diag(rap)[:-1] * hc ** 2, -2.0)
    assert np.isclose(rap[-1, -1] * hc ** 2, -3.0)


# ---- smoothers ---------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,n,dim", [("1d64", 64, "1d"), ("1d64s", 64, "1d"), ("2d16", 16, "2d"), ("2d16s", 16, "2d")])
def test_smoothers_match_reference(o, golden, tag, n, dim):
    sm, s, _ = o
    nn = n if dim == "1d" else n * n
    H = (-1. / np.pi ** 2) * sm.laplacian(n, dim)
    A = H - sp.eye(nn) * float(golden["sm_%s_shift" % tag])
    v0 = golden["sm_%s_v0" % tag]
    f = golden["sm_%s_f" % tag]
    assert rel(s.wjacobi(v0.copy(), f.copy(), A, nu=3), golden["sm_%s_wjacobi" % tag]) < RTOL
    assert rel(s.wjacobi(v0.copy(), f.copy(), A, nu=2, omega=0.8), golden["sm_%s_wjacobi_w08" % tag]) < RTOL
    assert rel(s.gseidel(v0.copy(), f.copy(), A, nu=3), golden["sm_%s_gseidel" % tag]) < RTOL
    assert rel(s.sor(v0.copy(), f.copy(), A, nu=3, omega=1.3), golden["sm_%s_sor" % tag]) < RTOL


# ---- V-cycles ----------------------------------------------------------------------------------
VC_TAGS = ["1d64", "1d64s", "1d256s_l8", "1d64_gs", "2d16_l8", "2d32_l8", "2d32_l2", "2d32_l8_nu",
           "2d64_l8", "2d16_l8_gs"]


@pytest.mark.parametrize("tag", VC_TAGS)
def test_vcycle_matches_reference(o, golden, tag):
    sm, s, _ = o
    n, dim, shift, lowest, nu1, nu2 = golden["vc_%s_meta" % tag]
    n, lowest, nu1, nu2 = int(n), int(lowest), int(nu1), int(nu2)
    dim = "1d" if dim == 1 else "2d"
    H = (-1. / np.pi ** 2) * sm.laplacian(n, dim)
    kw = {"smoother": s.gseidel} if tag.endswith("_gs") else {}
    out = s.vcycle(golden["vc_%s_v0" % tag].copy(), golden["vc_%s_f" % tag].copy(), H, sm, nu1=nu1, nu2=nu2,
                   shift=shift, lowest_level=lowest, dimension=dim, **kw)
    assert out.shape == (n if dim == "1d" else n * n,)
    # the coarsest exact solve (SuperLU on an indefinite shifted operator) limits agreement
    assert rel(out, golden["vc_%s_out" % tag]) < 1e-11


def test_vcycle_coarsest_returns_column(o):
    # quirk Q7 (MGCMTSolver.py:305-308): called directly at the coarsest size -> shape (n,1)
    sm, s, _ = o
    L = sm.laplacian(2)
    assert s.vcycle(np.ones(2), np.ones(2), L, sm).shape == (2, 1)


def test_twogrid_matches_reference(o, golden):
    sm, s, _ = o
    H = (-1. / np.pi ** 2) * sm.laplacian(64)
    out = s.twogrid(golden["tg_1d64_v0"].copy(), golden["tg_1d64_f"].copy(), H, sm, nu1=3, nu2=2, shift=3.9)
    assert rel(out, golden["tg_1d64_out"]) < 1e-11


@pytest.mark.parametrize("tag", ["1d64", "2d16"])
def test_vcycle_matrix_matches_reference(o, golden, tag):
    sm, s, _ = o
    n, dim, lowest = golden["vm_%s_meta" % tag]
    dim = "1d" if dim == 1 else "2d"
    H = (-1. / np.pi ** 2) * sm.laplacian(int(n), dim)
    out = s.vcycle_matrix(golden["vm_%s_v0" % tag].copy(), golden["vm_%s_f" % tag].copy(), H, sm,
                          shifts=golden["vm_%s_shifts" % tag], lowest_level=int(lowest), dimension=dim)
    assert rel(out, golden["vm_%s_out" % tag]) < 1e-10


# ---- RQ minimisation ---------------------------------------------------------------------------
def test_rqmin_matches_reference(o, golden):
    sm, s, _ = o
    n = 32
    H = sp.csr_matrix((-1. / np.pi ** 2) * sm.laplacian(n))
    M = sp.eye(n, format="csr")
    x, rho = s.rqmin(H, golden["rq_x0"].copy(), M, nu=4)
    assert rel(x, golden["rq_rqmin_x"]) < 1e-10 and abs(rho - float(golden["rq_rqmin_rho"])) < 1e-11
    x, rho = s.vcycle_rqmg(golden["rq_x0"].copy(), H, M)
    assert rel(x, golden["rq_rqmg_x"]) < 1e-9 and abs(rho - float(golden["rq_rqmg_rho"])) < 1e-10


# ---- Gram-Schmidt ------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["ill", "well", "rnd"])
def test_gramschmidt_matches_reference(o, golden, tag):
    _, _, p = o
    a = golden["gs_%s_in" % tag]
    assert np.allclose(p.gramschmidt(a.copy()), golden["gs_%s_mgs" % tag], rtol=0, atol=1e-14)
    assert np.allclose(p.gramschmidt(a.copy(), modified=0), golden["gs_%s_cgs" % tag], rtol=0, atol=1e-14)
    assert np.allclose(p.normalize(a.copy()), golden["gs_%s_norm" % tag], rtol=0, atol=1e-15)


def test_gramschmidt_unit_test_values(o):
    # UnitTests/GramSchmidt.py: CGS loses orthogonality on the ill-conditioned set, MGS keeps it
    _, _, p = o
    eps = 1e-8
    ill = np.array([[1, 1, 1], [eps, eps, 0], [eps, 0, eps]], dtype=float)
    c = p.gramschmidt(ill.copy(), modified=0)
    m = p.gramschmidt(ill.copy(), modified=1)
    assert abs(abs(np.dot(c[:, 1], c[:, 2])) - 0.7071067811865475) < 1e-6
    assert abs(np.dot(m[:, 1], m[:, 2])) < 1e-12
    well = np.array([[1, 1, 1], [2, 1, 0], [5, 1, 3]], dtype=float)
    q = p.gramschmidt(well.copy())
    assert np.allclose(q[:, 0], [0.18257419, 0.36514837, 0.91287093], atol=1e-8)
    assert np.allclose(q[:, 1], [0.78772636, 0.50128041, -0.35805744], atol=1e-8)
    assert np.allclose(q[:, 2], [0.58834841, -0.78446454, 0.19611614], atol=1e-8)


# ---- shift-method outer loop -------------------------------------------------------------------
def test_shift_method_loop_matches_reference(o, golden):
    sm, s, p = o
    N, N0, iters, lowest = [int(x) for x in golden["sh_meta"]]
    H = (-1. / np.pi ** 2) * sm.laplacian(N, "2d")
    V = golden["sh_V0"].copy()
    shifts = golden["sh_shifts"]
    lam = np.zeros((iters, 4))
    for it in range(iters):
        for c in range(4):
            w = s.vcycle(np.zeros(N * N), V[:, c].copy(), H, sm, shift=shifts[c], dimension="2d", lowest_level=lowest)
            V[:, c] = w / np.linalg.norm(w)
            lam[it, c] = np.dot(V[:, c], H.dot(V[:, c]))
        V = p.gramschmidt(V)
    assert np.allclose(lam, golden["sh_lambda"], rtol=1e-11, atol=0)
    # degenerate (1,2)/(2,1) pair: compare the invariant subspace, and the simple vectors directly
    assert rel(V[:, 0], golden["sh_V"][:, 0]) < 1e-9
    assert rel(V[:, 3], golden["sh_V"][:, 3]) < 1e-9


def test_closed_form_spectrum(o):
    # SURVEY.md section 4 "extra oracles"; report p.43: 2D n=128 -> 1.96901511, 4.92195391
    sm, _, _ = o
    assert abs(orc.well_eigenvalue_2d(128, 1, 1) - 1.96901511) < 5e-9
    assert abs(orc.well_eigenvalue_2d(128, 1, 2) - 4.92195391) < 5e-9
    assert abs(orc.well_eigenvalue_2d(16, 1, 1) - 1.76659015) < 5e-9
    H = (-1. / np.pi ** 2) * sm.laplacian(16, "2d")
    v = orc.well_eigenvector_2d(16, 1, 2)
    assert np.linalg.norm(H @ v - orc.well_eigenvalue_2d(16, 1, 2) * v) < 1e-12


# ---------------------------------------------------------------------------------------------------
# multiband (complex, 4 coupled bands) Hamiltonians: ThesisProblem.py / PotWellSolver.makeMatrix
MB_TAGS = ("z0", "z7", "x7")


def _mb(multiband, tag):
    import scipy.sparse as sp
    H = sp.csc_matrix(multiband[tag + "_H"])
    return H, H.shape[0], float(multiband[tag + "_shift"]), multiband[tag + "_x"], multiband[tag + "_f"]


def _crel(a, b):
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("tag", MB_TAGS)
def test_multiband_smoothers_match_reference(o, multiband, tag):
    import scipy.sparse as sp
    _, s, _ = o
    H, n, shift, x, f = _mb(multiband, tag)
    As = (H - sp.eye(n) * shift).tocsc()
    assert _crel(s.wjacobi(x.copy(), f.copy(), As, nu=3), multiband[tag + "_wj"]) < 1e-13
    assert _crel(s.gseidel(x.copy(), f.copy(), As, nu=3), multiband[tag + "_gs"]) < 1e-13
    assert _crel(s.sor(x.copy(), f.copy(), As, nu=3, omega=1.3), multiband[tag + "_sor"]) < 1e-13


@pytest.mark.parametrize("tag", MB_TAGS)
def test_multiband_vcycles_match_reference(o, multiband, tag):
    import functools
    sm, s, _ = o
    H, n, shift, x, f = _mb(multiband, tag)
    for sname, smo in (("wj", None), ("gs", s.gseidel), ("sor", functools.partial(s.sor, omega=1.3))):
        for low in (32, 8):
            w = s.vcycle(np.zeros((n, 1)), f.copy(), H, sm, shift=shift, lowest_level=low, smoother=smo)
            assert _crel(w, multiband["%s_vc_%s_%d" % (tag, sname, low)]) < 1e-12, (sname, low)
    w = s.vcycle(x.copy(), f.copy(), H, sm, nu1=2, nu2=3, shift=shift, lowest_level=16, smoother=s.gseidel)
    assert _crel(w, multiband[tag + "_vc_gs_x0"]) < 1e-12


@pytest.mark.parametrize("tag", MB_TAGS)
def test_multiband_shift_iteration_matches_reference(o, multiband, tag):
    """ThesisProblem.py:84-104: fixed-shift inverse iteration with V-cycles, Rayleigh quotient after each."""
    sm, s, _ = o
    H, n, shift, _, f = _mb(multiband, tag)
    v = f / np.linalg.norm(f)
    lam = []
    for _ in range(4):
        w = s.vcycle(np.zeros((n, 1)), v.copy(), H, sm, shift=shift, lowest_level=32, smoother=s.gseidel)
        v = np.asarray(w).reshape(-1)
        v = v / np.linalg.norm(v)
        lam.append(np.dot(v.conj(), H.dot(v)))
    assert np.max(np.abs(np.array(lam) - multiband[tag + "_it_lam"])) < 1e-11
    assert _crel(v, multiband[tag + "_it_v"]) < 1e-10


# ---------------------------------------------------------------------------------------------------
# BASELINE config 0 at its own size: 1-D well, 1024 points, shift method with Gauss-Seidel V-cycles
def test_config0_shift_loop_matches_reference(o, config0):
    import functools
    sm, s, p = o
    n, n0, low, iters, k = [int(x) for x in config0["meta"]]
    H = (-1. / np.pi ** 2) * sm.laplacian(n)
    V = config0["V0"].copy()
    lam = np.zeros((iters, k))
    for it in range(iters):
        for j in range(k):
            w = s.vcycle(np.zeros((n, 1)), V[:, j].copy(), H, sm, shift=config0["shifts"][j], lowest_level=low,
                         smoother=s.gseidel)
            V[:, j] = w / np.linalg.norm(w)
            lam[it, j] = V[:, j] @ (H @ V[:, j])
        V = p.gramschmidt(V)
    assert np.max(np.abs(lam - config0["lam"])) < 1e-12
    assert np.linalg.norm(V - config0["V"]) < 1e-12 * np.linalg.norm(config0["V"])
    f = config0["f"]
    w = s.vcycle(np.zeros((n, 1)), f.copy(), H, sm, shift=config0["shifts"][0], lowest_level=low)
    assert np.linalg.norm(w - config0["vc_wj"]) < 1e-13 * np.linalg.norm(w)
    w = s.vcycle(np.zeros((n, 1)), f.copy(), H, sm, shift=config0["shifts"][0], lowest_level=low,
                 smoother=functools.partial(s.sor, omega=1.2))
    assert np.linalg.norm(w - config0["vc_sor"]) < 1e-13 * np.linalg.norm(w)
    # and the loop does what the driver wants: the lowest eigenvalues of the 1024-point well
    assert np.all(np.abs(lam[-1] - [orc.well_eigenvalue_1d(n, j + 1) for j in range(k)]) < 1e-4)
