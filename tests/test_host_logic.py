"""CPU: host-side logic of the drop-in (no GPU compute): operator recognition, the stencil-maker
mirror, the C-ABI library's symbols, and the loud-failure contract."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import mgcmt_oracle as orc
from multigridcmt_b200 import MGCMTStencilMaker, SeparableOperator, UnsupportedOperator
from multigridcmt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same(a, b):
    return (sp.csc_matrix(a) != sp.csc_matrix(b)).nnz == 0


def test_stencil_maker_mirror_equals_oracle_and_reference(golden):
    sm, osm = MGCMTStencilMaker(), orc.StencilMaker()
    for a, b in [(8, 16), (4, 16), (2, 4), (16, 64)]:
        assert same(sm.interpolation(a, b), osm.interpolation(a, b))
        assert same(sm.restriction(b, a), osm.restriction(b, a))
    assert same(sm.interpolation(4, 8, "2d"), osm.interpolation(4, 8, "2d"))
    assert same(sm.restriction(8, 4, "2d"), osm.restriction(8, 4, "2d"))
    assert np.array_equal(sm.restriction(16, 8).toarray(), golden["op_R_16_8"])
    assert np.array_equal(sm.interpolation(8, 16).toarray(), golden["op_P_8_16"])
    assert np.array_equal(sm.interpolation(4, 16).toarray(), golden["op_P_4_16"])
    assert np.array_equal(sm.laplacian(16).toarray(), golden["op_L_16"])
    assert np.array_equal(sm.laplacian(8, "2d").toarray(), golden["op_L2d_8"])
    assert np.array_equal(sm.interpolation(4, 8, "2d").toarray(), golden["op_P2d_4_8"])
    assert np.array_equal(sm.restriction(8, 4, "2d").toarray(), golden["op_R2d_8_4"])
    assert sm.laplacian(8).format == "csc" and sm.interpolation(4, 8).format == "csc"


def test_stencil_maker_error_convention(capsys):
    sm = MGCMTStencilMaker()
    assert sm.interpolation(16, 8) is None
    assert sm.interpolation(6, 16) is None
    assert sm.interpolation(4, 12) is None
    assert sm.restriction(8, 16) is None
    assert sm.restriction(16, 6) is None
    out = capsys.readouterr().out
    assert out.count("!") == 5


@pytest.mark.parametrize("dim,n", [("1d", 32), ("2d", 16)])
def test_recognise_well_operator(dim, n):
    sm = MGCMTStencilMaker()
    H = (-1. / np.pi ** 2) * sm.laplacian(n, dim)
    op = SeparableOperator.from_sparse(H, dim)
    assert abs(op.tocsc() - H).max() == 0.0
    nn = n if dim == "1d" else n * n
    Hs = H - sp.eye(nn) * 4.386
    ops = SeparableOperator.from_sparse(Hs, dim)
    assert abs(ops.tocsc() - Hs).max() < 1e-13
    # matrix-free constructor gives the same operator as recognising the scipy matrix
    mf = (-1. / np.pi ** 2) * sm.laplacian(n, dim, matrix_free=True)
    assert abs(mf.tocsc() - H).max() < 1e-12
    assert np.allclose(mf.diagonal(), H.diagonal())


def test_recognise_separable_potential():
    n = 8
    sm = MGCMTStencilMaker()
    L = sm.laplacian(n, "2d")
    vx = np.linspace(0, 3, n)
    V = (vx[:, None] ** 2 + 2 * vx[None, :]).reshape(-1)
    H = -L + sp.diags(V)
    op = SeparableOperator.from_sparse(H, "2d")
    assert abs(op.tocsc() - H).max() < 1e-12


def test_refuses_non_separable():
    n = 8
    sm = MGCMTStencilMaker()
    L = sm.laplacian(n, "2d")
    rng = np.random.RandomState(0)
    with pytest.raises(UnsupportedOperator):
        SeparableOperator.from_sparse(-L + sp.diags(rng.random_sample(n * n)), "2d")
    with pytest.raises(UnsupportedOperator):
        SeparableOperator.from_sparse(sp.random(64, 64, 0.2, format="csc", random_state=1), "2d")
    with pytest.raises(UnsupportedOperator):
        SeparableOperator.from_sparse(sp.random(64, 64, 0.2, format="csc", random_state=1), "1d")
    with pytest.raises(UnsupportedOperator):
        SeparableOperator.from_sparse(sp.eye(16, format="csc") * (1 + 1j), "1d")


def test_library_exports_every_header_symbol():
    """include/mgcmt_b200.h <-> libmgcmt_b200.so <-> the ctypes table, symbol by symbol."""
    hdr = open(os.path.join(ROOT, "include", "mgcmt_b200.h")).read()
    declared = set(re.findall(r"\b(mgcmt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mgcmt_hier"}  # the opaque struct tag
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.load()  # loading needs no GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mgcmt_abi_version() == 1


def test_bad_arguments_are_errors_not_crashes():
    lib = _lib.load()
    handle = ctypes.c_void_p()
    z = np.zeros(4)
    p = z.ctypes.data_as(ctypes.c_void_p)
    rc = lib.mgcmt_hier_create(ctypes.byref(handle), 1, 6, 0, p, p, p, p, p, p, 2, None)
    assert rc == 1 and b"power of two" in lib.mgcmt_last_error()
    rc = lib.mgcmt_hier_create(ctypes.byref(handle), 4, 4, 1, None, p, p, p, p, p, 2, None)
    assert rc == 1
    assert lib.mgcmt_vcycle(None, 0.0, 4, 4, 0, 0.66, None, None, 0, None) == 1
    assert lib.mgcmt_dot(-1, None, None, None, None) == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multigridcmt_b200 import MGCMTSolver
    sm, s = MGCMTStencilMaker(), MGCMTSolver()
    with pytest.raises(_lib.MgcmtError):
        s.vcycle(np.ones(16), np.zeros(16), sm.laplacian(16), sm)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multigridcmt_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+(mgcmt_oracle|ref_loader|oracle)\b", src, re.M), fn
                assert "oracle/_ref" not in src and "/root/reference" not in src, fn


def test_reference_arm_json_contract():
    """`bench.py --impl reference` (CPU only: the C/OpenMP port of the path) prints one JSON line with the contract keys"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--grid", "512"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


# ---------------------------------------------------------------------------------------------------
# banded (multiband complex) operators: host-side storage and routing
def test_banded_operator_round_trip(multiband):
    import scipy.sparse as sp
    from multigridcmt_b200.banded import BandedOperator
    for tag, ndiag, real in (("z0", 3, True), ("z7", 9, False), ("x7", 13, False)):
        H = sp.csc_matrix(multiband[tag + "_H"])
        op = BandedOperator.from_sparse(H)
        assert len(op.offsets) == ndiag and op.is_real == real and 0 in op.offsets
        assert np.all(np.diff(op.offsets) > 0)
        assert abs(op.tocsc() - H).max() == 0
        assert np.array_equal(op.diagonal(), H.diagonal())
        # entries that fall outside the matrix are stored as zeros
        for k, off in enumerate(op.offsets):
            i = np.arange(op.n)
            assert not np.any(op.vals[k, (i + off < 0) | (i + off >= op.n)])


def test_banded_operator_limits():
    import scipy.sparse as sp
    from multigridcmt_b200.banded import BandedOperator
    from multigridcmt_b200.operators import UnsupportedOperator
    dense = sp.csc_matrix(np.ones((128, 128)))
    with pytest.raises(UnsupportedOperator):
        BandedOperator.from_sparse(dense)
    with pytest.raises(UnsupportedOperator):
        BandedOperator.from_sparse(sp.csc_matrix(np.ones((4, 5))))
    # a matrix without a stored main diagonal still gets offset 0
    op = BandedOperator.from_sparse(sp.diags([np.ones(7)], [1], shape=(8, 8), format="csc"))
    assert list(op.offsets) == [0, 1]


def test_routing_picks_the_banded_path(multiband):
    """real separable stencils -> fused path; complex / wide 1-D operators and complex vectors -> banded path"""
    import scipy.sparse as sp
    from multigridcmt_b200 import MGCMTSolver, MGCMTStencilMaker
    from multigridcmt_b200.banded import BandedOperator
    from multigridcmt_b200.operators import SeparableOperator, UnsupportedOperator
    route = MGCMTSolver._route
    L = MGCMTStencilMaker().laplacian(64)
    x = np.zeros(64)
    assert isinstance(route(L, "1d", x, x), SeparableOperator)
    assert isinstance(route(L, "1d", x, x.astype(complex)), BandedOperator)
    assert isinstance(route(sp.csc_matrix(multiband["z7_H"]), "1d", x, x), BandedOperator)
    assert isinstance(route(sp.csc_matrix(multiband["z0_H"]), "1d", x, x), BandedOperator)   # complex dtype
    penta = sp.diags([np.ones(62), np.ones(64), np.ones(62)], [-2, 0, 2], format="csc")
    assert isinstance(route(penta, "1d", x, x), BandedOperator)
    L2 = MGCMTStencilMaker().laplacian(8, dimension="2d")
    with pytest.raises(UnsupportedOperator):
        route(L2, "2d", np.zeros(64), np.zeros(64, dtype=complex))
    assert MGCMTSolver._guess_dimension(sp.csc_matrix(multiband["z7_H"]), 256) == "1d"


# ---------------------------------------------------------------------------------------------------
# numpy twins of the banded device formulas (csrc/band.cu), pinned to the REAL reference's outputs on CPU:
# a change of the formulas has to pass here before it goes to the GPU
def _galerkin_twin(op):
    """band_galerkin_kernel + the coarse-offset rule of mgcmt_band_create, restated with numpy"""
    from multigridcmt_b200.banded import BandedOperator
    n, offs = op.n, [int(o) for o in op.offsets]
    m = n // 2
    cs = sorted({D for d in offs for D in range((d - 1) // 2, (d + 2) // 2 + 1) if -m < D < m})
    vals = np.zeros((len(cs), m), dtype=complex)
    for kc, D in enumerate(cs):
        for J in range(m):
            K = J + D
            if not 0 <= K < m:
                continue
            acc = 0.0
            for a in range(3):
                ia = 2 * J + a
                if ia >= n:
                    continue
                for b in range(3):
                    ib = 2 * K + b
                    d = 2 * D + b - a
                    if ib >= n or d not in offs:
                        continue
                    acc += (0.5 if a == 1 else 0.25) * (1.0 if b == 1 else 0.5) * op.vals[offs.index(d), ia]
            vals[kc, J] = acc
    return BandedOperator(m, cs, vals)


@pytest.mark.parametrize("tag", ["z0", "z7", "x7"])
def test_banded_galerkin_formula_matches_reference_rap(multiband, tag):
    import scipy.sparse as sp
    from multigridcmt_b200.banded import BandedOperator
    op = BandedOperator.from_sparse(sp.csc_matrix(multiband[tag + "_H"]))
    coarse = _galerkin_twin(op)
    ref = multiband[tag + "_RAP"]
    assert np.max(np.abs(coarse.tocsc().toarray() - ref)) <= 1e-14 * np.max(np.abs(ref))
    assert len(coarse.offsets) == (3 if tag == "z0" else 15)
    # and once more down: the diagonal count of the 4-band operators stays at 15
    if tag != "z0":
        assert len(_galerkin_twin(coarse).offsets) == 15


@pytest.mark.parametrize("tag", ["z7", "x7"])
def test_banded_substitution_parameters_match_reference(multiband, tag):
    """(wl, cf, cd, cu, oscale) of band_lower_solve: gseidel = (1,1,0,1,1); sor = (1,1,0,0,w) then (w,0,1-w,w,1) + g"""
    H = multiband[tag + "_H"]
    n = H.shape[0]
    As = H - np.eye(n) * float(multiband[tag + "_shift"])
    x, f = multiband[tag + "_x"], multiband[tag + "_f"]
    D, Lo, Up = np.diag(np.diag(As)), np.tril(As, -1), np.triu(As, 1)

    def lower_solve(wl, cf, cd, cu, oscale, vin, g):
        y = np.linalg.solve(D + wl * Lo, cf * f + cd * (D @ vin) - cu * (Up @ vin))
        return oscale * y + (0 if g is None else g)
    v = x.copy()
    for _ in range(3):
        v = lower_solve(1, 1, 0, 1, 1, v, None)
    assert np.linalg.norm(v - multiband[tag + "_gs"]) < 1e-13 * np.linalg.norm(v)
    w = 1.3
    g = lower_solve(1, 1, 0, 0, w, x, None)
    v = x.copy()
    for _ in range(3):
        v = lower_solve(w, 0, 1 - w, w, 1, v, g)
    assert np.linalg.norm(v - multiband[tag + "_sor"]) < 1e-13 * np.linalg.norm(v)
    # the scan form of the in-chunk recurrence x_l = p_l + q_l x_{l-1} (Hillis-Steele over affine maps)
    rng = np.random.RandomState(0)
    p = rng.random_sample(32) + 1j * rng.random_sample(32)
    q = 0.5 * (rng.random_sample(32) + 1j * rng.random_sample(32))
    q[0] = 0
    seq = np.zeros(32, dtype=complex)
    for l in range(32):
        seq[l] = p[l] + (q[l] * seq[l - 1] if l else 0)
    P, Q, s = p.copy(), q.copy(), 1
    while s < 32:
        Pn, Qn = P.copy(), Q.copy()
        Pn[s:] = P[s:] + Q[s:] * P[:-s]
        Qn[s:] = Q[s:] * Q[:-s]
        P, Q, s = Pn, Qn, 2 * s
    assert np.max(np.abs(P - seq)) < 1e-14


def test_recognition_cache_notices_in_place_edits():
    """drivers reuse one matrix object; editing its values in place (e.g. adding a potential) must not hit a stale entry"""
    import scipy.sparse as sp
    from multigridcmt_b200 import MGCMTStencilMaker
    from multigridcmt_b200.banded import recognise_banded
    from multigridcmt_b200.operators import recognise
    L = sp.csc_matrix(MGCMTStencilMaker().laplacian(64))
    a = recognise(L, "1d")
    assert recognise(L, "1d") is a
    L.data[-1] *= 2.0                      # far from the first few entries
    b = recognise(L, "1d")
    assert b is not a and b.col[1][-1] == 2.0 * a.col[1][-1]
    C = sp.csc_matrix(np.diag(np.arange(1, 9) + 0j) + np.diag(np.ones(6), 2))
    c = recognise_banded(C)
    assert recognise_banded(C) is c
    C.data[-1] += 1j
    d = recognise_banded(C)
    assert d is not c and not d.is_real


# ---------------------------------------------------------------------------------------------------
# property tests of the host-side bookkeeping (hypothesis)
def test_banded_round_trip_random_matrices():
    import scipy.sparse as sp
    from hypothesis import given, settings, strategies as st
    from multigridcmt_b200.banded import BandedOperator

    @settings(max_examples=40, deadline=None)
    @given(st.integers(2, 40), st.lists(st.integers(-39, 39), min_size=1, max_size=6), st.integers(0, 2 ** 31 - 1), st.booleans())
    def check(n, offs, seed, cplx):
        rng = np.random.RandomState(seed)
        offs = sorted({o for o in offs if -n < o < n})
        if not offs:
            offs = [0]
        diags = []
        for o in offs:
            d = rng.random_sample(n - abs(o)) - 0.5
            if cplx:
                d = d + 1j * (rng.random_sample(n - abs(o)) - 0.5)
            d[rng.random_sample(d.size) < 0.2] = 0       # explicit zeros must not matter
            diags.append(d)
        A = sp.diags(diags, offs, shape=(n, n), format="csc")
        op = BandedOperator.from_sparse(A)
        assert abs(op.tocsc() - A).max() == 0
        assert 0 in op.offsets and np.all(np.diff(op.offsets) > 0)
        assert op.is_real == (not np.any(A.toarray().imag))
        x = rng.random_sample(n) + 1j * rng.random_sample(n)
        y = np.zeros(n, dtype=complex)
        for k, o in enumerate(op.offsets):              # the row-indexed storage the kernels use
            for i in range(n):
                if 0 <= i + o < n:
                    y[i] += op.vals[k, i] * x[i + o]
        assert np.allclose(y, A @ x, rtol=1e-13, atol=1e-13)
    check()


def test_plan_levels_invariants():
    from hypothesis import given, settings, strategies as st
    from multigridcmt_b200.slab import HALO, plan_levels

    @settings(max_examples=60, deadline=None)
    @given(st.integers(6, 15), st.sampled_from([1, 2, 4, 8]), st.sampled_from([128, 512, 2048]))
    def check(logn, world, gather):
        n = 1 << logn
        if n % world:
            return
        nlev = plan_levels(n, world, gather)
        own = n // world
        for l in range(nlev):
            assert (n >> l) > gather                       # only levels wider than gather_cols stay decomposed
            assert (own >> l) >= 64 and (own >> l) >= 2 * HALO
            assert own % (1 << (l + 1)) == 0               # cuts stay on even rows of every slab level
        # maximal: the next level would break one of the rules
        l = nlev
        assert not ((n >> l) > gather and (own >> l) >= 64 and own % (1 << (l + 1)) == 0)
    check()


def test_recognise_cache_notices_in_place_edits():
    """an in-place edit that keeps the sum of the values (a well moved along the diagonal) must not return
    the stale operator; a dead matrix's recycled id must not hit either."""
    import scipy.sparse as sp
    from multigridcmt_b200.operators import recognise, invalidate, _RECOGNISED
    n = 64
    V = np.zeros(n); V[10:20] = 5.0
    A = sp.diags([np.ones(n - 1), -2.0 * np.ones(n) + V, np.ones(n - 1)], [-1, 0, 1], format="csc")
    op1 = recognise(A, "1d")
    assert recognise(A, "1d") is op1
    A.setdiag(-2.0 * np.ones(n) + np.roll(V, 7))     # same sum, same nnz, same buffers
    op2 = recognise(A, "1d")
    assert op2 is not op1
    assert np.allclose(op2.diagonal().reshape(-1), A.diagonal())
    invalidate(A)
    assert id(A) not in _RECOGNISED
    op3 = recognise(A, "1d")
    key = id(A)
    del A
    import gc; gc.collect()
    assert key not in _RECOGNISED and op3 is not None


def test_uniform_leg_coefficients_are_consistent_to_double_double():
    """fused_uni.cu multiplies the data by rounded constants; what must hold far beyond 1e-16 is their RATIO (it is the
    diagonal of the operator the sweeps and the residual effectively use: 1e-16 * d = 7e-10 at 4096^2).  Host-only."""
    from fractions import Fraction as F
    import ctypes as C
    lib = _lib.load()
    out = (C.c_double * 7)()
    for N in (64, 1024, 4096, 16384):
        c = (-1.0 / np.pi ** 2) * float(N) ** 2
        d = -4.0 * c
        for shift in (0.0, 1.76659015, 4.38639582, 7.00620149):
            for omega in (2.0 / 3.0, 1.0, 1.3):
                assert lib.mgcmt_debug_uni_coefficients(c, d, shift, omega, out) == 0
                a_smooth, a_res, dlo, nbeta, wf, invw, drem = list(out)
                om = -(F(a_res) + F(dlo))                       # the weight the kernel applies: hi + lo
                exact = F(-nbeta) * (F(d) - F(shift)) / F(c)    # beta (d - shift) / c, exactly
                assert abs(om - exact) <= abs(exact) * F(1, 10 ** 28), (N, shift, omega, float(om - exact))
                assert F(a_smooth) + F(dlo) == 1 - om           # sweep coefficient 1 - om, same low part
                assert abs(om - F(omega)) <= F(omega) * F(1, 10 ** 15)      # and it IS the caller's omega to rounding
                assert abs(F(drem) - (F(d) + 4 * F(c) - F(shift))) <= abs(F(shift)) * F(1, 10 ** 15) + F(1, 10 ** 9) * 0 + abs(F(d)) * F(1, 10 ** 15)
                if abs(omega - 1.0) < 1e-12:
                    assert a_res == -1.0                        # a weight within ulps of 1 never multiplies the data


def test_leg_chunk_height_fills_whole_waves():
    """the chunk height of a streaming leg: even, within bounds, and never a few CTAs over a wave"""
    lib = _lib.load()
    for nrows, gx, slots, nstage in ((4096, 10, 296, 5), (4096, 10, 444, 5), (2048, 12, 296, 9), (2048, 5, 296, 5), (1024, 6, 296, 9),
                                     (512, 3, 296, 9), (128, 1, 296, 9), (16384, 37, 296, 5), (2060, 37, 296, 5)):
        rpc = lib.mgcmt_debug_leg_rows_per_chunk(nrows, gx, slots, nstage, 1 << 20)
        assert rpc % 2 == 0 and 2 <= rpc <= max(nrows, 16)
        ctas = gx * ((nrows + rpc - 1) // rpc)
        waves = (ctas + slots - 1) // slots
        # the last wave is at least 80 % full unless the whole grid is smaller than one wave
        assert ctas <= slots or ctas >= (waves - 1) * slots + 0.8 * slots or rpc == 16, (nrows, gx, slots, rpc, ctas)
        assert lib.mgcmt_debug_leg_rows_per_chunk(nrows, gx, slots, nstage, 128) <= 128


def test_vcycle_many_argument_checks_need_no_gpu():
    """MGCMTSolver.vcycle_many (the drivers' loop body as one call): argument errors are raised before anything touches
    the device, and the product path still refuses to run without CUDA (no CPU fallback)."""
    from multigridcmt_b200 import MGCMTSolver, MGCMTStencilMaker
    s, sm = MGCMTSolver(), MGCMTStencilMaker()
    H = (-1. / np.pi ** 2) * sm.laplacian(512, "2d", matrix_free=True)
    fs = [np.zeros(512 * 512), np.zeros(512 * 512)]
    with pytest.raises(ValueError):
        s.vcycle_many(fs, H, sm, [1.0], lowest_level=8, dimension="2d")          # one shift for two vectors
    assert s.vcycle_many([], H, sm, [], lowest_level=8, dimension="2d") == []
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            s.vcycle_many(fs, H, sm, [1.0, 2.0], lowest_level=8, dimension="2d")


def test_native_slab_phase_table():
    """csrc/slab_block.cu build_phases: the order of (halo exchange, leg) phases of one multi-GPU cycle.  Weighted Jacobi:
    one phase per level and direction.  Red-black Gauss-Seidel: the 5-point level still one phase per direction, every
    9-point level two (two passes of two sweeps with an exchange of the intermediate iterate in between)."""
    import ctypes as C
    from multigridcmt_b200 import _lib
    lib = _lib.load()
    DOWN, DOWN_A, DOWN_B, COARSE, UP, UP_A, UP_B, RQ = range(8)

    def phases(nlev, smoother, rq):
        kinds, levels = (C.c_int * 64)(), (C.c_int * 64)()
        n = lib.mgcmt_debug_slab_phases(nlev, smoother, rq, kinds, levels, 64)
        assert n > 0
        return [(kinds[i], levels[i]) for i in range(n)]

    for nlev in (1, 2, 5):
        wj = phases(nlev, _lib.SMOOTH_WJACOBI, 0)
        assert wj == ([(DOWN, l) for l in range(nlev)] + [(COARSE, nlev)] + [(UP, l) for l in range(nlev - 1, -1, -1)])
        assert phases(nlev, _lib.SMOOTH_WJACOBI, 1) == wj + [(RQ, 0)]
        gs = phases(nlev, _lib.SMOOTH_RBGS, 0)
        want = [(DOWN, 0)]
        for l in range(1, nlev):
            want += [(DOWN_A, l), (DOWN_B, l)]
        want.append((COARSE, nlev))
        for l in range(nlev - 1, 0, -1):
            want += [(UP_A, l), (UP_B, l)]
        want.append((UP, 0))
        assert gs == want
        assert len(gs) == 4 * nlev - 1          # what NativeSlabBlock.profile_read sizes its arrays for (<= 4 nlev + 2)
    kinds, levels = (C.c_int * 2)(), (C.c_int * 2)()
    assert lib.mgcmt_debug_slab_phases(5, _lib.SMOOTH_RBGS, 0, kinds, levels, 2) == -1
