"""CPU: the matrix-free C/OpenMP oracle (oracle/mgcmt_oracle.c, the timed CPU baseline) against the numpy oracle."""
import numpy as np
import pytest

import c_oracle
import mgcmt_oracle as orc


@pytest.mark.parametrize("N,low,shift,nu", [(32, 8, 0.0, (4, 4)), (64, 8, 4.38639582, (4, 4)), (128, 8, 1.76659015, (2, 3)),
                                             (64, 2, 0.0, (4, 4))])
def test_c_oracle_matches_numpy_oracle(N, low, shift, nu):
    osm, osv = orc.StencilMaker(), orc.Solver()
    H = (-1. / np.pi ** 2) * osm.laplacian(N, "2d")
    rs = np.random.RandomState(N)
    v0, f = rs.random_sample(N * N), rs.random_sample(N * N)
    want = osv.vcycle(v0.copy(), f.copy(), H, osm, nu1=nu[0], nu2=nu[1], shift=shift, lowest_level=low, dimension="2d")
    got = c_oracle.WellHierarchy(N, low).vcycle(v0, f, shift, nu1=nu[0], nu2=nu[1])
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-11


@pytest.mark.parametrize("N,low,shift,omega", [(32, 8, 0.0, 1.0), (64, 8, 4.38639582, 1.0), (64, 8, 1.76659015, 1.15)])
def test_c_oracle_rbgs_matches_numpy_oracle(N, low, shift, omega):
    """the red-black (four-colour) smoother: C twin == numpy twin (Solver.rbgs), whole V-cycles"""
    import functools
    osm, osv = orc.StencilMaker(), orc.Solver()
    H = (-1. / np.pi ** 2) * osm.laplacian(N, "2d")
    rs = np.random.RandomState(100 + N)
    f = rs.random_sample(N * N)
    want = osv.vcycle(np.zeros(N * N), f.copy(), H, osm, shift=shift, lowest_level=low, dimension="2d",
                      smoother=functools.partial(osv.rbgs, omega=omega))
    got = c_oracle.WellHierarchy(N, low).vcycle(np.zeros(N * N), f, shift, smoother="rbgs", omega=omega)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-11


def test_c_oracle_block_step_matches_numpy_oracle():
    """the reference arm's step (k V-cycles + Rayleigh sums + modified Gram-Schmidt, 2DPotGS.py:91-105) in C/OpenMP
    against the same loop written with the numpy oracle; lowest_level = 16 exercises the banded coarsest LU"""
    import c_oracle
    import mgcmt_oracle as orc
    N = 128
    osm, osv, op = orc.StencilMaker(), orc.Solver(), orc.Processor()
    H = (-1 / np.pi ** 2) * osm.laplacian(N, "2d")
    V = np.random.RandomState(0).random_sample((3, N * N))
    shifts = [1.7, 4.3, 4.4]
    for low, smoother in ((8, "wjacobi"), (16, "rbgs")):
        blk = c_oracle.ShiftBlock(N, low, shifts, smoother=smoother)
        W = np.zeros_like(V)
        lam = blk.step(V, W)
        kw = {"smoother": osv.rbgs} if smoother == "rbgs" else {}
        Wn = np.stack([osv.vcycle(np.zeros(N * N), V[c].copy(), H, osm, shift=shifts[c], dimension="2d", lowest_level=low, **kw)
                       for c in range(3)])
        lam_n = np.array([float(w @ (H @ w)) / float(w @ w) for w in Wn])
        Q = op.gramschmidt(Wn.T.copy()).T
        assert np.abs(lam - lam_n).max() < 1e-11
        assert np.abs(Q - W).max() < 1e-11
