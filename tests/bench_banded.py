"""Time the banded complex V-cycle (multiband Hamiltonians, ThesisProblem.py sizes and larger) on the GPU,
with the numpy/scipy oracle on the host beside it as checker and CPU reference.  Not part of bench.py's headline: a side
measurement for DESIGN.md.  Lives under tests/ because it uses oracle/ (test infrastructure); not collected by pytest.
usage: python tests/bench_banded.py [gridpoints_per_band ...]"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def multiband(g):
    """4 coupled bands, tridiagonal blocks, complex couplings: the structure PotWellSolver.makeMatrix produces"""
    def tri(d, e):
        return sp.diags([np.full(g - 1, e), np.full(g, d), np.full(g - 1, np.conj(e))], [-1, 0, 1], format="csc")
    step = 2.0 / g
    P = tri(6.85 * 2 / step ** 2 / np.pi ** 2, -6.85 / step ** 2 / np.pi ** 2)
    Q = tri(2.1 * 2 / step ** 2 / np.pi ** 2, -2.1 / step ** 2 / np.pi ** 2)
    S = tri(0.0, 0.3j / step)
    Rm = sp.diags([np.full(g, -0.2 + 0.1j)], [0], format="csc")
    V = sp.diags([np.where(np.abs(np.linspace(-1, 1, g)) > 0.5, 40.0, 0.0)], [0], format="csc")
    Z = sp.csc_matrix((g, g))
    return sp.bmat([[P + Q + V, -S, Rm, Z], [-S.conj().T, P - Q + V, Z, Rm], [Rm.conj().T, Z, P - Q + V, S],
                    [Z, Rm.conj().T, S.conj().T, P + Q + V]], format="csc")


def main():
    import torch
    import mgcmt_oracle as orc
    from multigridcmt_b200 import _lib
    from multigridcmt_b200.banded import BandedHierarchy, BandedOperator
    grids = [int(a) for a in sys.argv[1:]] or [256, 4096, 65536]
    osm, osolver = orc.StencilMaker(), orc.Solver()
    out = []
    for g in grids:
        H = multiband(g)
        n = 4 * g
        r = np.random.RandomState(3)
        f = (r.random_sample(n) - 0.5) + 1j * (r.random_sample(n) - 0.5)
        h = BandedHierarchy(BandedOperator.from_sparse(H), 32)
        fd = torch.from_numpy(f).cuda()
        v = torch.zeros_like(fd)
        row = {"unknowns": n, "levels": h.num_levels, "diagonals": [h.level_shape(l)[1] for l in range(h.num_levels)]}
        for name, code, osmo in (("wjacobi", _lib.SMOOTH_WJACOBI, None), ("gseidel", _lib.SMOOTH_GSLEX, osolver.gseidel)):
            om = 2. / 3. if name == "wjacobi" else 1.0
            for _ in range(3):
                v.zero_(); h.vcycle(3.0, 4, 4, code, om, v, fd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                v.zero_(); h.vcycle(3.0, 4, 4, code, om, v, fd)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            t0 = time.perf_counter()
            ref = osolver.vcycle(np.zeros(n), f.copy(), H, osm, shift=3.0, lowest_level=32, smoother=osmo)
            t_first = time.perf_counter() - t0
            t0 = time.perf_counter()
            ref = osolver.vcycle(np.zeros(n), f.copy(), H, osm, shift=3.0, lowest_level=32, smoother=osmo)
            t_cpu = time.perf_counter() - t0
            err = float(np.linalg.norm(v.cpu().numpy() - ref) / np.linalg.norm(ref))
            row[name] = {"gpu_ms_per_vcycle": ms, "oracle_ms_per_vcycle_hierarchy_cached": t_cpu * 1e3,
                         "oracle_ms_first_call": t_first * 1e3, "rel_diff": err}
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
