"""GPU V-cycle against the C oracle at a large size, for every implementation switch:  python tools/parity_large.py [N]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import c_oracle
from multigridcmt_b200 import MGCMTSolver, MGCMTStencilMaker, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lows = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 64]
lib = _lib.load()
sm, s = MGCMTStencilMaker(), MGCMTSolver()
H = (-1. / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
rs = np.random.RandomState(N)
v0, f = rs.random_sample(N * N), rs.random_sample(N * N)
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
for low in lows:
    orc = c_oracle.WellHierarchy(N, low)
    for shift in (0.0, 4.38639582, 1.76659015):
        for smoother in ("wjacobi", "rbgs"):
            want = orc.vcycle(v0, f, shift, smoother=smoother)
            outs = {}
            for uni in (1, 0):
                lib.mgcmt_set_option(b"fused_uni", uni)
                t = time.time()
                got = s.vcycle(v0.copy(), f.copy(), H, sm, shift=shift, lowest_level=low, dimension="2d",
                               **({"smoother": s.rbgs} if smoother == "rbgs" else {}))
                outs[uni] = (rel(got, want), time.time() - t, got)
            print("N=%d low=%d shift=%.4f %-8s  uni: %.2e (%.2fs)  general: %.2e (%.2fs)  uni vs general: %.2e"
                  % (N, low, shift, smoother, outs[1][0], outs[1][1], outs[0][0], outs[0][1], rel(outs[1][2], outs[0][2])), flush=True)
lib.mgcmt_set_option(b"fused_uni", 1)
