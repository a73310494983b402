import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigridcmt_b200 import MGCMTStencilMaker, MGCMTSolver
from multigridcmt_b200.slab import LocalComm, SlabVCycle
N, world, gather = 512, 2, 128
sm, s = MGCMTStencilMaker(), MGCMTSolver()
H = (-1. / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
f = np.random.RandomState(3).random_sample(N * N)
sv = SlabVCycle(H, world, LocalComm(world), range(world), lowest_level=8, gather_cols=gather)
print("nlev", sv.nlev)
sv.scatter("f", f)
sv.vcycle(4.386, v0_is_zero=True)
got = sv.gather_local("v").reshape(N, N)
want = s.vcycle(np.zeros(N * N), f.copy(), H, sm, shift=4.386, lowest_level=8, dimension="2d").reshape(N, N)
err = np.abs(got - want)
rows = err.max(axis=1); cols = err.max(axis=0)
print("max err", err.max(), "scale", np.abs(want).max())
print("rows with err>1e-12:", np.nonzero(rows > 1e-12)[0][:40], "count", (rows > 1e-12).sum())
print("cols with err>1e-12:", np.nonzero(cols > 1e-12)[0][:40], "count", (cols > 1e-12).sum())
print("row err profile", rows[:12], rows[250:262])
