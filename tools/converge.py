"""How far does the shift method get?  python tools/converge.py [N] [lowest] [max_iters]
For each smoother / form / shift policy: iterations until ||H v - rho v|| <= 1e-10 and |rho - closed form| <= 1e-10."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigridcmt_b200 import MGCMTStencilMaker
from multigridcmt_b200.eigensolver import ShiftMethod, well_eigenvalue_1d, well_start_block
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
low = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 60
MODES = [(1, 1), (1, 2), (2, 1), (2, 2)]
sm = MGCMTStencilMaker()
H = (-1. / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
V0, shifts = well_start_block(N, MODES)
exact = np.array([well_eigenvalue_1d(N, a) + well_eigenvalue_1d(N, b) for a, b in MODES])
for smoother in ("wjacobi", "rbgs"):
    for form, upd in (("reference", False), ("correction", False), ("reference", True), ("correction", True)):
        loop = ShiftMethod(H, shifts, V0, dimension="2d", lowest_level=low, smoother=smoother, ortho="gram")
        torch.cuda.synchronize(); t = time.time()
        r = loop.solve(tol=1e-10, max_iters=iters, form=form, update_shift=upd, exact=exact)
        torch.cuda.synchronize(); dt = time.time() - t
        h = r["history"]
        trace = " ".join("%d:%.0e" % (i, res.max()) for i, res, rho in h[:: max(1, len(h) // 8)])
        print("N=%d low=%d %-8s %-10s update_shift=%d: converged=%s iters=%d  max res %.1e  max |rho-exact| %.1e  %.2fs  [%s]"
              % (N, low, smoother, form, upd, r["converged"], r["iterations"], r["residual_norms"].max(),
                 np.abs(r["eigenvalues"] - exact).max(), dt, trace), flush=True)
