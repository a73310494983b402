"""Rayleigh sums of a V-cycle output: fused stage (uni / general kernels) vs the separate pass vs long-double on the host."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from multigridcmt_b200 import MGCMTStencilMaker, _lib
from multigridcmt_b200.hierarchy import _ptr, _stream_ptr, get_hierarchy
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _lib.load(); sm = MGCMTStencilMaker()
H = (-1. / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
h = get_hierarchy(H, 8)
f = torch.from_numpy(np.random.RandomState(33).random_sample(N * N)).cuda()
L = np.longdouble
c = L(-1.) / L(np.pi) ** 2 * 0 + L((-1. / np.pi ** 2) * float(N) ** 2)
for smoother, om in ((_lib.SMOOTH_WJACOBI, 2. / 3.), (_lib.SMOOTH_RBGS, 1.0)):
    for shift in (4.38639582, 1.7):
        res = {}
        for uni in (1, 0):
            lib.mgcmt_set_option(b"fused_uni", uni)
            w = torch.zeros(N * N, dtype=torch.float64, device="cuda"); out = torch.zeros(2, dtype=torch.float64, device="cuda")
            _lib.check(lib.mgcmt_vcycle_rq(h.handle, shift, 4, 4, smoother, om, _ptr(w), _ptr(f), 1, _ptr(out), _stream_ptr(torch)))
            ref = torch.zeros(2, dtype=torch.float64, device="cuda"); h.rayleigh(0, w, ref)
            x = w.cpu().numpy().reshape(N, N).astype(L)
            X = np.pad(x, 1)
            Ax = c * (X[:-2, 1:-1] + X[2:, 1:-1] + X[1:-1, :-2] + X[1:-1, 2:] - 4 * x)
            exact = (float((x * Ax).sum()), float((x * x).sum()))
            o, r = out.cpu().numpy(), ref.cpu().numpy()
            print("smoother %d shift %.3f uni %d: fused num rel err %.2e  separate-pass rel err %.2e   (den %.1e / %.1e)"
                  % (smoother, shift, uni, abs(o[0] - exact[0]) / abs(exact[0]), abs(r[0] - exact[0]) / abs(exact[0]),
                     abs(o[1] - exact[1]) / exact[1], abs(r[1] - exact[1]) / exact[1]), flush=True)
lib.mgcmt_set_option(b"fused_uni", 1)
