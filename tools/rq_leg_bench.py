"""Rayleigh sums: fused into the finest up leg (mgcmt_vcycle_rq) vs up leg + separate pass, per smoother."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigridcmt_b200 import MGCMTStencilMaker, _lib
from multigridcmt_b200.hierarchy import get_hierarchy, _ptr, _stream_ptr
N = 4096; lib = _lib.load(); sm = MGCMTStencilMaker()
H = (-1.0 / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
h = get_hierarchy(H, 64)
g = torch.Generator(device="cuda"); g.manual_seed(0)
f = torch.rand(N * N, dtype=torch.float64, device="cuda", generator=g); out = torch.empty_like(f); rq = torch.zeros(2, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, reps=8):
    fn(); fn(); ts = []
    for _ in range(reps):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3
for sname, code, om in (("wjacobi", _lib.SMOOTH_WJACOBI, 2. / 3.), ("rbgs", _lib.SMOOTH_RBGS, 1.0)):
    a = t(lambda: h.vcycle(1.7, 4, 4, code, om, out, f, v0_is_zero=True))
    b = t(lambda: _lib.check(lib.mgcmt_vcycle_rq(h.handle, 1.7, 4, 4, code, om, _ptr(out), _ptr(f), 1, _ptr(rq), _stream_ptr(torch))))
    c = t(lambda: h.rayleigh(0, out, rq))
    print("%-8s vcycle %.1f us  vcycle_rq (fused stage) %.1f us  (+%.1f)   separate Rayleigh pass %.1f us" % (sname, a, b, b - a, c))
