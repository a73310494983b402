"""Times the finest-level V-cycle legs (and level-1 legs) of the 2-D well with CUDA events, one kernel at a time,
for every implementation switch given on the command line.  One line per (variant, leg).

  python tools/leg_bench.py [N] [variant ...]      variant = name:opt=val,opt=val   (mgcmt_set_option pairs)

e.g.  python tools/leg_bench.py 4096 general:fused_uni=0 uni3:fused_uni=1 uni2:fused_uni=1,uni_minctas=2
Bytes per unknown: zero-start down leg 18 (f in, v + r/4 out), down leg 26, up leg 26 (v, f, e/4 in, v out).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multigridcmt_b200 import MGCMTStencilMaker, _lib
from multigridcmt_b200.hierarchy import get_hierarchy

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
variants = sys.argv[2:] or ["default:"]
lib = _lib.load()
sm = MGCMTStencilMaker()
H = (-1.0 / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
h = get_hierarchy(H, 8)
g = torch.Generator(device="cuda"); g.manual_seed(0)
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0


def bufs(level):
    n = (N >> level) ** 2
    v = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    f = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    e = torch.rand(n // 4, dtype=torch.float64, device="cuda", generator=g)
    return n, v, f, torch.empty_like(v), e, torch.empty(n // 4, dtype=torch.float64, device="cuda")


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2


def timeit(fn, reps=10):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for var in variants:
    name, _, opts = var.partition(":")
    for kv in filter(None, opts.split(",")):
        k, v_ = kv.split("=")
        _lib.check(lib.mgcmt_set_option(k.encode(), int(v_)))
    for level in (0, 1):
        n, v, f, out, e, rc = bufs(level)
        rqs = torch.zeros(2, dtype=torch.float64, device="cuda")
        legs = [("wj down0", 2, 4, 18, lambda: h.fused_leg(level, 2, 4, 1.7, 2. / 3., None, f, out, None, rc)),
                ("wj down", 1, 4, 26, lambda: h.fused_leg(level, 1, 4, 1.7, 2. / 3., v, f, out, None, rc)),
                ("wj up", 3, 4, 26, lambda: h.fused_leg(level, 3, 4, 1.7, 2. / 3., v, f, out, e, None)),
                ("gs down0", 2, 4, 18, lambda: h.fused_leg(level, 32 | 2, 4 if level == 0 else 2, 1.7, 1.0, None, f, out, None, rc)),
                ("gs down", 1, 4, 26, lambda: h.fused_leg(level, 32 | 1, 4 if level == 0 else 2, 1.7, 1.0, v, f, out, None, rc)),
                ("gs up", 3, 4, 26, lambda: h.fused_leg(level, 32 | 3, 4 if level == 0 else 2, 1.7, 1.0, v, f, out, e, None))]
        for lname, mode, nu, bpu, fn in legs:
            med, best = timeit(fn)
            gbs = bpu * n / (med * 1e-3) / 1e9
            print("%-10s L%d %-9s %8.1f us (best %7.1f)  %6.0f GB/s algorithmic = %.2f of %.0f" % (name, level, lname, med * 1e3, best * 1e3, gbs, gbs / peak, peak), flush=True)
    if N >= 512:
        # whole V-cycles (zero start) for both smoothers, with and without the Rayleigh sums
        n, v, f, out, e, rc = bufs(0)
        rqs = torch.zeros(2, dtype=torch.float64, device="cuda")
        for sname, code, om in (("wjacobi", _lib.SMOOTH_WJACOBI, 2. / 3.), ("rbgs", _lib.SMOOTH_RBGS, 1.0)):
            med, best = timeit(lambda: h.vcycle(1.7, 4, 4, code, om, out, f, v0_is_zero=True), reps=5)
            print("%-10s vcycle %-8s %8.1f us (best %7.1f)" % (name, sname, med * 1e3, best * 1e3), flush=True)
            def vrq():
                _lib.check(lib.mgcmt_vcycle_rq(h.handle, 1.7, 4, 4, code, om, out.data_ptr(), f.data_ptr(), 1, rqs.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
            med, best = timeit(vrq, reps=5)
            print("%-10s vcycle_rq %-5s %8.1f us (best %7.1f)" % (name, sname, med * 1e3, best * 1e3), flush=True)
