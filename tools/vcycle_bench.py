"""Whole V-cycle time (zero start) for both smoothers under option variants:  python tools/vcycle_bench.py N lowest name:opt=val,... ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigridcmt_b200 import MGCMTStencilMaker, _lib
from multigridcmt_b200.hierarchy import get_hierarchy
N = int(sys.argv[1]); low = int(sys.argv[2]); variants = sys.argv[3:] or ["default:"]
lib = _lib.load(); sm = MGCMTStencilMaker()
H = (-1.0 / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
h = get_hierarchy(H, low)
g = torch.Generator(device="cuda"); g.manual_seed(0)
f = torch.rand(N * N, dtype=torch.float64, device="cuda", generator=g); out = torch.empty_like(f)
ref = {}
for var in variants:
    name, _, opts = var.partition(":")
    for kv in filter(None, opts.split(",")):
        k, v_ = kv.split("="); _lib.check(lib.mgcmt_set_option(k.encode(), int(v_)))
    for sname, code, om in (("wjacobi", _lib.SMOOTH_WJACOBI, 2. / 3.), ("rbgs", _lib.SMOOTH_RBGS, 1.0)):
        for _ in range(3):
            h.vcycle(1.7, 4, 4, code, om, out, f, v0_is_zero=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            h.vcycle(1.7, 4, 4, code, om, out, f, v0_is_zero=True)
        e1.record(); torch.cuda.synchronize()
        r = ref.setdefault(sname, out.clone())
        print("%-12s N=%d low=%d %-8s %8.1f us per cycle   rel diff to first variant %.1e" % (name, N, low, sname, e0.elapsed_time(e1) / 20 * 1e3, float((out - r).norm() / r.norm())), flush=True)
