"""Runs a few launches of one kernel family on the 4096^2 well so `ncu -k regex:...` can capture it.

  python tools/profile_sweep.py [jacobi|vcycle|down|down0|up|gsdown|gsdown0|gsup|down1|up1|gsdown1] [N] [option=value ...]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multigridcmt_b200 import MGCMTStencilMaker, _lib
from multigridcmt_b200.hierarchy import get_hierarchy

what = sys.argv[1] if len(sys.argv) > 1 else "jacobi"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
for kv in sys.argv[3:]:   # option=value pairs for mgcmt_set_option
    k, v = kv.split("=")
    _lib.check(_lib.load().mgcmt_set_option(k.encode(), int(v)))
sm = MGCMTStencilMaker()
H = (-1.0 / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
LOW = int(os.environ.get("LOWEST", "8"))
h = get_hierarchy(H, LOW)
g = torch.Generator(device="cuda"); g.manual_seed(0)
v = torch.rand(N * N, dtype=torch.float64, device="cuda", generator=g)
f = torch.rand(N * N, dtype=torch.float64, device="cuda", generator=g)
out = torch.empty_like(v); rc = torch.rand(N * N // 4, dtype=torch.float64, device="cuda", generator=g)
v1 = torch.rand(N * N // 4, dtype=torch.float64, device="cuda", generator=g); f1 = v1 * 0.5; out1 = torch.empty_like(v1)
rc1 = torch.empty(N * N // 16, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
for it in range(2):
    e0.record()
    for _ in range(reps):
        if what == "jacobi":
            h.smooth(0, _lib.SMOOTH_WJACOBI, 1.7, 2.0 / 3.0, 4, v, f)
        elif what == "down":
            h.fused_leg(0, 1, 4, 1.7, 2.0 / 3.0, v, f, out, None, rc)
        elif what == "up":
            h.fused_leg(0, 3, 4, 1.7, 2.0 / 3.0, v, f, out, rc, None)
        elif what in ("vcycle_rq_gs", "vcycle_rq_wj"):
            gs = what.endswith("gs")
            _lib.check(_lib.load().mgcmt_vcycle_rq(h.handle, 1.7, 4, 4, _lib.SMOOTH_RBGS if gs else _lib.SMOOTH_WJACOBI,
                                                   1.0 if gs else 2.0 / 3.0, out.data_ptr(), f.data_ptr(), 1, rc.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream))
        elif what == "vcycle_gs":
            h.vcycle(1.7, 4, 4, _lib.SMOOTH_RBGS, 1.0, out, f, v0_is_zero=True)
        elif what == "vcycle_wj":
            h.vcycle(1.7, 4, 4, _lib.SMOOTH_WJACOBI, 2.0 / 3.0, out, f, v0_is_zero=True)
        elif what == "down0":
            h.fused_leg(0, 2, 4, 1.7, 2.0 / 3.0, None, f, out, None, rc)
        elif what == "gsdown":
            h.fused_leg(0, 32 | 1, 4, 1.7, 1.0, v, f, out, None, rc)
        elif what == "gsdown0":
            h.fused_leg(0, 32 | 2, 4, 1.7, 1.0, None, f, out, None, rc)
        elif what == "gsup":
            h.fused_leg(0, 32 | 3, 4, 1.7, 1.0, v, f, out, rc, None)
        elif what == "up1":
            h.fused_leg(1, 3, 4, 1.7, 2.0 / 3.0, v1, f1, out1, rc1, None)
        elif what == "gsdown1":
            h.fused_leg(1, 32 | 1, 2, 1.7, 1.0, v1, f1, out1, None, rc1)
        elif what == "down1":
            h.fused_leg(1, 1, 4, 1.7, 2.0 / 3.0, v1, f1, out1, None, rc1)
        else:
            h.vcycle(1.7, 4, 4, _lib.SMOOTH_WJACOBI, 2.0 / 3.0, v, f)
    e1.record()
    torch.cuda.synchronize()
print(what, N, "ms per call:", e0.elapsed_time(e1) / reps)
