#!/usr/bin/env python
"""bench.py -- throughput of the multigrid V-cycle eigensolver path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n GRID] [--smoother S]

A "step" is one outer iteration of the shift method on the block of the 4 lowest eigenvectors of the 2-D
infinite well (2DPotGS.py:91-105): for each vector one V(4,4)-cycle of (H - mu_i I) w = v_i, normalise,
Rayleigh quotient; then modified Gram-Schmidt of the block.  Default workload = BASELINE.json configs[2]:
4096^2, red-black Gauss-Seidel smoother, 7 levels (lowest_level = 64, exact coarsest solve on 64^2); the
weighted-Jacobi / lowest_level = 8 variant (the reference's own 2-D choice, 2DPot.py:89) is a side line.
`value` counts smoother unknown-updates/s (8 sweeps on every smoothing level of every cycle) with all vectors
resident in HBM; `e2e` is the same cycles through the reference-facing MGCMTSolver.vcycle call with HOST numpy
arrays, host<->device copies inside the timed region.  `roofline` is per kernel: bytes that leg must move
(18 B per unknown for the zero-start down leg, 26 B for the up leg) / its CUDA-event time / measured HBM peak.

Prints ONE JSON line (rank 0).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "2D well V-cycle unknown-updates/s"
UNIT = "unknown-updates/s"
MODES = [(1, 1), (1, 2), (2, 1), (2, 2)]
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
# capture (profiles/r1_fused_down_4096.txt); only valid for the default 4096^2 workload
TRAFFIC_NCU = {}
try:
    TRAFFIC_NCU = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass
N0 = 16  # coarse grid the initial guesses / shifts come from (2DPotGS.py:54-63; closed form instead of eigsh)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=0, help="grid size N (N x N unknowns); default 4096 (16384 for --gpus > 1)")
    ap.add_argument("--lowest", type=int, default=0,
                    help="lowest_level; default 64 (BASELINE config 3: 7 levels at 4096^2) on 1 GPU, 8 (2DPot.py:89) on the slab path")
    ap.add_argument("--smoother", default="", choices=["", "wjacobi", "rbgs"],
                    help="default rbgs (BASELINE config 3) on 1 GPU, wjacobi on the slab path")
    ap.add_argument("--no-side", action="store_true", help="skip the side lines (other smoother / lowest_level, convergence run, scaling base)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-n", type=int, default=2048, help="grid size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams for the independent V-cycles of a step")
    ap.add_argument("--ortho", default="gram", choices=["mgs", "gram"],
                    help="block orthonormalisation: column-by-column MGS (MGCMTProcessor.py:44-50) or its Gram-matrix form")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--replicas", action="store_true", help="multi-GPU: independent replicas instead of row slabs")
    ap.add_argument("--gather-cols", type=int, default=512, help="slab path: levels at most this wide are replicated")
    ap.add_argument("--slab-stagger", type=int, default=1,
                    help="native slab driver: run the block as two halves half a phase apart (exchange of one overlaps kernels of the other)")
    ap.add_argument("--slab-driver", choices=["native", "python"], default="native",
                    help="slab path: step issued from C++ with NCCL called directly (csrc/slab_block.cu), or from Python over torch.distributed")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# closed-form pieces: no RNG, no eigsh
# ---------------------------------------------------------------------------------------------------
def ev1(n, k):
    import numpy as np
    return (4.0 * n * n / np.pi ** 2) * np.sin(k * np.pi / (2.0 * (n + 1))) ** 2


def vec1(n, k):
    import numpy as np
    v = np.sin(k * np.pi * (np.arange(n) + 1.0) / (n + 1.0))
    return v / np.linalg.norm(v)


def interp1(N):
    """1-D interpolation N0 -> N as a dense (N, N0) matrix (MGCMTStencilMaker.py:27-43)."""
    from multigridcmt_b200 import MGCMTStencilMaker
    return MGCMTStencilMaker().interpolation(N0, N).toarray()


def initial_block(N):
    """(4, N*N) start vectors P(16->N,'2d') * closed-form N=16 eigenvectors, normalised; and shifts."""
    import numpy as np
    P = interp1(N)
    shifts = [ev1(N0, a) + ev1(N0, b) for a, b in MODES]
    V = np.empty((4, N * N))
    for c, (a, b) in enumerate(MODES):
        V[c] = np.kron(P @ vec1(N0, a), P @ vec1(N0, b))  # (P (x) P)(va (x) vb)
        V[c] /= np.linalg.norm(V[c])
    return V, shifts


def updates_per_cycle(N, lowest):
    """smoother unknown-updates in one V(4,4): 8 sweeps on every level except the coarsest."""
    tot, g = 0, N
    while g > lowest:
        tot += 8 * g * g
        g //= 2
    return tot


# ---------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------------
class StdoutToStderr:
    """fd-level redirect of stdout into stderr (C libraries that printf: stdout must carry ONE JSON line)"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout=3.0):
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if t_begin is not None and not (t_begin <= ts <= t_end + 0.1):
                continue
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        med = statistics.median(sm) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (the oracle, single-threaded scipy: the reference's own CPU path cannot run at these sizes,
# SURVEY.md D9 / section 8(d))
# ---------------------------------------------------------------------------------------------------
def omp_threads():
    try:
        return int(os.environ.get("OMP_NUM_THREADS", "")) or (os.cpu_count() or 1)
    except ValueError:
        return os.cpu_count() or 1


def workload_config(N, smoother, lowest, k=4):
    """the `config` both arms print: what is computed, nothing about how"""
    levels, g = 1, N
    while g > lowest:
        levels += 1
        g //= 2
    return {"workload": "2D infinite well %d^2, lowest %d eigenpairs, shift method (2DPotGS.py:91-105): %d x V(4,4) + Rayleigh "
                        "quotient + modified Gram-Schmidt per step" % (N, k, k),
            "grid": N, "eigenpairs": k, "smoother": smoother, "lowest_level": lowest, "levels": levels,
            "l2": ("working set %.1f GB >> 126 MB L2 (inputs larger than L2, no flush needed)" % (10 * N * N * 8 / 1e9)
                   if 10 * N * N * 8 > 4 * 126e6 else "working set %.2f GB: NOT larger than the 126 MB L2 (non-default --n)" % (10 * N * N * 8 / 1e9))}


def start_block_host(N):
    import numpy as np
    P = interp1(N)
    shifts = [ev1(N0, a) + ev1(N0, b) for a, b in MODES]
    V = np.stack([np.kron(P @ vec1(N0, a), P @ vec1(N0, b)) for a, b in MODES])
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    return np.ascontiguousarray(V), shifts


def cpu_block_sample(N, smoother, lowest, steps=2, budget_s=25.0):
    """The matrix-free C/OpenMP port of the path (oracle/mgcmt_oracle.c, held to the numpy oracle by
    tests/test_c_oracle.py) on all host cores: `steps` whole shift-method steps of the bench workload (k V-cycles +
    Rayleigh sums + modified Gram-Schmidt), after one untimed step that also factors the coarsest operators."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    V, shifts = start_block_host(N)
    W = np.zeros_like(V)
    blk = c_oracle.ShiftBlock(N, lowest, shifts, smoother=smoother)
    blk.step(V, W)
    V, W = W, V
    done = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        blk.step(V, W)
        V, W = W, V
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return len(MODES) * updates_per_cycle(N, lowest) * done / dt, dt, done


def cpu_sample(n_cpu, lowest, repeats=1):
    """Time the numpy/scipy oracle (the port that is pinned to the real reference) on one V(4,4)-cycle at n_cpu^2 with
    the hierarchy already built (the reference rebuilds it every call; leaving that out favours the CPU)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mgcmt_oracle as orc
    osm, osv = orc.StencilMaker(), orc.Solver(cache_hierarchy=True)
    H = (-1.0 / np.pi ** 2) * osm.laplacian(n_cpu, "2d")
    shift = ev1(N0, 1) + ev1(N0, 2)
    P = osm.interpolation(N0, n_cpu).toarray()
    v = np.kron(P @ vec1(N0, 1), P @ vec1(N0, 2))
    v /= np.linalg.norm(v)
    osv.vcycle(np.zeros(n_cpu * n_cpu), v.copy(), H, osm, shift=shift, dimension="2d", lowest_level=lowest)  # builds R,P,RAP
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        osv.vcycle(np.zeros(n_cpu * n_cpu), v.copy(), H, osm, shift=shift, dimension="2d", lowest_level=lowest)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return updates_per_cycle(n_cpu, lowest) / best, best


def defaults(args, multi):
    smoother = args.smoother or ("wjacobi" if multi else "rbgs")
    lowest = args.lowest or (8 if multi else 64)
    return smoother, lowest


def run_reference(args):
    """The reference arm: the CPU implementation of the path on the host cores.  The reference itself is Python 2 and
    cannot be installed or run on this box, so this is the C/OpenMP port of its arithmetic (kind "port") with all host
    threads, on the bench workload of the same --gpus (4096^2; 16384^2 for N > 1): each step is the same step our arm
    times (k V-cycles + Rayleigh sums + modified Gram-Schmidt, all threaded), as many steps as fit the time budget."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    multi = (int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus > 1) and not args.replicas
    N = args.n or (16384 if multi else 4096)      # the same workload as our arm at this --gpus (16384^2 when slab-decomposed)
    smoother, lowest = defaults(args, multi)
    V, shifts = start_block_host(N)
    W = np.zeros_like(V)
    k = len(MODES)
    blk = c_oracle.ShiftBlock(N, lowest, shifts, smoother=smoother)
    bufs = [V, W]

    def step():
        blk.step(bufs[0], bufs[1])
        bufs.reverse()
    budget_s = 150.0
    t_start = time.perf_counter()
    warm = 0
    for _ in range(max(1, min(args.warmup, 3))):
        step(); warm += 1
        if time.perf_counter() - t_start > 0.3 * budget_s:
            break
    done = 0
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        step(); done += 1
        if time.perf_counter() - t_start > budget_s:
            break
    dt = time.perf_counter() - t0
    value = k * updates_per_cycle(N, lowest) * done / dt
    cores = omp_threads()
    sample = ("%d of the requested %d steps (time-bounded); each step = %d V(4,4)-cycles (%s, lowest_level=%d) + Rayleigh sums + "
              "modified Gram-Schmidt at %d^2; matrix-free C/OpenMP port, %d threads" % (done, args.steps, k, smoother, lowest, N, cores))
    cfg = workload_config(N, smoother, lowest, k)
    if multi:
        cfg["l2"] = "per-rank working set >> 126 MB L2"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": warm, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "impl_detail": "CPU port of the reference's arithmetic (oracle/mgcmt_oracle.c); the Python-2 reference itself cannot run here",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
LEG_KINDS = {1: ("down leg, zero start", 18.0), 2: ("down leg", 26.0), 3: ("up leg", 26.0), 4: ("up leg + Rayleigh sums", 26.0)}


def time_steps(torch, run, nsteps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(nsteps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def leg_profile(lib, torch, loop, nserial, n, smoother, peak, traffic_tab, N):
    """serial pass of the same step on ONE stream with CUDA events around every finest-level leg (inside the timed
    region the cycles overlap on several streams, so a bracketed kernel would be timed sharing the GPU)"""
    lib.mgcmt_profile_enable(1)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(nserial):
        loop.step(serial=True)
    es1.record()
    torch.cuda.synchronize()
    ms_serial = es0.elapsed_time(es1) / nserial
    ms_k = (C.c_double * 5)()
    cnt_k = (C.c_longlong * 5)()
    lib.mgcmt_profile_read_kinds(ms_k, cnt_k)
    lib.mgcmt_profile_enable(0)
    kernels = []
    for kind, (label, bpu) in LEG_KINDS.items():
        if cnt_k[kind] == 0:
            continue
        per = ms_k[kind] / cnt_k[kind]
        gbs = bpu * n / (per * 1e-3) / 1e9
        name = ("uni5_leg_kernel<%s>" % ("GS: 8 colour stages" if smoother == "rbgs" else "4 Jacobi sweeps")) + " finest-level " + label
        kernels.append({"kernel": name, "kind": label, "launch_ms": per, "launches_timed": int(cnt_k[kind]),
                        "algorithmic_bytes_per_launch": bpu * n, "bytes_per_unknown": bpu, "achieved": gbs, "frac": gbs / peak,
                        "share_of_step": ms_k[kind] / (ms_serial * nserial),
                        "traffic": traffic_tab.get("%s:%s" % (smoother, label)) if N == 4096 else None})
    return kernels, ms_serial


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        with StdoutToStderr():     # NCCL printf()s its version banner to stdout at the first communicator init
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()

    from multigridcmt_b200 import MGCMTProcessor, MGCMTSolver, MGCMTStencilMaker, ZeroVector, _lib
    from multigridcmt_b200.hierarchy import _ptr, _stream_ptr
    from multigridcmt_b200.eigensolver import ShiftMethod
    lib = _lib.load()
    for kv in filter(None, os.environ.get("MGCMT_OPTIONS", "").split(",")):   # A/B switches: MGCMT_OPTIONS=name=value,...
        name, _, val = kv.partition("=")
        _lib.check(lib.mgcmt_set_option(name.encode(), int(val)))
    sm, solver, proc = MGCMTStencilMaker(), MGCMTSolver(), MGCMTProcessor()

    N = args.n or 4096
    smoother, lowest = defaults(args, False)
    n = N * N
    H = (-1.0 / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
    V_host, shifts = initial_block(N)
    k = len(MODES)
    exact = np.array([ev1(N, a) + ev1(N, b) for a, b in MODES])
    # The timed object is the package's own device-resident outer loop (multigridcmt_b200/eigensolver.py): the k
    # eigenvectors of a step are independent until the Gram-Schmidt, so each V-cycle gets its own CUDA stream and its
    # own hierarchy (level work vectors) -- the latency-bound coarse levels of one cycle overlap the HBM-bound fine
    # levels of another.  With N ranks and --replicas every rank runs the same independent block.
    t_setup = time.perf_counter()
    loop = ShiftMethod(H, shifts, V_host, dimension="2d", lowest_level=lowest, nu1=4, nu2=4, smoother=smoother,
                       ortho=("gram" if args.ortho == "gram" else "mgs"), streams=args.streams)
    nstreams = len(loop.streams)
    rq = loop.rq
    ortho_mode = loop.ortho
    step = loop.step

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        clocks.wait_first()
    step()                       # builds the k coarsest inverses (one per shift)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    # the Gram-matrix orthonormalisation must reproduce column-by-column MGS on this block (parity bar 1e-12)
    ortho_check = None
    if ortho_mode == 2:
        A_ = loop.blocks[1 - loop.cur].clone(); B_ = A_.clone()
        _lib.check(lib.mgcmt_gramschmidt(n, k, _ptr(A_), 1, _stream_ptr(torch)))
        _lib.check(lib.mgcmt_gramschmidt(n, k, _ptr(B_), 2, _stream_ptr(torch)))
        ortho_check = float((A_ - B_).norm() / A_.norm())
        del A_, B_
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = lib.mgcmt_launch_count()
    step(); step()
    launches_per_step = (lib.mgcmt_launch_count() - launches0) // 2
    run, graphed = make_runner(step, torch, not args.no_graph)
    run(2)
    torch.cuda.synchronize()
    t_begin = time.time()
    ms = time_steps(torch, run, args.steps)
    t_end = time.time()
    launches = launches_per_step * args.steps
    lam = (rq[:, 0] / rq[:, 1]).cpu().tolist()   # Rayleigh quotients of the last timed step's V-cycle outputs
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop(t_begin, t_end) if rank == 0 else None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kernels, ms_serial = leg_profile(lib, torch, loop, min(args.steps, 10), n, smoother, peak, TRAFFIC_NCU, N)

    ups_step = k * updates_per_cycle(N, lowest)
    value = world * ups_step * args.steps / (ms * 1e-3)

    # ---- side lines (not the headline): the other smoother / depth, convergence to 1e-10, the strong-scaling base ----
    side, converge, scaling_base = [], None, None
    if world == 1 and not args.no_side:
        for sm_o, low_o in (("wjacobi", 8), ("rbgs", 8), ("wjacobi", 64)):
            if (sm_o, low_o) == (smoother, lowest):
                continue
            lo = ShiftMethod(H, shifts, V_host, dimension="2d", lowest_level=low_o, smoother=sm_o, ortho="gram", streams=args.streams)
            for _ in range(4):
                lo.step()
            run_o, _g = make_runner(lo.step, torch, not args.no_graph)
            run_o(2)
            torch.cuda.synchronize()
            ms_o = time_steps(torch, run_o, 20) / 20
            kern_o, _ = leg_profile(lib, torch, lo, 4, n, sm_o, peak, TRAFFIC_NCU, N)
            side.append({"smoother": sm_o, "lowest_level": low_o, "ms_per_step": ms_o, "vcycles_per_s": k / (ms_o * 1e-3),
                         "value": k * updates_per_cycle(N, low_o) / (ms_o * 1e-3),
                         "legs": [{"kind": kk["kind"], "launch_ms": kk["launch_ms"], "frac": kk["frac"]} for kk in kern_o]})
            del lo, run_o
        # north_star: lowest 4 eigenpairs to 1e-10.  The reference's loop has no stopping rule and, as a
        # power iteration on the V-cycle operator, stagnates at the cycle's own mode mixing; the correction form
        # (ShiftMethod.correction_step, same cycles, right-hand side = eigen-residual) has exact eigenvectors as fixed points.
        floor = 0.5 * np.finfo(float).eps * 8.0 * N * N / np.pi ** 2   # fp64 evaluation floor of ||H v - rho v||: eps/2 * ||H||_inf
        conv = {}
        for form in ("reference", "correction"):
            lc = ShiftMethod(H, shifts, V_host, dimension="2d", lowest_level=lowest, smoother=smoother, ortho="gram", streams=args.streams)
            lc.step(); lc.blocks[lc.cur].copy_(torch.from_numpy(V_host).cuda()); lc.iterations = 0   # coarsest inverses built, block reset
            torch.cuda.synchronize()
            tc = time.perf_counter()
            r = lc.solve(tol=max(1e-10, floor), max_iters=(40 if form == "correction" else 25), form=form, exact=None)
            torch.cuda.synchronize()
            tcs = time.perf_counter() - tc
            hist = r["history"]
            it_eig = next((it for it, res, rho in hist if np.all(np.abs(rho - exact) <= 1e-10)), None)
            conv[form] = {"converged": bool(r["converged"]), "iterations": int(r["iterations"]), "time_ms": 1e3 * tcs,
                          "iters_to_eigenvalues_1e-10": it_eig,
                          "residual_norms": [float(x) for x in r["residual_norms"]],
                          "eigenvalue_abs_err": [float(x) for x in np.abs(r["eigenvalues"] - exact)],
                          "residual_tolerance": max(1e-10, floor)}
            del lc
        converge = {"criterion": "||H v - rho v||_2 <= max(1e-10, eps/2 ||H||_inf) and |rho - closed form| <= 1e-10 (SURVEY D8); "
                                 "at %d^2 ||H||_inf = %.2e, so the fp64 floor of the residual norm is %.1e" % (N, 8.0 * N * N / np.pi ** 2, floor),
                    "reference_form": conv["reference"], "correction_form": conv["correction"]}
        if N == 4096:
            try:   # the 1-GPU time of the multi-GPU workload (16384^2, Jacobi, lowest 8): denominator of the strong-scaling curve
                Nb, lowb = 16384, 8
                Hb = (-1.0 / np.pi ** 2) * sm.laplacian(Nb, "2d", matrix_free=True)
                Vb, shb = initial_block(Nb)
                lb = ShiftMethod(Hb, shb, Vb, dimension="2d", lowest_level=lowb, smoother="wjacobi", ortho="gram", streams=args.streams)
                del Vb
                for _ in range(3):
                    lb.step()
                torch.cuda.synchronize()
                ms_b = time_steps(torch, lambda m: [lb.step() for _ in range(m)], 6) / 6
                scaling_base = {"grid": Nb, "smoother": "wjacobi", "lowest_level": lowb, "n_gpus": 1, "ms_per_step": ms_b,
                                "value": k * updates_per_cycle(Nb, lowb) / (ms_b * 1e-3),
                                "note": "same step as `--gpus N>1` (row slabs) on one GPU: divide the N-GPU ms_per_step into this"}
                del lb
                torch.cuda.empty_cache()
            except Exception as e:   # e.g. out of memory on a smaller part
                scaling_base = {"error": str(e).splitlines()[0] if str(e) else type(e).__name__}

    # ---- e2e: the reference-facing call with host buffers, copies inside the timed region -----
    e2e = None
    if args.e2e_steps > 0:
        Vh = torch.from_numpy(V_host.copy()).pin_memory()
        Vnp = Vh.numpy()
        Vpage = V_host.copy()
        zero = ZeroVector(n)
        kw = {"smoother": solver.rbgs} if smoother == "rbgs" else {}

        def e2e_step(src, v0):
            # one call per eigenvector, as a user of the reference writes it (2DPotGS.py:95), host arrays in, host array out;
            # what the caller then does with w on the host (numpy normalisation) is not part of the path
            for c in range(k):
                w = solver.vcycle(v0, src[c], H, sm, shift=shifts[c], dimension="2d", lowest_level=lowest, **kw)
                assert w.shape == (n,)

        def e2e_block(src, v0):
            # the same loop body as ONE call: the k cycles are independent, so the upload of vector c+1 and the download
            # of vector c-1 ride beside cycle c (MGCMTSolver.vcycle_many -> mgcmt_vcycle_host_block)
            ws = solver.vcycle_many([src[c] for c in range(k)], H, sm, shifts, dimension="2d", lowest_level=lowest, **kw)
            assert len(ws) == k and ws[0].shape == (n,)

        def timed(src, v0, fn=None):
            fn = fn or e2e_step
            fn(src, v0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                fn(src, v0)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return world * ups_step * args.e2e_steps / dt
        v_block = timed(Vnp, zero, e2e_block)
        v_block_page = timed(Vpage, zero, e2e_block)
        v_pinned = timed(Vnp, zero)
        v_page = timed(Vpage, zero)
        zeros_np = np.zeros(n)

        def with_zero_upload():
            r = timed(Vnp, zeros_np)
            zeros_np.shape = (n,)
            return r
        v_zero_upload = with_zero_upload()
        # what the copies alone cost: one pinned H2D + one D2H of a vector per call (the PCIe floor of this API)
        dbuf = torch.empty(n, dtype=torch.float64, device="cuda")
        hout = torch.empty(n, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in range(k):
            dbuf.copy_(Vh[c], non_blocking=True)
            hout.copy_(dbuf, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        copy_s = time.perf_counter() - t0
        del dbuf, hout
        e2e = {"value": v_block, "unit": UNIT, "h2d_bytes_per_step": k * n * 8, "d2h_bytes_per_step": k * n * 8,
               "one_call_per_vector": v_pinned, "block_call_pageable_f": v_block_page,
               "copies_only": {"ms_per_step": 1e3 * copy_s, "GBs_h2d_plus_d2h": k * n * 8 / copy_s / 1e9 * 2,
                               "value_if_compute_were_free": world * ups_step / copy_s,
                               "note": "a one-vector call is synchronous (numpy in, numpy out), so its upload, cycle and download "
                                       "cannot overlap: this is the PCIe floor of one_call_per_vector; the block call overlaps the "
                                       "two directions and is bounded by one direction's %.0f MB per step instead" % (k * n * 8 / 1e6)},
               "steps": args.e2e_steps,
               "call": "MGCMTSolver.vcycle_many([numpy f_c (pinned)], H, sm, shifts, dimension='2d', lowest_level=%d%s): the k "
                       "zero-start cycles of a step in one call, copies of neighbouring vectors pipelined around each cycle "
                       "(mgcmt_vcycle_host_block); one_call_per_vector = MGCMTSolver.vcycle(ZeroVector(n), f_c, ...) k times"
                       % (lowest, ", smoother=solver.rbgs" if smoother == "rbgs" else ""),
               "pageable_f": v_page, "with_numpy_zero_v0_upload": v_zero_upload}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roofline = None
    if kernels:
        dom = max(kernels, key=lambda kk: kk["share_of_step"])
        roofline = {"bound": "hbm", "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                    "traffic": dom["traffic"], "kernel": dom["kernel"], "launch_ms": dom["launch_ms"],
                    "launches_timed": dom["launches_timed"], "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                    "peak_source": peak_kind, "share_of_step": dom["share_of_step"],
                    "timed_in": "serial pass of %d steps after the timed region (%.3f ms/step on one stream), CUDA events on the launching stream"
                                % (min(args.steps, 10), ms_serial),
                    "kernels": kernels,
                    "algorithmic_speedup": {"unfused_bytes_per_unknown_and_cycle": 304.0,
                                            "step_GBs_at_304B": 304.0 * n * k * args.steps / (ms * 1e-3) / 1e9,
                                            "note": "SURVEY 8(d)'s figure for one kernel per operator; the fused legs move 18-26 B per "
                                                    "unknown and leg instead of 114 B, which is a traffic saving, not bandwidth"}}

    cpu_baseline = None
    if not args.no_cpu and world == 1:
        v_c, t_c, done_c = cpu_block_sample(N, smoother, lowest, steps=4)
        cpu_baseline = {"value": v_c, "unit": UNIT, "cores": omp_threads(), "kind": "port",
                        "sample": "%d shift-method steps (%d V(4,4)-cycles each, %s, lowest_level=%d, + Rayleigh sums + modified "
                                  "Gram-Schmidt) of the %d^2 workload: %.2f s with the matrix-free C/OpenMP port on %d threads"
                                  % (done_c, k, smoother, lowest, N, t_c, omp_threads())}
        if not args.no_side:
            v_cpu, t_cpu = cpu_sample(args.cpu_n, 8)
            cpu_baseline["scipy_port_1core"] = {"value": v_cpu, "sample": "1 V(4,4)-cycle (wjacobi, lowest_level=8) at %d^2, hierarchy prebuilt, "
                                                "%.2f s, scipy CSC SpMV (the port pinned to the real reference; single-threaded)" % (args.cpu_n, t_cpu)}

    cfg = workload_config(N, smoother, lowest, k)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "impl_detail": {"parallelism": "replicas x%d" % world if world > 1 else "1 GPU", "streams": nstreams, "cuda_graph": graphed,
                        "ortho": ("Gram-matrix form of the Gram-Schmidt step (12 instead of 29 vector passes); rel. difference to "
                                  "column-by-column MGS on this block: %.1e" % ortho_check) if ortho_mode == 2 else "column-by-column MGS",
                        "setup_s": setup_s, "setup": "hierarchies + %d coarsest inverses (%d unknowns each, banded LU above 256)" % (k, lowest * lowest),
                        "scaling_note": "--gpus 1 runs BASELINE config 3 (4096^2); --gpus N>1 runs config 5 (16384^2 in row slabs): "
                                        "use scaling_base as the 1-GPU point of the strong-scaling curve"},
        "vcycles_per_s": world * k * args.steps / (ms * 1e-3),
        "eigenvalues": lam, "eigenvalue_abs_err": [abs(a - b) for a, b in zip(lam, exact)],
        "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "side_lines": side, "converge": converge, "scaling_base": scaling_base,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



def make_runner(step, torch, use_graph):
    """Returns (run(nsteps), graphed?).  With CUDA graphs, two consecutive steps (the block buffers swap roles every
    step) are captured once and replayed: the ~100 kernel launches / NCCL calls of a step then cost no CPU time."""
    if not use_graph:
        return (lambda n: [step() for _ in range(n)]), False
    try:
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            step(); step()                      # warm the capture stream (allocator, NCCL channels)
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=cap):
            step(); step()
        torch.cuda.synchronize()

        def run(n):
            for _ in range(n // 2):
                g.replay()
            if n % 2:
                step()
        return run, True
    except Exception as e:      # capture not possible (e.g. an NCCL build without graph support): eager launches
        sys.stderr.write("bench.py: CUDA graph capture failed (%s); running eagerly\n" % (str(e).splitlines()[0] if str(e) else type(e).__name__))
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        return (lambda n: [step() for _ in range(n)]), False

def run_slab(args):
    """--gpus N > 1: 2-D well 16384^2, row-slab decomposed over N B200s (one rank per GPU), NCCL halo exchange
    per fused leg, coarse levels (<= 2048^2) replicated after an all-gather.  Same step as the single-GPU arm."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    with StdoutToStderr():         # NCCL printf()s its version banner to stdout at the first communicator init
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from multigridcmt_b200 import MGCMTStencilMaker, _lib
    from multigridcmt_b200.slab import HALO, SlabVCycle, TorchDistComm
    lib = _lib.load()
    N = args.n or 16384
    smoother_s, lowest = defaults(args, True)
    k = len(MODES)
    sm = MGCMTStencilMaker()
    H = (-1.0 / np.pi ** 2) * sm.laplacian(N, "2d", matrix_free=True)
    native = (args.slab_driver == "native" and args.ortho == "gram")
    P = interp1(N)
    shifts = [ev1(N0, a) + ev1(N0, b) for a, b in MODES]
    lam = torch.zeros(k, 2, dtype=torch.float64, device="cuda")
    svs = []
    if native:
        # the whole step is two C calls: k lock-step cycles (+ Rayleigh sums) and the Gram-form orthonormalisation
        from multigridcmt_b200.slab import NativeSlabBlock
        with StdoutToStderr():
            nb = NativeSlabBlock(H, world, rank, k, lowest_level=lowest, gather_cols=args.gather_cols,
                                 stagger=bool(args.slab_stagger), smoother=smoother_s)
        own, begin, nlev_slab, lockstep = nb.own0, nb.begin0, nb.nlev, True
        owned = nb.owned
        bd = {"V": nb.new_block(), "W": nb.new_block()}

        def step():
            nb.cycle(shifts, bd["V"], bd["W"], lam)
            nb.gram(bd["W"])
            bd["V"], bd["W"] = bd["W"], bd["V"]

        def one_cycle_all():      # e2e: the k cycles of a step without the orthonormalisation
            nb.cycle(shifts, bd["V"], bd["W"], None)
    else:
        comm = TorchDistComm()
        # one solver state (level buffers) per CUDA stream: the 4 independent V-cycles of a step overlap each other's
        # halo exchanges and replicated coarse parts
        nstreams = max(1, min(args.streams, k))
        svs = [SlabVCycle(H, world, comm, [rank], lowest_level=lowest, gather_cols=args.gather_cols, smoother=smoother_s)
               for _ in range(nstreams)]
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        sv = svs[0]
        st = sv.states[0]
        own, begin, nlev_slab = st.own0, st.begin0, sv.nlev
        owned = lambda x: st.owned(x, 0)
        bd = {"V": sv.new_block(k)[0], "W": sv.new_block(k)[0]}       # (k, slab_size) blocks in slab layout
        from multigridcmt_b200.slab import vcycle_block
        lockstep = (nstreams == k)

        def cycles(with_lam):
            V = [[bd["V"][c]] for c in range(k)]
            W = [[bd["W"][c]] for c in range(k)]
            if lockstep:
                # the k cycles advance level by level together: every halo-exchange phase is one NCCL group for all k
                vcycle_block(svs, shifts, V, W, lam=[lam] if with_lam else None, streams=streams)
            else:
                for c in range(k):
                    svc = svs[c % nstreams]
                    svc.vcycle(shifts[c], v0_is_zero=True, f0=V[c], v0=W[c])
                    if with_lam:
                        svc.rayleigh(W[c], sync=False)
                        lam[c].copy_(svc.states[0].scal[:2])
            return W

        def step():
            W = cycles(True)
            if args.ortho == "gram":
                sv.gramschmidt_gram([bd["W"]])
            else:
                sv.gramschmidt(W)
            bd["V"], bd["W"] = bd["W"], bd["V"]

        def one_cycle_all():
            cycles(False)
    for c, (a, b) in enumerate(MODES):
        ya, yb = P @ vec1(N0, a), P @ vec1(N0, b)
        blk = np.outer(ya[begin:begin + own], yb) / (np.linalg.norm(ya) * np.linalg.norm(yb))
        owned(bd["V"][c]).copy_(torch.from_numpy(blk).cuda())

    clocks = Clocks(local)
    if rank == 0:
        clocks.start(); clocks.wait_first()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    lib.mgcmt_profile_enable(0)
    launches0 = lib.mgcmt_launch_count()
    step(); step()
    launches_per_step = (lib.mgcmt_launch_count() - launches0) // 2
    # eager launches on the slab path: capturing the NCCL send/recv groups of the halo exchange into a CUDA graph
    # hung on this stack (torch 2.11 / NCCL 2.28), so the step is launched call by call here
    run, graphed = make_runner(step, torch, False)
    run(2)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    run(args.steps)
    e1.record()
    host_issue_ms = (time.time() - t_begin) * 1e3 / args.steps   # CPU time to enqueue a step (no sync inside)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = launches_per_step * args.steps
    clk = clocks.stop(t_begin, t_end) if rank == 0 else None
    lam_h = (lam[:, 0] / lam[:, 1]).cpu().tolist()
    exact = [ev1(N, a) + ev1(N, b) for a, b in MODES]
    ups_step = k * updates_per_cycle(N, lowest)
    value = ups_step * args.steps / (ms * 1e-3)

    # e2e: owned rows of f come from pinned host memory, the owned rows of the result go back, every cycle
    e2e = None
    if args.e2e_steps > 0:
        hostf = torch.empty(k, own * N, dtype=torch.float64).pin_memory()
        hostv = torch.empty(k, own * N, dtype=torch.float64).pin_memory()
        for c in range(k):
            hostf[c].copy_(owned(bd["V"][c]).reshape(-1))
        def e2e_step():
            for c in range(k):
                owned(bd["V"][c]).copy_(hostf[c].view(own, N), non_blocking=True)
            one_cycle_all()
            for c in range(k):
                hostv[c].view(own, N).copy_(owned(bd["W"][c]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_step()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": ups_step * args.e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": k * own * N * 8 * world, "d2h_bytes_per_step": k * own * N * 8 * world,
               "steps": args.e2e_steps, "call": "slab block cycle (k V-cycles) on owned rows copied from / to pinned host memory"}
    LEG_B = (18.0 + 26.0) * 4.0 / 3.0
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(N, smoother_s, lowest, k), l2="per-rank working set >> 126 MB L2"),
            "impl_detail": {"parallelism": "row slabs x%d, NCCL send/recv halos per fused leg + all-gather of the first replicated level" % world,
                            "slab_levels": nlev_slab, "halo_rows": HALO, "lockstep_block": lockstep, "cuda_graph": graphed,
                            "driver": (("native: step issued from C++, NCCL called directly (csrc/slab_block.cu)"
                                        + ("; two halves half a phase apart on two communicators (one half's exchange overlaps the other's kernels)"
                                           if args.slab_stagger else "; lock-step")) if native else "python: torch.distributed"),
                            "replicated_from": "%d^2" % (N >> nlev_slab),
                            "l2": "per-rank working set %.1f GB >> 126 MB L2" % (12 * (own + 2 * HALO) * N * 8 / 1e9),
                            "strong_scaling_base": "the --gpus 1 line carries the 1-GPU time of this workload as scaling_base"},
            "vcycles_per_s": k * args.steps / (ms * 1e-3), "host_issue_ms_per_step": host_issue_ms,
            "eigenvalues": lam_h, "eigenvalue_abs_err": [abs(a - b) for a, b in zip(lam_h, exact)],
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": LEG_B * N * N * k * args.steps / (ms * 1e-3) / 1e9 / world, "peak": peak,
                         "unit": "GB/s", "frac": LEG_B * N * N * k * args.steps / (ms * 1e-3) / 1e9 / world / peak, "traffic": None,
                         "kernel": "whole step per GPU at %.1f B per fine unknown and V-cycle = the fused legs' compulsory traffic "
                                   "(18 B zero-start down leg + 26 B up leg per level, x 4/3 over the levels; halo exchanges, the "
                                   "replicated coarse part and the orthonormalisation are inside the time, not the bytes); "
                                   "per-kernel numbers: N=1 run" % LEG_B,
                         "unfused_step_GBs_at_304B": 304.0 * N * N * k * args.steps / (ms * 1e-3) / 1e9 / world,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback"},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    for s_ in svs:
        s_.close()
    if native:
        nb.close()
    dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1 and not args.replicas:
        run_slab(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
