"""Row-slab decomposition of the 2-D V-cycle across GPUs.

One process per GPU.  Rank r owns the rows [r N/G, (r+1) N/G) of the finest grid and the matching rows of
every *distributed* level; each slab array carries HALO = 10 rows above and below its owned rows.  A V(4,4)
leg is ONE fused kernel per level (csrc/fused.cu: 4 sweeps + transfer), whose dependency cone is exactly
those 6 rows -- so halos are exchanged once per leg, not once per sweep:

    down, level l:   leg(v_l, f_l) -> tmp_l, f_{l+1}(owned rows);   exchange halos of f_{l+1}
    coarse part:     levels narrower than `gather_cols` are REPLICATED: the restricted residual of the last slab
                     level is written straight into a full-size array, all-gathered (NCCL over NVLink), and every
                     rank runs the same small V-cycle (mgcmt_vcycle_from) -- no scatter back is needed
    up, level l:     exchange halos of tmp_l and of the coarse correction;   leg(tmp_l + P e, f_l) -> v_l

Slab cuts sit on multiples of 2^(number of slab levels), so coarse row j <-> fine rows 2j, 2j+1, 2j+2 stays aligned on
every level (the reference's coarse point sits on the odd fine point, MGCMTStencilMaker.py:39-42).

The communicator is abstract: `TorchDistComm` uses torch.distributed (NCCL on GPUs, gloo in the CPU tests of the
exchange plumbing); `LocalComm` keeps all ranks in one process and copies between their arrays, which lets a single GPU
run -- and check bit-for-bit against the undecomposed V-cycle -- the exact kernels and halo logic of the multi-GPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .operators import SeparableOperator

HALO = 10   # rows: the dependency cone of the longest fused leg (8 Gauss-Seidel colour stages + transfer = 10; Jacobi: 6)
MODE_SMOOTH, MODE_DOWN, MODE_DOWN_ZERO, MODE_UP = 0, 1, 2, 3


def plan_levels(n, world, gather_cols=2048, min_own_rows=64):
    """Number of slab (distributed) levels for an n x n grid on `world` ranks: levels wider than gather_cols,
    while every rank still owns >= min_own_rows rows and the cuts stay even on every slab level."""
    if n % world:
        raise ValueError("grid rows must divide evenly among the ranks")
    own = n // world
    nlev = 0
    while (n >> nlev) > gather_cols and (own >> nlev) >= min_own_rows and own % (1 << (nlev + 1)) == 0:
        nlev += 1
    return nlev


# ----------------------------------------------------------------------------------------------------
# halo exchange plumbing (device-agnostic: works on any torch tensors, so gloo/CPU can test it)
# ----------------------------------------------------------------------------------------------------
def halo_views(x, own_rows, ncols, halo=HALO):
    """(top_halo, top_owned, bottom_owned, bottom_halo) row-block views of a slab array of (own + 2 halo) rows."""
    a = x.view(own_rows + 2 * halo, ncols)
    return a[:halo], a[halo:2 * halo], a[own_rows:own_rows + halo], a[own_rows + halo:]


class LocalComm:
    """All ranks live in this process (their slab arrays are passed in as lists)."""

    def __init__(self, world):
        self.world = world

    def exchange(self, arrays, own_rows, ncols):
        for r in range(self.world):
            _, top_own, bot_own, _ = halo_views(arrays[r], own_rows, ncols)
            if r > 0:
                halo_views(arrays[r - 1], own_rows, ncols)[3].copy_(top_own)
            if r + 1 < self.world:
                halo_views(arrays[r + 1], own_rows, ncols)[0].copy_(bot_own)

    def exchange_many(self, items):
        for (arrays, own_rows, ncols) in items:
            self.exchange(arrays, own_rows, ncols)

    def allgather_rows(self, fulls, own_rows, ncols):
        """every rank's full array gets every rank's owned row block"""
        for src in range(self.world):
            blk = fulls[src].view(-1, ncols)[src * own_rows:(src + 1) * own_rows]
            for dst in range(self.world):
                if dst != src:
                    fulls[dst].view(-1, ncols)[src * own_rows:(src + 1) * own_rows].copy_(blk)

    def allreduce_sum(self, scalars):
        tot = sum(scalars[1:], scalars[0].clone())
        for s in scalars:
            s.copy_(tot)


class TorchDistComm:
    """One rank per process over torch.distributed (NCCL: send/recv pairs over NVLink; gloo for CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def exchange(self, arrays, own_rows, ncols):
        self.exchange_many([(arrays, own_rows, ncols)])

    def exchange_many(self, items):
        """items: [(arrays, own_rows, ncols), ...] -- all halo exchanges of one phase in ONE NCCL group call"""
        dist = self.dist
        ops = []
        for (arrays, own_rows, ncols) in items:
            (x,) = arrays
            top_halo, top_own, bot_own, bot_halo = halo_views(x, own_rows, ncols)
            if self.rank > 0:
                ops += [dist.P2POp(dist.isend, top_own, self.rank - 1, self.group),
                        dist.P2POp(dist.irecv, top_halo, self.rank - 1, self.group)]
            if self.rank + 1 < self.world:
                ops += [dist.P2POp(dist.isend, bot_own, self.rank + 1, self.group),
                        dist.P2POp(dist.irecv, bot_halo, self.rank + 1, self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def allgather_rows(self, fulls, own_rows, ncols):
        (full,) = fulls
        mine = full.view(-1, ncols)[self.rank * own_rows:(self.rank + 1) * own_rows].clone()
        self.dist.all_gather_into_tensor(full.view(-1), mine.view(-1), group=self.group)

    def allreduce_sum(self, scalars):
        (s,) = scalars
        self.dist.all_reduce(s, group=self.group)


# ----------------------------------------------------------------------------------------------------
# per-rank device state
# ----------------------------------------------------------------------------------------------------
class _RankState:
    def __init__(self, op, rank, world, nlev_slab, lowest_level):
        torch = _lib.require_cuda()
        lib = _lib.load()
        N = op.ncols
        self.rank, self.world, self.N = rank, world, N
        self.own0 = N // world
        self.begin0 = rank * self.own0
        self.nlev = nlev_slab
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.slab = C.c_void_p()
        _lib.check(lib.mgcmt_hier_create_slab(C.byref(self.slab), N, N, self.begin0, self.own0, nlev_slab, HALO,
                                              hp(op.row[0]), hp(op.row[1]), hp(op.row[2]),
                                              hp(op.col[0]), hp(op.col[1]), hp(op.col[2]), stream))
        # replicated coarse part: full hierarchy whose first nlev_slab levels carry operators only
        self.coarse = C.c_void_p()
        _lib.check(lib.mgcmt_hier_create2(C.byref(self.coarse), N, N, 1, hp(op.row[0]), hp(op.row[1]), hp(op.row[2]),
                                          hp(op.col[0]), hp(op.col[1]), hp(op.col[2]), int(lowest_level), nlev_slab, stream))
        f64 = torch.float64
        self.v, self.f, self.tmp = [], [], []
        for l in range(nlev_slab):
            n = ((self.own0 >> l) + 2 * HALO) * (N >> l)
            self.v.append(torch.zeros(n, dtype=f64, device="cuda"))
            self.f.append(torch.zeros(n, dtype=f64, device="cuda"))
            self.tmp.append(torch.zeros(n, dtype=f64, device="cuda"))
        ng = (N >> nlev_slab) ** 2
        self.fg = torch.zeros(ng, dtype=f64, device="cuda")   # full (replicated) coarse right-hand side
        self.vg = torch.zeros(ng, dtype=f64, device="cuda")   # full coarse correction
        self.scal = torch.zeros(32, dtype=f64, device="cuda")

    def close(self):
        lib = _lib.load()
        for h in (self.slab, self.coarse):
            if h:
                lib.mgcmt_hier_destroy(h)
        self.slab = self.coarse = None

    def own_rows(self, l):
        return self.own0 >> l

    def ncols(self, l):
        return self.N >> l

    def owned(self, x, l):
        """view of the owned rows of a slab array of level l"""
        return x.view(self.own_rows(l) + 2 * HALO, self.ncols(l))[HALO:HALO + self.own_rows(l)]


class SlabVCycle:
    """V(4,4) weighted-Jacobi cycles on a row-slab decomposed N x N grid.

    states: one _RankState per rank handled by THIS process (all ranks with LocalComm, one with TorchDistComm)."""

    def __init__(self, op: SeparableOperator, world, comm, ranks, lowest_level=8, gather_cols=2048, omega=None,
                 smoother="wjacobi"):
        """smoother: "wjacobi" (omega default 2/3) or "rbgs" (red-black / four-colour Gauss-Seidel, omega default 1: BASELINE
        config 3's smoother).  Red-black legs: one pass of 4 sweeps on the 5-point level, two passes of 2 sweeps on the
        9-point levels with a halo exchange of the intermediate iterate between them."""
        if op.nrows != op.ncols:
            raise ValueError("slab decomposition is for square 2-D grids")
        if smoother not in ("wjacobi", "rbgs"):
            raise ValueError("slab smoother must be wjacobi or rbgs")
        self.op, self.world, self.comm = op, world, comm
        self.gs = (smoother == "rbgs")
        self.omega = (1.0 if self.gs else 2. / 3.) if omega is None else omega
        self.nlev = plan_levels(op.ncols, world, gather_cols)
        if self.nlev < 1:
            raise ValueError("grid too small to decompose over %d ranks (use the single-GPU path)" % world)
        self.states = [_RankState(op, r, world, self.nlev, lowest_level) for r in ranks]

    def close(self):
        for s in self.states:
            s.close()

    # -- helpers ------------------------------------------------------------------------------------
    def _leg(self, st, l, mode, vin, f, vout, e=None, rc=None, nu=4):
        torch = _lib.require_cuda()
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        if self.gs:
            mode |= 32
        _lib.check(_lib.load().mgcmt_fused_leg(st.slab, l, mode, nu, float(self.shift), float(self.omega), p(vin), p(f),
                                               p(vout), p(e), p(rc), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def _exchange(self, name, l):
        arrs = [getattr(st, name)[l] for st in self.states]
        st0 = self.states[0]
        self.comm.exchange(arrs, st0.own_rows(l), st0.ncols(l))

    def new_vector(self):
        """one finest-level slab array (owned rows + halos, zero) per local rank"""
        torch = _lib.require_cuda()
        return [torch.zeros_like(st.v[0]) for st in self.states]

    # -- the cycle ------------------------------------------------------------------------------------
    def vcycle(self, shift, v0_is_zero=True, f0=None, v0=None):
        """In: the owned rows of f (and of v unless v0_is_zero).  Out: the owned rows of v (halo rows are scratch).
        f0 / v0: per-local-rank finest-level slab arrays to use instead of st.f[0] / st.v[0] (so a block of
        eigenvectors kept in slab layout needs no staging copies); f0's halo rows are refreshed in place."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        self.shift = shift
        nl = self.nlev
        saved = None
        if f0 is not None or v0 is not None:
            saved = [(st.f[0], st.v[0]) for st in self.states]
            for i, st in enumerate(self.states):
                if f0 is not None:
                    st.f[0] = f0[i]
                if v0 is not None:
                    st.v[0] = v0[i]
        try:
            self._vcycle_impl(shift, v0_is_zero)
        finally:
            if saved is not None:
                for st, (f, v) in zip(self.states, saved):
                    st.f[0], st.v[0] = f, v

    def _vcycle_impl(self, shift, v0_is_zero):
        torch = _lib.require_cuda()
        lib = _lib.load()
        nl = self.nlev
        gs = self.gs
        st0 = self.states[0]
        self._exchange("f", 0)
        if not v0_is_zero:
            self._exchange("v", 0)
        cur = [None] * nl     # name of the array that holds the smoothed iterate of level l after its down leg
        for l in range(nl):
            last = (l + 1 == nl)
            zero = (l > 0 or v0_is_zero)
            if gs and l > 0:
                # 9-point level, red-black: 2 sweeps from zero, exchange, 2 sweeps + residual + restriction
                for st in self.states:
                    st.v[l].zero_()
                    self._leg(st, l, MODE_SMOOTH, st.v[l], st.f[l], st.tmp[l], nu=2)
                self._exchange("tmp", l)
                for st in self.states:
                    self._leg(st, l, MODE_DOWN, st.tmp[l], st.f[l], st.v[l], rc=(st.fg if last else st.f[l + 1]), nu=2)
                cur[l] = "v"
            else:
                for st in self.states:
                    mode = MODE_DOWN_ZERO if zero else MODE_DOWN
                    self._leg(st, l, mode, None if zero else st.v[l], st.f[l], st.tmp[l],
                              rc=(st.fg if last else st.f[l + 1]))
                cur[l] = "tmp"
            if last:
                self.comm.allgather_rows([st.fg for st in self.states], st0.own_rows(nl), st0.ncols(nl))
            else:
                self._exchange("f", l + 1)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for st in self.states:   # replicated coarse part (identical on every rank)
            _lib.check(lib.mgcmt_vcycle_from(st.coarse, nl, float(shift), _lib.SMOOTH_RBGS if gs else _lib.SMOOTH_WJACOBI,
                                             float(self.omega), C.c_void_p(st.vg.data_ptr()), C.c_void_p(st.fg.data_ptr()), stream))
        for l in range(nl - 1, -1, -1):
            last = (l + 1 == nl)
            # halos of the smoothed iterate and (below the last slab level) of the coarse correction: one NCCL group
            items = [([getattr(st, cur[l])[l] for st in self.states], st0.own_rows(l), st0.ncols(l))]
            if not last:
                items.append(([st.v[l + 1] for st in self.states], st0.own_rows(l + 1), st0.ncols(l + 1)))
            self.comm.exchange_many(items)
            if gs and l > 0:
                for st in self.states:   # iterate is in v[l]: correction + 2 sweeps -> tmp[l]; exchange; 2 sweeps -> v[l]
                    self._leg(st, l, MODE_UP, st.v[l], st.f[l], st.tmp[l], e=(st.vg if last else st.v[l + 1]), nu=2)
                self._exchange("tmp", l)
                for st in self.states:
                    self._leg(st, l, MODE_SMOOTH, st.tmp[l], st.f[l], st.v[l], nu=2)
            else:
                for st in self.states:
                    self._leg(st, l, MODE_UP, st.tmp[l], st.f[l], st.v[l], e=(st.vg if last else st.v[l + 1]))

    def rayleigh(self, x=None, sync=True):
        """x^T H x and x^T x of a finest-level slab vector (default st.v[0]); its halo rows are refreshed first.
        sync=True returns (quotient, x^T x) as floats; sync=False leaves [num, den] in st.scal[:2] on the device."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        arrs = x if x is not None else [st.v[0] for st in self.states]
        st0 = self.states[0]
        self.comm.exchange(arrs, st0.own_rows(0), st0.ncols(0))
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for st, a in zip(self.states, arrs):
            _lib.check(lib.mgcmt_slab_rayleigh(st.slab, 0, C.c_void_p(a.data_ptr()),
                                               C.c_void_p(st.scal.data_ptr()), stream))
        self.comm.allreduce_sum([st.scal[:2] for st in self.states])
        if not sync:
            return None
        num, den = self.states[0].scal[:2].cpu().tolist()
        return num / den, den

    def new_block(self, k):
        """k finest-level slab vectors in one (k, slab_size) tensor per local rank (equal stride: what the Gram-matrix
        kernels want); block[i][c] is vector c of local rank i"""
        torch = _lib.require_cuda()
        return [torch.zeros(k, st.v[0].numel(), dtype=torch.float64, device="cuda") for st in self.states]

    def gramschmidt_gram(self, blocks):
        """Orthonormalise the k vectors of a block (new_block layout) in Gram-matrix / Cholesky-QR form: local packed
        Gram matrix of the owned rows -> one all-reduce of k(k+1)/2 scalars -> Q = W R^-1 locally."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        k = blocks[0].shape[0]
        ng = k * (k + 1) // 2
        for st, blk in zip(self.states, blocks):
            off = HALO * st.ncols(0)
            n_own = st.own_rows(0) * st.ncols(0)
            base = C.c_void_p(blk.data_ptr() + 8 * off)
            _lib.check(lib.mgcmt_gram(n_own, k, base, blk.shape[1], C.c_void_p(st.scal.data_ptr()), stream))
        self.comm.allreduce_sum([st.scal[:ng] for st in self.states])
        for st, blk in zip(self.states, blocks):
            off = HALO * st.ncols(0)
            n_own = st.own_rows(0) * st.ncols(0)
            base = C.c_void_p(blk.data_ptr() + 8 * off)
            _lib.check(lib.mgcmt_cholqr_apply(n_own, k, base, blk.shape[1], C.c_void_p(st.scal.data_ptr()), stream))
        return blocks

    def gramschmidt(self, block):
        """Modified Gram-Schmidt (MGCMTProcessor.py:44-50) of k slab vectors; block[c] = per-local-rank arrays.
        Dots run over the owned rows of each rank and are all-reduced (one scalar vector per column)."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        k = len(block)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        own = [st.owned(block[0][i], 0).numel() for i, st in enumerate(self.states)]
        views = [[st.owned(block[c][i], 0).reshape(-1) for i, st in enumerate(self.states)] for c in range(k)]
        for i in range(k):
            for r, st in enumerate(self.states):
                _lib.check(lib.mgcmt_dot(own[r], p(views[i][r]), p(views[i][r]), p(st.scal[0:1]), stream))
            self.comm.allreduce_sum([st.scal[0:1] for st in self.states])
            for r, st in enumerate(self.states):                   # q_i = w_i / ||w_i||
                _lib.check(lib.mgcmt_scale_inv_norm(own[r], p(views[i][r]), p(st.scal[0:1]), stream))
            if i + 1 == k:
                break
            for r, st in enumerate(self.states):                   # <q_i,q_i> and <w_j,q_i>, j > i
                _lib.check(lib.mgcmt_dot(own[r], p(views[i][r]), p(views[i][r]), p(st.scal[0:1]), stream))
                for j in range(i + 1, k):
                    _lib.check(lib.mgcmt_dot(own[r], p(views[j][r]), p(views[i][r]), p(st.scal[j - i:j - i + 1]), stream))
            self.comm.allreduce_sum([st.scal[:k - i] for st in self.states])
            for r, st in enumerate(self.states):
                st.scal[1:k - i].div_(st.scal[0])                  # <w_j,q_i> / <q_i,q_i> (MGCMTProcessor.py:10-20)
                for j in range(i + 1, k):
                    _lib.check(lib.mgcmt_axpy_dev(own[r], p(st.scal[j - i:j - i + 1]), -1.0, p(views[i][r]), p(views[j][r]), stream))
        return block

    def scale_v(self, alpha):
        for st in self.states:
            st.v[0].mul_(alpha)

    # -- host <-> slab helpers (tests, bench) ----------------------------------------------------------
    def scatter(self, name, full_host):
        """full N*N numpy vector -> owned rows of st.<name>[0] on every local rank"""
        torch = _lib.require_cuda()
        N = self.op.ncols
        a = np.asarray(full_host, dtype=np.float64).reshape(N, N)
        for st in self.states:
            blk = torch.from_numpy(np.ascontiguousarray(a[st.begin0:st.begin0 + st.own0])).cuda()
            st.owned(getattr(st, name)[0], 0).copy_(blk)

    def gather_local(self, name):
        """owned rows of all LOCAL ranks stacked (LocalComm: the full vector)"""
        torch = _lib.require_cuda()
        return torch.cat([st.owned(getattr(st, name)[0], 0).reshape(-1) for st in self.states]).cpu().numpy()


# ----------------------------------------------------------------------------------------------------
# block of independent cycles in lock-step: one NCCL group per phase for ALL vectors
# ----------------------------------------------------------------------------------------------------
def vcycle_block(svs, shifts, f0s, v0s, lam=None, streams=None):
    """k independent V(4,4) cycles (zero initial guess), one SlabVCycle (= one set of level buffers) per vector, advanced
    level by level together so that every halo-exchange phase is ONE batched NCCL send/recv group for all k vectors
    (the exchanges are latency-bound: 6 rows each).  svs[c] solves (H - shifts[c]) w = f0s[c] into v0s[c]
    (f0s[c] / v0s[c]: per-local-rank finest-level slab arrays).  If `lam` (k x 2 device tensor per local rank, list) is
    given, the Rayleigh numerators / denominators w^T H w, w^T w of the results are left in it (all-reduced).
    streams: optional list of k CUDA streams -- within a phase vector c's kernels run on streams[c] (forked from /
    joined to the current stream around every phase), so the latency-bound legs of the small slab levels and the
    replicated coarse parts of different vectors overlap; the exchanges stay on the current stream."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    k = len(svs)
    main = torch.cuda.current_stream()

    class _On:   # run vector c's launches on its stream, ordered after everything already on the main stream
        def __init__(self, c):
            self.c = c
        def __enter__(self):
            if streams is not None:
                streams[self.c].wait_stream(main)
                self.ctx = torch.cuda.stream(streams[self.c])
                self.ctx.__enter__()
        def __exit__(self, *a):
            if streams is not None:
                self.ctx.__exit__(*a)

    def join():
        if streams is not None:
            for st_ in streams:
                main.wait_stream(st_)

    sv0 = svs[0]
    comm, nl, st0 = sv0.comm, sv0.nlev, sv0.states[0]
    if sv0.gs:
        # red-black legs: the lock-step phase table lives in the native driver (NativeSlabBlock); here one cycle at a time
        for c, sv in enumerate(svs):
            sv.vcycle(shifts[c], v0_is_zero=True, f0=f0s[c], v0=v0s[c])
            if lam is not None:
                sv.rayleigh(v0s[c], sync=False)
                for i, st in enumerate(sv.states):
                    lam[i][c].copy_(st.scal[:2])
        return
    saved = []
    for c, sv in enumerate(svs):
        sv.shift = shifts[c]
        saved.append([(st.f[0], st.v[0]) for st in sv.states])
        for i, st in enumerate(sv.states):
            st.f[0], st.v[0] = f0s[c][i], v0s[c][i]
    try:
        def items(name, l):
            return [([getattr(st, name)[l] for st in sv.states], st0.own_rows(l), st0.ncols(l)) for sv in svs]
        comm.exchange_many(items("f", 0))
        for l in range(nl):
            last = (l + 1 == nl)
            for c, sv in enumerate(svs):
                with _On(c):
                    for st in sv.states:
                        sv._leg(st, l, MODE_DOWN_ZERO, None, st.f[l], st.tmp[l], rc=(st.fg if last else st.f[l + 1]))
            join()
            if last:
                for sv in svs:
                    comm.allgather_rows([st.fg for st in sv.states], st0.own_rows(nl), st0.ncols(nl))
            else:
                comm.exchange_many(items("f", l + 1))
        for c, sv in enumerate(svs):
            with _On(c):
                stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                for st in sv.states:
                    _lib.check(lib.mgcmt_vcycle_from(st.coarse, nl, float(sv.shift), _lib.SMOOTH_WJACOBI, float(sv.omega),
                                                     C.c_void_p(st.vg.data_ptr()), C.c_void_p(st.fg.data_ptr()), stream))
        join()
        for l in range(nl - 1, -1, -1):
            last = (l + 1 == nl)
            it = items("tmp", l)
            if not last:
                it += items("v", l + 1)
            comm.exchange_many(it)
            for c, sv in enumerate(svs):
                with _On(c):
                    for st in sv.states:
                        sv._leg(st, l, MODE_UP, st.tmp[l], st.f[l], st.v[l], e=(st.vg if last else st.v[l + 1]))
            join()
        if lam is not None:
            comm.exchange_many(items("v", 0))
            for c, sv in enumerate(svs):
                with _On(c):
                    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                    for i, st in enumerate(sv.states):
                        _lib.check(lib.mgcmt_slab_rayleigh(st.slab, 0, C.c_void_p(st.v[0].data_ptr()),
                                                           C.c_void_p(lam[i][c].data_ptr()), stream))
            join()
            comm.allreduce_sum(lam)
    finally:
        for sv, sav in zip(svs, saved):
            for st, (f, v) in zip(sv.states, sav):
                st.f[0], st.v[0] = f, v


# ----------------------------------------------------------------------------------------------------
# the same step issued natively (csrc/slab_block.cu): NCCL called from C++, no Python between the launches
# ----------------------------------------------------------------------------------------------------
def _nccl_library_path():
    """the NCCL shared object this process already mapped (PyTorch's), else the bundled wheel, else the soname"""
    import os
    try:
        with open("/proc/self/maps") as fh:
            for line in fh:
                if "libnccl" in line and ".so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    try:
        import nvidia.nccl
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return "libnccl.so.2"


class NativeSlabBlock:
    """k slab-decomposed V(4,4) cycles per call, driven from C++ (mgcmt_slabblock_*): the production form of
    `vcycle_block` + `SlabVCycle.gramschmidt_gram`.  One rank per process; with world > 1 torch.distributed must be
    initialised (it carries the 128-byte NCCL id from rank 0 to the others, nothing else).  Bit-for-bit the same
    results as the Python-driven path (same kernels, same order): tools/check_native_slab.py."""

    def __init__(self, op: SeparableOperator, world, rank, k, lowest_level=8, gather_cols=2048, omega=None,
                 stagger=False, smoother="wjacobi"):
        torch = _lib.require_cuda()
        lib = _lib.load()
        if op.nrows != op.ncols:
            raise ValueError("slab decomposition is for square 2-D grids")
        self.op, self.world, self.rank, self.k = op, int(world), int(rank), int(k)
        self.N = op.ncols
        if smoother not in ("wjacobi", "rbgs"):
            raise ValueError("slab smoother must be wjacobi or rbgs")
        self.smoother = smoother
        if omega is None:
            omega = 1.0 if smoother == "rbgs" else 2. / 3.
        self.nlev = plan_levels(self.N, world, gather_cols)
        if self.nlev < 1:
            raise ValueError("grid too small to decompose over %d ranks (use the single-GPU path)" % world)
        self.own0 = self.N // world
        self.begin0 = rank * self.own0
        self.slab_size = (self.own0 + 2 * HALO) * self.N
        self.comm, self.comm2 = C.c_void_p(), C.c_void_p()
        self.handle = C.c_void_p()
        if world > 1:
            import torch.distributed as dist
            _lib.check(lib.mgcmt_nccl_load(_nccl_library_path().encode()))
            # two communicators: the block runs as two halves half a phase apart (stagger=False: one, lock-step)
            for comm in ([self.comm, self.comm2] if stagger and self.k >= 2 else [self.comm]):
                ident = torch.zeros(128, dtype=torch.uint8)
                if rank == 0:
                    _lib.check(lib.mgcmt_nccl_unique_id(C.c_void_p(ident.data_ptr())))
                dev = ident.cuda() if dist.get_backend() == "nccl" else ident
                dist.broadcast(dev, src=0)
                ident = dev.cpu()
                _lib.check(lib.mgcmt_nccl_comm_create(C.c_void_p(ident.data_ptr()), self.world, self.rank, C.byref(comm)))
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.mgcmt_slabblock_create(self.comm, self.comm2, self.world, self.rank, self.N, self.nlev, int(lowest_level),
                                              self.k, hp(op.row[0]), hp(op.row[1]), hp(op.row[2]), hp(op.col[0]), hp(op.col[1]),
                                              hp(op.col[2]), float(omega), stream, C.byref(self.handle)))
        _lib.check(lib.mgcmt_slabblock_set_smoother(self.handle, _lib.SMOOTH_RBGS if smoother == "rbgs" else _lib.SMOOTH_WJACOBI,
                                                    float(omega)))
        self._ptrs = (C.c_void_p * self.k)
        self._shifts = (C.c_double * self.k)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        lib = _lib.load()
        if getattr(self, "handle", None):
            lib.mgcmt_slabblock_destroy(self.handle)
            self.handle = None
        for name in ("comm", "comm2"):
            if getattr(self, name, None):
                lib.mgcmt_nccl_comm_destroy(getattr(self, name))
                setattr(self, name, None)

    def new_block(self):
        """(k, slab_size) zero tensor: k finest-level slab vectors (owned rows + halos) of this rank"""
        torch = _lib.require_cuda()
        return torch.zeros(self.k, self.slab_size, dtype=torch.float64, device="cuda")

    def owned(self, x):
        """view of the owned rows of one finest-level slab vector"""
        return x.view(self.own0 + 2 * HALO, self.N)[HALO:HALO + self.own0]

    def cycle(self, shifts, F, W, lam=None):
        """W[c] <- one V(4,4) cycle on (H - shifts[c]) w = F[c] from a zero start, for all k vectors; lam (k, 2)
        receives the Rayleigh sums w^T H w, w^T w (all ranks).  Asynchronous on the current stream."""
        torch = _lib.require_cuda()
        es = F.shape[1] * 8
        f0 = self._ptrs(*[F.data_ptr() + c * es for c in range(self.k)])
        v0 = self._ptrs(*[W.data_ptr() + c * es for c in range(self.k)])
        _lib.check(_lib.load().mgcmt_slabblock_cycle(self.handle, self._shifts(*[float(s) for s in shifts]), f0, v0,
                                                     C.c_void_p(lam.data_ptr()) if lam is not None else None,
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def profile(self, on=True):
        _lib.check(_lib.load().mgcmt_slabblock_profile(self.handle, 1 if on else 0))

    def profile_read(self, with_lam=True):
        """per-stage milliseconds of the last profiled cycle: list of (stage name, comm ms, compute ms)"""
        n = 4 * self.nlev + 2
        comm, comp, ns = (C.c_double * n)(), (C.c_double * n)(), C.c_int()
        _lib.check(_lib.load().mgcmt_slabblock_profile_read(self.handle, 1 if with_lam else 0, comm, comp, C.byref(ns)))
        two = lambda l: self.smoother == "rbgs" and l > 0      # 9-point red-black legs are two passes
        names = []
        for l in range(self.nlev):
            names += ["down L%d pass 1" % l, "down L%d pass 2" % l] if two(l) else ["down L%d" % l]
        names.append("coarse (replicated)")
        for l in range(self.nlev - 1, -1, -1):
            names += ["up L%d pass 1" % l, "up L%d pass 2" % l] if two(l) else ["up L%d" % l]
        names.append("rayleigh (separate pass)")
        return [(names[i], comm[i], comp[i]) for i in range(ns.value)]

    def gram(self, W):
        torch = _lib.require_cuda()
        _lib.check(_lib.load().mgcmt_slabblock_gram(self.handle, C.c_void_p(W.data_ptr()), W.shape[1],
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
