"""Multi-GPU parity: the natively driven row-slab block (real NCCL halo exchange, one rank per GPU) against the
undecomposed single-GPU V-cycle on identical inputs.  Exit code 0 iff every rank's owned rows agree to 1e-12 relative
and the Rayleigh sums to 1e-12.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_slab_vs_single.py [N] [gather_cols] [wjacobi|rbgs]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from multigridcmt_b200 import MGCMTStencilMaker, _lib
    from multigridcmt_b200.hierarchy import _ptr, _stream_ptr, get_hierarchy
    from multigridcmt_b200.slab import NativeSlabBlock
    lib = _lib.load()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    gather = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    smoother = sys.argv[3] if len(sys.argv) > 3 else "wjacobi"
    code, omega = (_lib.SMOOTH_RBGS, 1.0) if smoother == "rbgs" else (_lib.SMOOTH_WJACOBI, 2. / 3.)
    k = 4
    H = (-1.0 / np.pi ** 2) * MGCMTStencilMaker().laplacian(N, "2d", matrix_free=True)
    shifts = [1.7665, 4.3863, 4.3864, 7.0062]
    nb = NativeSlabBlock(H, world, rank, k, lowest_level=8, gather_cols=gather, smoother=smoother, stagger=True)
    g = torch.Generator(device="cuda"); g.manual_seed(11)          # the same full right-hand sides on every rank
    full = torch.rand(k, N, N, dtype=torch.float64, device="cuda", generator=g) - 0.5
    F, W = nb.new_block(), nb.new_block()
    for c in range(k):
        nb.owned(F[c]).copy_(full[c, nb.begin0:nb.begin0 + nb.own0])
    lam = torch.zeros(k, 2, dtype=torch.float64, device="cuda")
    nb.cycle(shifts, F, W, lam)
    torch.cuda.synchronize()
    # the undecomposed cycle (every rank computes it; only its own rows are compared)
    h = get_hierarchy(H, 8)
    worst, worst_lam, bits = 0.0, 0.0, True
    for c in range(k):
        w = torch.zeros(N * N, dtype=torch.float64, device="cuda")
        out = torch.zeros(2, dtype=torch.float64, device="cuda")
        _lib.check(lib.mgcmt_vcycle_rq(h.handle, shifts[c], 4, 4, code, omega, _ptr(w), _ptr(full[c].reshape(-1)), 1,
                                       _ptr(out), _stream_ptr(torch)))
        mine = nb.owned(W[c])
        ref = w.view(N, N)[nb.begin0:nb.begin0 + nb.own0]
        worst = max(worst, float((mine - ref).norm() / ref.norm()))
        bits = bits and bool(torch.equal(mine, ref))
        worst_lam = max(worst_lam, float(((lam[c] - out).abs() / out.abs()).max()))
    res = torch.tensor([worst, worst_lam, 0.0 if bits else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    ok = bool(res[0] <= 1e-12 and res[1] <= 1e-12)
    if rank == 0:
        print(json.dumps({"world": world, "N": N, "smoother": smoother, "slab_levels": nb.nlev, "max_rel_diff_owned_rows": float(res[0]),
                          "max_rel_diff_rayleigh_sums": float(res[1]), "bit_identical": bool(res[2] == 0.0), "ok": ok}), flush=True)
    nb.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
