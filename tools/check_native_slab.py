"""Multi-GPU check of the natively driven slab block against the Python-driven slab path (same kernels, same order:
the owned rows and the Rayleigh sums must be identical), plus the host issue time of both.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_native_slab.py [N]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from multigridcmt_b200 import MGCMTStencilMaker
    from multigridcmt_b200.slab import NativeSlabBlock, SlabVCycle, TorchDistComm, vcycle_block
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    gather = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    k = 4
    H = (-1.0 / np.pi ** 2) * MGCMTStencilMaker().laplacian(N, "2d", matrix_free=True)
    shifts = [1.7665, 4.3863, 4.3864, 7.0062]
    nb = NativeSlabBlock(H, world, rank, k, lowest_level=8, gather_cols=gather, stagger=True)
    nb1 = NativeSlabBlock(H, world, rank, k, lowest_level=8, gather_cols=gather, stagger=False)
    comm = TorchDistComm()
    svs = [SlabVCycle(H, world, comm, [rank], lowest_level=8, gather_cols=gather) for _ in range(k)]
    streams = [torch.cuda.Stream() for _ in range(k)]
    st = svs[0].states[0]
    g = torch.Generator(device="cuda"); g.manual_seed(7 + rank)
    own = torch.rand(k, nb.own0, N, dtype=torch.float64, device="cuda", generator=g) - 0.5
    F1, W1, F2, W2 = nb.new_block(), nb.new_block(), nb.new_block(), nb.new_block()
    for c in range(k):
        nb.owned(F1[c]).copy_(own[c]); nb.owned(F2[c]).copy_(own[c])
    lam1 = torch.zeros(k, 2, dtype=torch.float64, device="cuda"); lam2 = torch.zeros_like(lam1)

    def native():
        nb.cycle(shifts, F1, W1, lam1)
        nb.gram(W1)

    F3, W3 = nb.new_block(), nb.new_block()
    for c in range(k):
        nb.owned(F3[c]).copy_(own[c])
    lam3 = torch.zeros_like(lam1)

    def native_lockstep():
        nb1.cycle(shifts, F3, W3, lam3)
        nb1.gram(W3)

    def python():
        vcycle_block(svs, shifts, [[F2[c]] for c in range(k)], [[W2[c]] for c in range(k)], lam=[lam2], streams=streams)
        svs[0].gramschmidt_gram([W2])

    native(); python(); native_lockstep()
    torch.cuda.synchronize(); dist.barrier()
    same_w = all(torch.equal(nb.owned(W1[c]), nb.owned(W2[c])) and torch.equal(nb.owned(W1[c]), nb.owned(W3[c])) for c in range(k))
    close = lambda a, b: bool(((a - b).abs() <= 1e-12 * b.abs()).all())   # fused Rayleigh stage vs separate pass
    same_lam = close(lam1, lam2) and close(lam3, lam2)
    res = {"rank": rank, "world": world, "N": N, "slab_levels": nb.nlev, "identical_vectors": same_w, "identical_rayleigh": same_lam,
           "lam": (lam1[:, 0] / lam1[:, 1]).cpu().tolist()}
    for name, fn in (("native", native), ("native_lockstep", native_lockstep), ("python", python)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        issue = (time.perf_counter() - t0) / reps * 1e3
        torch.cuda.synchronize(); dist.barrier()
        res[name] = {"ms_per_step": e0.elapsed_time(e1) / reps, "host_issue_ms_per_step": issue}
    # per-stage breakdown of one lock-step cycle (CUDA events on the block's ordering stream)
    nb1.profile(True)
    native_lockstep()
    torch.cuda.synchronize()
    res["stages_ms(name, comm, compute)"] = [(n, round(a, 4), round(c, 4)) for n, a, c in nb1.profile_read(True)]
    nb1.profile(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        nb1.gram(W3)
    e1.record(); torch.cuda.synchronize()
    res["gram_ms"] = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(json.dumps(res), flush=True)
    ok = same_w and same_lam
    for sv in svs:
        sv.close()
    nb.close(); nb1.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
