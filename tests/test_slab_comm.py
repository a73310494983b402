"""CPU: the halo-exchange / all-gather plumbing of the multi-GPU path (multigridcmt_b200/slab.py) with
world_size-2 and -4 gloo process groups, against the single-process LocalComm emulation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multigridcmt_b200.slab import HALO, LocalComm, TorchDistComm, halo_views, plan_levels


def test_plan_levels():
    assert plan_levels(16384, 8) == 3            # 16384, 8192, 4096 distributed; 2048^2 and below replicated
    assert plan_levels(16384, 1) == 3
    assert plan_levels(4096, 2, gather_cols=512) == 3
    assert plan_levels(512, 2, gather_cols=128) == 2
    assert plan_levels(1024, 8, gather_cols=512) == 1
    assert plan_levels(1024, 16, gather_cols=64) == 1   # 64 rows per rank: only the finest level is distributed
    assert plan_levels(1024, 32, gather_cols=64) == 0   # 32 rows per rank: nothing to distribute
    with pytest.raises(ValueError):
        plan_levels(1000, 3)


def _slab(rank, own, ncols):
    x = torch.zeros((own + 2 * HALO) * ncols, dtype=torch.float64)
    a = x.view(own + 2 * HALO, ncols)
    a[HALO:HALO + own] = (rank * own + torch.arange(own, dtype=torch.float64))[:, None] * 1000 + torch.arange(ncols, dtype=torch.float64)
    return x


def _expected(rank, world, own, ncols):
    x = _slab(rank, own, ncols).view(own + 2 * HALO, ncols)
    glob = torch.cat([_slab(r, own, ncols).view(own + 2 * HALO, ncols)[HALO:HALO + own] for r in range(world)])
    lo = rank * own - HALO
    for i in range(own + 2 * HALO):
        g = lo + i
        if 0 <= g < world * own:
            x[i] = glob[g]
    return x.reshape(-1)


def test_local_comm_exchange_and_allgather():
    world, own, ncols = 4, 16, 8
    arrs = [_slab(r, own, ncols) for r in range(world)]
    LocalComm(world).exchange(arrs, own, ncols)
    for r in range(world):
        assert torch.equal(arrs[r], _expected(r, world, own, ncols))
    fulls = [torch.zeros(world * own * ncols, dtype=torch.float64) for _ in range(world)]
    for r in range(world):
        fulls[r].view(-1, ncols)[r * own:(r + 1) * own] = r + 1
    LocalComm(world).allgather_rows(fulls, own, ncols)
    for r in range(world):
        assert torch.equal(fulls[r], fulls[0]) and float(fulls[r].view(-1, ncols)[:, 0].sum()) == own * sum(range(1, world + 1))
    sc = [torch.tensor([float(r), 1.0], dtype=torch.float64) for r in range(world)]
    LocalComm(world).allreduce_sum(sc)
    assert all(torch.equal(x, torch.tensor([6.0, 4.0], dtype=torch.float64)) for x in sc)


def _worker(rank, world, port, own, ncols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = TorchDistComm()
        x = _slab(rank, own, ncols)
        comm.exchange([x], own, ncols)
        ok = torch.equal(x, _expected(rank, world, own, ncols))
        full = torch.zeros(world * own * ncols, dtype=torch.float64)
        full.view(-1, ncols)[rank * own:(rank + 1) * own] = rank + 1
        comm.allgather_rows([full], own, ncols)
        want = torch.cat([torch.full((own, ncols), float(r + 1), dtype=torch.float64) for r in range(world)]).reshape(-1)
        ok = ok and torch.equal(full, want)
        s = torch.tensor([float(rank), 1.0], dtype=torch.float64)
        comm.allreduce_sum([s])
        ok = ok and float(s[0]) == sum(range(world)) and float(s[1]) == world
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_torch_dist_comm_gloo(world):
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 16, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert all(res[r] for r in range(world)), res
