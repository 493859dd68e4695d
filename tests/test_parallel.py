"""Host-side logic of member sharding, exercised with world_size-2 gloo process groups on the CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marigold_b200 import parallel


def test_member_indices_cover_and_partition():
    for E in (1, 2, 3, 8, 10, 16):
        for G in (1, 2, 4, 8):
            got = sorted(m for r in range(G) for m in parallel.member_indices(E, r, G))
            assert got == list(range(E))
            assert max(len(parallel.member_indices(E, r, G)) for r in range(G)) == parallel.slots_per_rank(E, G)
    # config 4 of BASELINE.json: E=10 on 8 GPUs -> 2,2,1,1,1,1,1,1 (makespan 2 => 5x ceiling)
    assert [len(parallel.member_indices(10, r, 8)) for r in range(8)] == [2, 2, 1, 1, 1, 1, 1, 1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, E, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        mine = parallel.member_indices(E, rank, world_size)
        # member m's "prediction" is a tensor filled with m + 1
        local = torch.stack([torch.full((1, 4, 6), float(m + 1)) for m in mine]) if mine else torch.empty(0, 1, 4, 6)
        out = parallel.gather_members(local, E)
        ok = out.shape == (E, 1, 4, 6) and all(bool((out[m] == m + 1).all()) for m in range(E))
        ms = parallel.barrier_max_ms(10.0 * (rank + 1))
        q.put((rank, ok, ms))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("E", [2, 3, 5])
def test_gather_members_gloo_world2(E):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, E, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(abs(ms - 20.0) < 1e-9 for _, _, ms in res)      # max over ranks


def test_single_process_is_identity():
    x = torch.randn(3, 1, 4, 4)
    assert parallel.world() == (0, 1)
    assert torch.equal(parallel.gather_members(x, 3), x)
