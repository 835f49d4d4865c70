"""world_size-2 gloo tests (CPU) of the multi-GPU host logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tvm_b200 import dist as tdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # bench timing rule: max over ranks
        ms = tdist.max_over_ranks(10.0 + 5.0 * rank)
        assert ms == 15.0
        # coset sharding: every evaluation-domain row has exactly one owner and a unique local slot
        r, n = 8, 16
        mine = tdist.cosets_of_rank(r, world, rank)
        assert mine == list(range(rank, r, world))
        local = torch.full((len(mine) * n, 3), -1, dtype=torch.int64)
        for i in range(r * n):
            owner, m = tdist.local_row(i, r, world, n)
            assert owner == tdist.row_owner(i, r, world)
            if owner == rank:
                assert local[m, 0] == -1
                local[m] = torch.tensor([i, i + 1, i + 2])
        assert (local[:, 0] >= 0).all()
        # gather and restore natural order
        g = tdist.all_gather_rows(local, world)
        for i in range(r * n):
            owner, m = tdist.local_row(i, r, world, n)
            assert g[owner, m, 0].item() == i
        out.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_coset_sharding_and_timing_reduction_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]


def test_more_ranks_than_cosets_is_rejected():
    with pytest.raises(ValueError):
        tdist.cosets_of_rank(8, 16, 0)
