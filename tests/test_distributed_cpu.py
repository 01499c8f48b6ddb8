"""world_size-2 (and 3) gloo tests of the multi-GPU host plumbing on CPU: shard ranges, padding
for uneven T/W, block packing and both gather helpers.  The per-rank compute is replaced by a
deterministic stand-in (the real kernels need a GPU and are covered by the -m gpu tests)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_blocks(t_idx, n, m):
    """Deterministic stand-in for (At, Bt, ct) at global timestep indices t_idx."""
    t = t_idx.to(torch.float64)
    At = t[:, None, None] + torch.arange(n * n, dtype=torch.float64).reshape(1, n, n) * 1e-3
    Bt = -t[:, None, None] + torch.arange(n * m, dtype=torch.float64).reshape(1, n, m) * 1e-2
    ct = 0.5 * t[:, None] + torch.arange(n, dtype=torch.float64)[None, :]
    return At, Bt, ct


def _worker(rank, world, port, T, n, m):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from irs_mpc_b200 import distributed as D
        start, stop, per = D.shard_range(T, world, rank)
        At, Bt, ct = _fake_blocks(torch.arange(start, stop), n, m)
        local = D.pack_abc(At, Bt, ct)
        assert local.shape == (stop - start, n * (n + m + 1))
        full = D.gather_rows(local, T)
        A2, B2, c2 = D.unpack_abc(full, n, m)
        Ae, Be, ce = _fake_blocks(torch.arange(T), n, m)
        assert torch.equal(A2, Ae) and torch.equal(B2, Be) and torch.equal(c2, ce)
        # sample-sharded exchange: per-rank blocks gathered in rank order
        mine = torch.full((T, 5), float(rank + 1), dtype=torch.float64)
        allr = D.gather_ranks(mine)
        assert allr.shape == (world, T, 5)
        for r in range(world):
            assert torch.all(allr[r] == r + 1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 100), (2, 7), (3, 100), (3, 2)])
def test_gather_plumbing_gloo(world, T):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, T, 5, 2), nprocs=world, join=True)


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from irs_mpc_b200.distributed import shard_range
    for total in (1, 2, 7, 100, 101):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                a, b, per = shard_range(total, world, r)
                assert 0 <= a <= b <= total and b - a <= per
                covered += list(range(a, b))
            assert covered == list(range(total))
