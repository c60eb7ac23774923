"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (env sharding, the
flat-bucket gradient all-reduce of shared nets, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_scheduling_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, n = D.shard_envs(65536 * world + 3, world, rank)
        a = torch.nn.Parameter(torch.zeros(5, 7))
        b = torch.nn.Parameter(torch.zeros(11))
        a.grad = torch.full((5, 7), float(rank + 1))
        b.grad = torch.arange(11.0) * (rank + 1)
        D.allreduce_gradients([a, b])
        mx = D.max_over_ranks(10.0 + rank)
        # the PPO update kernels keep both heads' gradients in ONE flat buffer: a single all-reduce per epoch
        flat = torch.arange(9.0) * (rank + 1)
        D.allreduce_mean_(flat)
        assert torch.allclose(flat, torch.arange(9.0) * 1.5)
        q.put((rank, off, n, a.grad.clone(), b.grad.clone(), mx))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharding_allreduce_and_timing_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (r0, off0, n0, a0, b0, m0), (r1, off1, n1, a1, b1, m1) = res
    total = 65536 * world + 3
    assert off0 == 0 and off1 == n0 and n0 + n1 == total and abs(n0 - n1) <= 1
    assert torch.equal(a0, torch.full((5, 7), 1.5)) and torch.equal(a1, a0)
    assert torch.allclose(b0, torch.arange(11.0) * 1.5) and torch.equal(b0, b1)
    assert m0 == m1 == 11.0


def test_shards_tile_the_range():
    for total in (1, 7, 65536, 262144 + 5):
        for world in (1, 2, 3, 8):
            nxt = 0
            for r in range(world):
                off, n = D.shard_envs(total, world, r)
                assert off == nxt
                nxt += n
            assert nxt == total


def test_number_to_n_dimensional_action_matches_reference_rule():
    """src/Agent.py:644-666: index 0 is the least significant digit; out of range raises."""
    import pytest
    from marl_scheduling_b200.SchedulingEnvironment import numberToNDimensionalAction
    assert numberToNDimensionalAction(torch.tensor(0), 13, 4).tolist() == [0, 0, 0, 0]
    assert numberToNDimensionalAction(torch.tensor(28560), 13, 4).tolist() == [12, 12, 12, 12]
    assert numberToNDimensionalAction(torch.tensor(13 * 5 + 2), 13, 4).tolist() == [2, 5, 0, 0]
    assert numberToNDimensionalAction(torch.tensor([7, 124]), 5, 3).tolist() == [[2, 1, 0], [4, 4, 4]]
    with pytest.raises(ValueError):
        numberToNDimensionalAction(torch.tensor(125), 5, 3)
    with pytest.raises(ValueError):
        numberToNDimensionalAction(torch.tensor(-1), 5, 3)
