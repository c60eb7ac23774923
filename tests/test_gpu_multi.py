"""GPU, two ranks over NCCL (skipped on a one-GPU box): the only exchange of the data-parallel path -- the
all-reduce of the PPO gradient of nets shared across env shards.  Each rank runs msched_ppo_grad on
its shard of the buffer; the averaged gradient must equal the oracle's gradient over the whole buffer, and after
msched_adam_step both ranks must hold bit-identical weights."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n_in, A, U, TB = 15, 7, 3, 640                       # globally shared net: every unit of every env shard
        g = torch.Generator().manual_seed(11)
        x = torch.randint(-1, 9, (TB, U, n_in), generator=g, dtype=torch.int16)
        act = torch.randint(0, A, (TB, U), generator=g, dtype=torch.int32)
        lp = (torch.randn(TB, U, generator=g) * 0.3 - float(np.log(A)))
        G = torch.randn(TB, U, generator=g)
        lo, hi = D.shard_envs(TB, world, rank)
        hi += lo
        sl = slice(lo, hi)
        aw = policy.MlpGroup.random(n_in, 16, A, 1, dev, seed=3).weights
        cw = policy.MlpGroup.random(n_in, 16, 1, 1, dev, seed=4).weights
        flat = torch.zeros(aw.numel() + cw.numel(), device=dev)
        ga, gc = flat[: aw.numel()].view_as(aw), flat[aw.numel():].view_as(cw)
        net_ids = torch.zeros(1, dtype=torch.int32, device=dev)
        unit_ids = torch.arange(U, dtype=torch.int32, device=dev).view(1, U)
        policy.ppo_grad(aw, cw, n_in, A, x[sl].to(dev), act[sl].contiguous().to(dev), lp[sl].contiguous().to(dev),
                        G[sl].contiguous().to(dev), net_ids, unit_ids, ga, gc)
        D.allreduce_mean_(flat)                               # NCCL, one flat bucket
        ma, va, mc, vc = (torch.zeros_like(t) for t in (aw, aw, cw, cw))
        a1, c1 = aw.clone(), cw.clone()
        policy.adam_step(a1.view(-1), ga.view(-1), ma.view(-1), va.view(-1), 3e-4, 1)
        policy.adam_step(c1.view(-1), gc.view(-1), mc.view(-1), vc.view(-1), 1e-3, 1)
        torch.cuda.synchronize()
        out = dict(ga=ga.cpu().numpy(), gc=gc.cpu().numpy(), a1=a1.cpu().numpy(), c1=c1.cpu().numpy())
        if rank == 0:
            out.update(aw=aw.cpu().numpy(), cw=cw.cpu().numpy(), x=x.numpy(), act=act.numpy(), lp=lp.numpy(), G=G.numpy())
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_nccl_allreduced_ppo_gradient_equals_whole_batch_gradient():
    import torch
    import torch.multiprocessing as mp
    from oracle import oracle as O
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=120) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    r0, r1 = res[0], res[1]
    for k in ("ga", "gc", "a1", "c1"):
        assert np.array_equal(r0[k], r1[k]), k               # both ranks hold the same averaged gradient and weights
    n_in = r0["x"].shape[-1]
    oa, oc, _ = O.ppo_loss_grads(r0["aw"][0], r0["cw"][0], r0["x"].reshape(-1, n_in), r0["act"].reshape(-1).astype(np.int64),
                                 r0["lp"].reshape(-1).astype(np.float64), r0["G"].reshape(-1), 0.2)
    np.testing.assert_allclose(r0["ga"][0], oa, rtol=1e-4, atol=1e-5 * np.abs(oa).max())
    np.testing.assert_allclose(r0["gc"][0], oc, rtol=1e-4, atol=1e-5 * np.abs(oc).max())
    assert np.abs(r0["a1"] - r0["aw"]).max() > 1e-4           # the step moved the weights
