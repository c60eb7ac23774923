#!/usr/bin/env python
"""Size check beyond BASELINE's batch: B environments of the config-3 domain (default 2^21) stepped with random
actions; the last 128 environments must walk through exactly the states of a 128-environment batch created with
env_offset = B - 128 and fed the same actions (every device draw is keyed on the GLOBAL environment index), and the
job-conservation identity must hold over the whole batch.  Run on the GPU box: python tools/big_batch_check.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom



def check(B=1 << 21, steps=40):
    dom = dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1])
    mk = lambda n, off: BatchedSchedulingEnv(n, world_params_from_dom(dom, True), reward="free_comm", auction="random",
                                             spawn="philox", seed=5, env_offset=off)
    big, win = mk(B, 0), mk(128, B - 128)
    lay, dev = big.layout, big.device
    g = torch.Generator(device=dev).manual_seed(3)
    term = torch.zeros(B, dtype=torch.int64, device=dev)
    for t in range(steps):
        big.acceptor_actions.random_(0, 7, generator=g)
        big.offer_core_actions.random_(0, 4, generator=g)
        big.offer_price_actions.random_(0, 9, generator=g)
        win.action[:128].copy_(big.action[B - 128:B])
        big.step_observe_records()
        win.step_observe_records()
        term += (big.result[:B, lay.r_counts] >> 16) & 0xFF
        torch.cuda.synchronize()
        assert torch.equal(big.state[B - 128:B], win.state[:128]), t
        assert torch.equal(big.result[B - 128:B], win.result[:128]), t
        assert torch.equal(big._obs_buffer()[B - 128:B], win._obs_buffer()[:128]), t
    assert int(big.result[:B, lay.r_flags].max()) == 0
    C, NL = 3, 6
    st = big.state[:B]
    s_slot = 2 + 3 * C + (C + 3) // 4
    present = (st[:, 3:2 + 3 * C:3] > 0).sum(1) + (st[:, s_slot + 1::4][:, :NL] > 0).sum(1)
    spawned = st[:, 0].long() - 1
    assert torch.equal(spawned, term + present), "job conservation"
    return int(term.sum())


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    print(f"ok: {B} environments x 40 steps, last window == offset batch, {check(B)} jobs terminated")
