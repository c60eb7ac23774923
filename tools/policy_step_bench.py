#!/usr/bin/env python
"""Times msched_policy_step (every PPO unit of a rollout step in one launch) on the cfg3 / cfg2 shapes at 65,536
envs, with and without the experience-buffer outputs.  Run on the GPU box: python tools/policy_step_bench.py [cfg3|cfg2] [variant]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from marl_scheduling_b200 import policy
from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
only = sys.argv[2] if len(sys.argv) > 2 else None
doms = {"cfg3": (dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1]), True),
        "cfg2": (dict(N=4, C=4, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7]), False)}
dom, free = doms[which]
B = int(os.environ.get("B", 65536))
N, C, L = dom["N"], dom["C"], dom["L"]
NL, P = N * L, max(dom["prios"])
env = BatchedSchedulingEnv(B, world_params_from_dom(dom, free), reward="free_comm" if free else "fix", auction="random",
                           spawn="philox", seed=0)
dev, lay = env.device, env.layout
g = torch.Generator(device=dev).manual_seed(1)
for t in range(30):
    env.acceptor_actions.random_(0, 2, generator=g)
    env.offer_core_actions.random_(0, C + 1, generator=g)
    if free:
        env.offer_price_actions.random_(0, P + 1, generator=g)
    env.step_observe_records()
Ua, Uo = N * C, NL
ga = policy.MlpGroup.random(3 + 2 * NL, 16, NL + 1, Ua, dev, seed=1)
go = policy.MlpGroup.random(2 * C + 2, 16, C + 1, Uo, dev, seed=2)
gp = policy.MlpGroup.random(4, 16, P + 1, Uo, dev, seed=3) if free else None
buf = dict(xa=torch.zeros((B, Ua, lay.o_acc_row), dtype=torch.int16, device=dev), xo=torch.zeros((B, Uo, lay.o_off_row), dtype=torch.int16, device=dev),
           xp=torch.zeros((B, Uo, 4), dtype=torch.int16, device=dev),
           a=[torch.zeros((B, n), dtype=torch.int32, device=dev) for n in (Ua, Uo, Uo)],
           lp=[torch.zeros((B, n), dtype=torch.float32, device=dev) for n in (Ua, Uo, Uo)])


def run(variant, reps=300):
    x = variant == "full"
    o = variant in ("full", "noxused")
    A_ = policy.policy_step_group(ga, Ua, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor, 1, buf["a"][0] if o else None, buf["lp"][0] if o else None, x_used=buf["xa"] if x else None)
    O_ = policy.policy_step_group(go, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_core, 2, buf["a"][1] if o else None, buf["lp"][1] if o else None, x_used=buf["xo"] if x else None)
    P_ = policy.policy_step_group(gp, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_price, 3, buf["a"][2] if o else None, buf["lp"][2] if o else None, x_used=buf["xp"] if x else None) if free else None
    for i in range(300):   # warm-up long enough for the SM clock to settle
        policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, C, A_, O_, P_, action_rec=env.action, action_rec_stride=lay.action_halfs, step=i, input_bound=16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, C, A_, O_, P_, action_rec=env.action, action_rec_stride=lay.action_halfs, step=i, input_bound=16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    macs = Ua * (16 * (3 + 2 * NL) + 256 + 16 * (NL + 1)) + Uo * (16 * (2 * C + 2) + 256 + 16 * (C + 1)) + (Uo * (64 + 256 + 16 * (P + 1)) if free else 0)
    print(f"{os.environ.get('MSCHED_POLICY_STEP_IMPL', 'default'):7s} {which} {variant:8s} B={B}: {us:8.1f} us  {2 * macs * B / us / 1e6:6.2f} TFLOP/s  rows/us {B * (Ua + Uo * (2 if free else 1)) / us:8.1f}", flush=True)


for v in ("full", "noxused", "actiononly"):
    if only in (None, v):
        run(v, reps=3 if only == 'ncu' else 300)
