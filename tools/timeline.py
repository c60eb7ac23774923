#!/usr/bin/env python
"""Per-CTA timeline of the fused step kernel (msched_debug_timeline): where a tile's lifetime goes.
Run on the GPU box: python tools/timeline.py [cfg3|cfg4] [obs 0/1]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ctypes as C
import bench
from marl_scheduling_b200 import _lib as L
from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
obs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = bench.CONFIGS[name]
dom, mode = cfg["dom"], cfg["mode"]
B = 65536
env = BatchedSchedulingEnv(B, world_params_from_dom(dom, mode.startswith("free")), reward=mode, auction="random",
                           spawn="philox", seed=0, net_zero_offer_reward=dom.get("netZero", 0.5))
ring, gen = bench.make_actions(torch, env, 4, 1)
res = torch.zeros_like(env.result)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
def step(i):
    bench.refresh_actions(env, ring[i % 4], gen)
    if obs: env.step_observe_records(ring[i % 4], res)
    else: env.step_records(ring[i % 4], res)
for i in range(300): step(i)
nc = env.layout.padded_envs // 32
tl = torch.zeros((nc, 8), dtype=torch.int64, device=env.device)
L.check(env.lib.msched_debug_timeline(env.handle, tl.data_ptr()))
flush.fill_(1)
torch.cuda.synchronize()
step(0)
torch.cuda.synchronize()
t = tl.cpu().numpy()
clk = 1.965  # GHz
g0 = t[:, 1].min()
print("CTAs", nc, "kernel span by globaltimer: %.2f us" % ((t[:, 7].max() - g0) / 1e3))
st = (t[:, 1] - g0) / 1e3
en = (t[:, 7] - g0) / 1e3
print("CTA start  (us) pct 0/25/50/75/100:", np.percentile(st, [0, 25, 50, 75, 100]).round(2))
print("CTA end    (us) pct 0/25/50/75/100:", np.percentile(en, [0, 25, 50, 75, 100]).round(2))
d = lambda a, b: (t[:, b] - t[:, a]) / clk / 1e3
for nm, a, b in (("pre-work (P0)", 2, 3), ("wait for tile", 3, 4), ("P1..P5 compute", 4, 5), ("bulk store read", 5, 6), ("whole CTA", 2, 6)):
    x = d(a, b)
    print(f"{nm:18s} us mean {x.mean():6.2f}  pct 5/50/95: {np.percentile(x, [5, 50, 95]).round(2)}")
sm = t[:, 0]
per = np.bincount(sm.astype(int))
print("CTAs per SM min/mean/max", per[per > 0].min(), per[per > 0].mean().round(2), per.max(), "SMs used", (per > 0).sum())
# concurrency on one SM
s0 = int(sm[0])
rows = t[sm == s0]
rows = rows[np.argsort(rows[:, 2])]
base = rows[0, 2]
for r in rows:
    print("  SM%d cta: start %7.2f  tile %7.2f  computed %7.2f  done %7.2f us" % (s0, (r[2]-base)/clk/1e3, (r[4]-base)/clk/1e3, (r[5]-base)/clk/1e3, (r[6]-base)/clk/1e3))
