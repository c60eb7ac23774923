#!/usr/bin/env python
"""Per-CTA phase timing of the tensor-core actor kernel (first tile of every CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from marl_scheduling_b200 import policy
dev = torch.device("cuda", 0)
B, nin, h, A, units = 65536, 15, 16, 7, 6
grp = policy.MlpGroup.random(nin, h, A, units, dev, seed=1)
x = torch.randint(-2, 11, (B, units * nin), dtype=torch.int16, device=dev)
act = torch.empty(B * units, dtype=torch.int32, device=dev); lp = torch.empty(B * units, dtype=torch.float32, device=dev)
for _ in range(3): policy.actor_forward(grp, x, nin, units, B, action=act, logprob=lp)
tl = torch.zeros((4096, 8), dtype=torch.int64, device=dev)
policy.actor_forward(grp, x, nin, units, B, action=act, logprob=lp, timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().numpy(); t = t[t[:, 0] > 0]
clk = 1.965e3
print("CTAs", len(t))
for nm, a_, b_ in (("alloc+weights", 0, 1), ("x load+panels", 1, 2), ("sync+L1 mma+tanh+store", 2, 3), ("sync+L2 mma+tanh+store", 3, 4), ("sync+L3 mma+ld", 4, 5), ("epilogue", 5, 6), ("first tile total", 1, 6)):
    d = (t[:, b_] - t[:, a_]) / clk
    print(f"{nm:26s} mean {d.mean():7.2f} us  pct 5/50/95 {np.percentile(d, [5, 50, 95]).round(2)}")
sm = t[:, 7]; s0 = sm[0]; rows = t[sm == s0]; rows = rows[np.argsort(rows[:, 0])]; base = rows[0, 0]
for r in rows: print("  cta on SM%d:" % s0, " ".join("%7.2f" % ((r[i] - base) / clk) for i in range(7)))
