#!/usr/bin/env python
"""Times the actor-forward kernel (tensor-core vs fp32 SIMT) and the returns kernel on the cfg3
shapes at 65,536 envs.  Run on the GPU box: python tools/policy_bench.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from marl_scheduling_b200 import policy

dev = torch.device("cuda", 0)
B = 65536
shapes = [("acceptor cfg3", 15, 16, 7, 6), ("core chooser cfg3", 8, 16, 4, 6), ("price chooser cfg3", 4, 16, 9, 6),
          ("acceptor cfg2", 27, 16, 13, 16), ("offer cfg2", 10, 16, 5, 12), ("agg offer 12->32->64", 12, 32, 64, 2),
          ("agg acceptor cfg3 343", 45, 32, 343, 2), ("fully agg cfg3 21952", 57, 64, 21952, 2),
          ("agg acceptor cfg2 28561", 108, 32, 28561, 4)]
only = sys.argv[1] if len(sys.argv) > 1 else None   # e.g. "tc:acceptor cfg3"
for impl in ("tc", "simt"):
    os.environ["MSCHED_ACTOR_IMPL"] = impl
    for name, nin, h, A, units in shapes:
        if only and only != f"{impl}:{name}":
            continue
        if A > 64 and impl == "simt":
            continue
        grp = policy.MlpGroup.random(nin, h, A, units, dev, seed=1)
        x = torch.randint(-2, 11, (B, units * nin), dtype=torch.int16, device=dev)
        act = torch.empty(B * units, dtype=torch.int32, device=dev)
        lp = torch.empty(B * units, dtype=torch.float32, device=dev)
        reps = 50 if A <= 64 else (10 if A < 1000 else 2)
        for _ in range(2 if A > 64 else 5):
            policy.actor_forward(grp, x, nin, units, B, seed=1, step=0, action=act, logprob=lp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            policy.actor_forward(grp, x, nin, units, B, seed=1, step=i, action=act, logprob=lp)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        mac = nin * h + h * h + h * A
        print(f"{impl:5s} {name:24s} rows {B*units:8d}  {us:8.1f} us  {B*units/us:8.1f} rows/us  {2*mac*B*units/us/1e6:7.2f} TFLOP/s")
if only:
    sys.exit(0)
r = torch.randn(200, B * 6, device=dev)
for _ in range(3):
    policy.returns(r, 0.87, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    policy.returns(r, 0.87, True)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 10
print(f"returns T=200 M={B*6}: {us:.1f} us, {12*r.numel()/us/1e3:.1f} GB/s (12 B/element)")
