"""Times msched_returns (TMA-tiled kernel) on a few buffer shapes; MSCHED_RETURNS_W=32|64|128 forces a tile width.
Run on the GPU box: python tools/returns_bench.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from marl_scheduling_b200 import policy
dev = torch.device("cuda", 0)
for T, M in ((200, 393216), (16, 393216), (64, 1 << 20)):
    rew = torch.randn(T, M, device=dev)
    for _ in range(3): policy.returns(rew, 0.8733, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): policy.returns(rew, 0.8733, True)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(os.environ.get("MSCHED_RETURNS_W", "auto"), T, M, f"{us:.1f} us", f"{8 * T * M / us / 1e3:.0f} GB/s (8 B/elem: read r, write G)")
