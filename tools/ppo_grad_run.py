"""Launch msched_ppo_grad a few times at the cfg3 acceptor shape (for ncu captures).
Usage: python tools/ppo_grad_run.py [T] [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_scheduling_b200 import policy  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device("cuda", 0)
U, n_in, A = 6, 15, 7
g = torch.Generator(device=dev).manual_seed(1)
A_w = policy.MlpGroup.random(n_in, 16, A, U, dev, seed=1)
C_w = policy.MlpGroup.random(n_in, 16, 1, U, dev, seed=2)
X = torch.randint(-1, 9, (T * B, U, n_in), generator=g, dtype=torch.int16, device=dev)
act = torch.randint(0, A, (T * B, U), generator=g, dtype=torch.int32, device=dev)
lp = torch.full((T * B, U), -1.9459, device=dev)
G = torch.randn((T * B, U), generator=g, device=dev)
ids = torch.arange(U, dtype=torch.int32, device=dev)
ga, gc = torch.zeros_like(A_w.weights), torch.zeros_like(C_w.weights)
ws = None
for _ in range(3):
    _, ws = policy.ppo_grad(A_w.weights, C_w.weights, n_in, A, X, act, lp, G, ids, ids.view(U, 1), ga, gc, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
policy.ppo_grad(A_w.weights, C_w.weights, n_in, A, X, act, lp, G, ids, ids.view(U, 1), ga, gc, workspace=ws)
e1.record()
torch.cuda.synchronize()
print(f"ppo_grad: {T * B * U} samples, {e0.elapsed_time(e1):.3f} ms")
