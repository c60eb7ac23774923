"""Diagnostic: long-horizon GPU vs oracle comparison with bench-like uniform random actions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
from oracle import oracle as O

dom = dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1])
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 1200
env = BatchedSchedulingEnv(B, world_params_from_dom(dom, True), reward="free_comm", auction="random",
                           spawn="philox", seed=0)
orc = O.Oracle(B, dom, "free_comm", tie_mode=O.TIE_PHILOX, seed=0, chain_cap=64)
rng = np.random.default_rng(1)
mx = 0
for t in range(T):
    offc = rng.integers(0, 4, (B, 2, 3)); offp = rng.integers(0, 9, (B, 2, 3)); acc = rng.integers(0, 7, (B, 2, 3))
    r = env.step(offc, acc, None, offer_price=offp)
    orc.step(offc, acc, None, offp=offp)
    ok = np.array_equal(r["agent"].cpu().numpy(), orc.r_agent) and np.array_equal(r["flags"].cpu().numpy().astype(np.uint32), orc.flags)
    if t % 100 == 99 or not ok:
        e = env.export_state()
        mx = max(mx, int(e["chain_len"].max()))
        print(t, "ok" if ok else "MISMATCH", "max chain", int(e["chain_len"].max()), "flags", int(r["flags"].max()),
              "oracle flags", int(orc.flags.max()), "hist", np.bincount(e["chain_len"].ravel(), minlength=8)[:12].tolist())
    if not ok:
        break
print("done, max chain", mx)
