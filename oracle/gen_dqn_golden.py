"""Generate tests/golden/dqn.npz by running the UNMODIFIED reference DQN code (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/gen_dqn_golden.py

For three Q-net shapes of the BASELINE configs (acceptor 27->16->13 and offer 10->16->5 of cfg2, acceptor
15->16->7 of cfg3) a reference `DQNEntity` (src/DQNmodules.py:34-76) evaluates integer observation rows
(forward, arg-max), runs its epsilon-greedy `selectAction` with the Python RNG scripted (the exploration draw and
the randrange result are recorded), and the reference `optimize_model` (src/DQNmodules.py:97-154: SmoothL1Loss
against the target net, gradients clamped to [-1, 1], torch.optim.Adam with default learning rate like
src/Agent.py:313-320) takes three steps over batches drawn from a reference `ReplayMemory`, np.random.choice
scripted so that the sampled indices are recorded.  Stored: initial parameters in the C-ABI's flat layout
[W1 | b1 | W2 | b2], the rows, Q-values, actions, the transitions, the batch indices and the parameters after every
step.  tests/test_oracle_golden.py replays them through oracle.dqn_optimize on the CPU, tests/test_gpu_dqn.py
through msched_dqn_select / msched_dqn_grad / msched_adam_step on the GPU.
"""
import copy
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

sys.path.insert(0, H.REFERENCE_SRC)
import DQNmodules as RD  # noqa: E402  (torch + numpy only)


def flat(net):
    seq = net.model
    parts = []
    for i in (0, 2):
        parts += [seq[i].weight.detach().reshape(-1), seq[i].bias.detach().reshape(-1)]
    return torch.cat(parts).numpy().astype(np.float32)


def main():
    out = {}
    cases = [("acc_cfg2", 27, 13, 0.9, 32), ("off_cfg2", 10, 5, 0.5, 32), ("acc_cfg3", 15, 7, 0.8733, 64)]
    for tag, n_in, A, gamma, batch in cases:
        torch.manual_seed(len(tag) * 17 + n_in)
        rng = np.random.default_rng(n_in * 5 + A)
        world = types.SimpleNamespace(randomPolicy=False, round=150)
        env = types.SimpleNamespace(RUN_START=0.9, RUN_END=0.05, RUN_DECAY=200.0)
        net = RD.DQNEntity(world, env, n_in, A)
        out[tag + ".w0"] = flat(net)
        # ---- forward / selectAction (src/DQNmodules.py:52-76) ----
        M = 96
        x = rng.integers(-2, 11, (M, n_in)).astype(np.int16)
        with torch.no_grad():
            q = net.forward(torch.as_tensor(x)).numpy()
        samples = rng.random(M)
        randr = rng.integers(0, A, M)
        acts = np.zeros(M, np.int32)
        orig_random, orig_randrange = random.random, random.randrange
        try:
            for i in range(M):
                random.random = lambda i=i: float(samples[i])
                random.randrange = lambda n, i=i: int(randr[i])
                acts[i] = int(net.selectAction(torch.as_tensor(x[i])))
        finally:
            random.random, random.randrange = orig_random, orig_randrange
        eps = env.RUN_END + (env.RUN_START - env.RUN_END) * np.exp(-1.0 * world.round / env.RUN_DECAY)
        out[tag + ".x"], out[tag + ".q"], out[tag + ".sample"], out[tag + ".randrange"] = x, q, samples, randr.astype(np.int32)
        out[tag + ".action"], out[tag + ".eps"] = acts, np.float64(eps)
        # ---- optimize_model (src/DQNmodules.py:97-154) ----
        n_tr = 200
        S = rng.integers(-2, 11, (n_tr, n_in)).astype(np.int16)
        S2 = rng.integers(-2, 11, (n_tr, n_in)).astype(np.int16)
        Aa = rng.integers(0, A, n_tr).astype(np.int32)
        Rr = rng.integers(-6, 12, n_tr).astype(np.int32)
        mem = RD.ReplayMemory(256)
        for i in range(n_tr):
            # what DQNSchedulingEnv.update*MemoriesAndOptimize pushes (src/SchedulingEnvironment.py:378-389): tuples of
            # ints, the action (an int, or a float after a random action), the reward row of src/Reward.py ([r])
            mem.push(tuple(S[i].tolist()), float(Aa[i]) if i % 3 == 0 else int(Aa[i]), tuple(S2[i].tolist()), np.array([int(Rr[i])]))
        target = copy.deepcopy(net)
        opt = torch.optim.Adam(net.parameters())
        idxs, ws = [], []
        orig_choice = np.random.choice
        try:
            for step in range(3):
                idx = rng.integers(0, n_tr, batch)

                def scripted(arr, n, idx=idx):
                    assert n == len(idx) and len(arr) == n_tr
                    return arr[idx]
                np.random.choice = scripted
                RD.optimize_model(None, mem, batch, net, target, gamma, opt)
                idxs.append(idx.astype(np.int32))
                ws.append(flat(net))
        finally:
            np.random.choice = orig_choice
        out[tag + ".S"], out[tag + ".S2"], out[tag + ".A"], out[tag + ".R"] = S, S2, Aa, Rr
        out[tag + ".idx"], out[tag + ".w_after"] = np.stack(idxs), np.stack(ws)
        out[tag + ".gamma"], out[tag + ".n_in"], out[tag + ".n_actions"] = np.float64(gamma), np.int32(n_in), np.int32(A)
        print(tag, "eps", round(float(eps), 4), "greedy share", float((samples > eps).mean()), "moved",
              float(np.abs(ws[-1] - out[tag + ".w0"]).max()))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "dqn.npz"), **out)
    print("dqn golden:", len(out), "arrays")


if __name__ == "__main__":
    main()
