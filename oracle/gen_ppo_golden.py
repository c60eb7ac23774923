"""Generate tests/golden/ppo_update.npz by running the UNMODIFIED reference PPO.update (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/gen_ppo_golden.py

For three net shapes of the BASELINE configs (acceptor 15->16->16->7 and price chooser 4->16->16->9 of
cfg3, acceptor 27->16->16->13 of cfg2) a reference `PPO` (src/PPOmodules.py:75-174) selects actions for T
integer observation rows (PPO.selectAction, global torch RNG seeded), receives integer rewards and runs
`update()`.  Recorded: initial actor / critic parameters in the C-ABI's flat layout, the buffer (states,
actions, log-probs, rewards), hyper-parameters, and the parameters after the update.  The GPU box never sees
/root/reference; tests/test_gpu_ppo_update.py replays the buffer through msched_returns / msched_ppo_grad /
msched_adam_step.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, "/root/reference/src")
import PPOmodules as RP  # noqa: E402  (imports torch only)


def flat(seq):
    parts = []
    for i in (0, 2, 4):
        parts += [seq[i].weight.detach().reshape(-1), seq[i].bias.detach().reshape(-1)]
    return torch.cat(parts).numpy().astype(np.float32)


def main():
    out = {}
    cases = [("acc_cfg3", 15, 7, 200, 0.9, 0.2, 5, 3e-4, 1e-3), ("price_cfg3", 4, 9, 200, 0.5, 0.2, 8, 3e-4, 1e-3),
             ("acc_cfg2", 27, 13, 120, 0.95, 0.1, 3, 1e-3, 3e-3)]
    for tag, n_in, A, T, gamma, eps_clip, K, lr_a, lr_c in cases:
        torch.manual_seed(len(tag) * 31 + n_in)
        rng = np.random.default_rng(n_in * 7 + A)
        ppo = RP.PPO(None, n_in, A, lr_a, lr_c, gamma, eps_clip, K, 16)
        out[tag + ".actor0"], out[tag + ".critic0"] = flat(ppo.policy.actor), flat(ppo.policy.critic)
        states = rng.integers(-2, 9, (T, n_in)).astype(np.int16)
        rewards = rng.integers(-6, 11, T).astype(np.int64)
        for t in range(T):
            ppo.selectAction(torch.tensor(states[t]))
            ppo.buffer.rewards.append(int(rewards[t]))
        out[tag + ".states"] = states
        out[tag + ".actions"] = torch.stack(ppo.buffer.actions).numpy().astype(np.int32)
        out[tag + ".logprobs"] = torch.stack(ppo.buffer.logprobs).numpy().astype(np.float32)
        out[tag + ".rewards"] = rewards.astype(np.int32)
        ppo.update()
        out[tag + ".actor1"], out[tag + ".critic1"] = flat(ppo.policy.actor), flat(ppo.policy.critic)
        for k, v in (("n_in", n_in), ("A", A), ("K", K), ("gamma", gamma), ("eps_clip", eps_clip),
                     ("lr_actor", lr_a), ("lr_critic", lr_c)):
            out[tag + "." + k] = np.float64(v)
        moved = np.abs(out[tag + ".actor1"] - out[tag + ".actor0"]).max()
        print(f"{tag}: T={T} K={K} max |delta actor| = {moved:.3e}")
    np.savez_compressed(os.path.join(GOLDEN, "ppo_update.npz"), **out)


if __name__ == "__main__":
    main()
