"""Result and checkpoint formats.

`ResultLog` collects the per-episode aggregates under exactly the keys of the reference's
`argsDict` (src/trainPPO.py:229-243) and pickles them as `data{i}.pkl` like the train scripts do
(src/trainPPO.py:245-251), so the reference's plotting scripts (src/Plot.py) read batched runs
unchanged.  `save_checkpoint` / `load_checkpoint` store the batched PPO units' parameters and
optimizer state (the reference only checkpoints its DQN nets, src/SavingAndLoading.py:4-43; the
top-level layout "agentNets"/"... optim" is kept).
"""
from __future__ import annotations

import os
import pickle

import torch

KEYS = ("acceptorRew", "coreChooserRew", "priceChooserRew", "prices", "auctioneerRew", "dwellTimes",
        "agentRew", "acceptionQuality", "acceptionAmount", "terminationRevenues", "tradeRevenues")


class ResultLog:
    def __init__(self, parameters, plot_path="", mean_job_fraction=None):
        self.args = {k: [] for k in KEYS}
        self.args["plotPath"] = plot_path
        self.args["meanJob"] = mean_job_fraction
        self.args["params"] = dict(parameters)

    def append(self, summary):
        """summary: EpisodeMetrics.summary() of one finished episode."""
        for k in KEYS:
            self.args[k].append(summary[k])

    def argsDict(self):
        return self.args

    def dump(self, file_name="data{}.pkl"):
        """First free data{i}.pkl, like src/trainPPO.py:245-251.  Returns the path."""
        i = 0
        while os.path.isfile(file_name.format(i)):
            i += 1
        path = file_name.format(i)
        with open(path, "wb") as f:
            pickle.dump(self.args, f)
        return path


def _units(agents):
    """name -> BatchedPPO of an agents object (agents.py)."""
    return {n: getattr(agents, n) for n in ("acceptor", "offer", "core", "price", "unit") if hasattr(agents, n)}


def save_checkpoint(path, world):
    top = {}
    for name, ppo in _units(world.agents).items():
        top["agentNets " + name] = {
            "actor": ppo.actor.detach().cpu(), "critic": ppo.critic.detach().cpu(),
            "policy_old": ppo.policy_old.weights.detach().cpu(),
            name + " optim": ppo.optimizer.state_dict(), "optim kind": type(ppo.optimizer).__name__,
            "step_no": ppo.step_no,
            "shape": (ppo.n_in, ppo.H, ppo.A, ppo.n_nets)}
    torch.save(top, path)


def load_checkpoint(path, world):
    top = torch.load(path, map_location="cpu", weights_only=True)  # tensors, lists and scalars only
    for name, ppo in _units(world.agents).items():
        d = top["agentNets " + name]
        if tuple(d["shape"]) != (ppo.n_in, ppo.H, ppo.A, ppo.n_nets):
            raise ValueError(f"checkpoint shape of {name} {d['shape']} does not match the agents")
        with torch.no_grad():
            ppo.actor.copy_(d["actor"].to(ppo.actor.device))
            ppo.critic.copy_(d["critic"].to(ppo.critic.device))
            ppo.policy_old.weights.copy_(d["policy_old"].to(ppo.policy_old.weights.device))
        kind = d.get("optim kind", "FlatAdam" if d[name + " optim"].get("kind") == "FlatAdam" else "Adam")
        if kind != type(ppo.optimizer).__name__:
            raise ValueError(f"checkpoint of {name} holds a {kind} optimizer state but the agents use "
                             f"{type(ppo.optimizer).__name__} (MSCHED_PPO_UPDATE selects the kernel or the autograd path)")
        ppo.optimizer.load_state_dict(d[name + " optim"])
        ppo.step_no = int(d["step_no"])
