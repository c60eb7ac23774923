"""Episode aggregates on the device.

The reference's train loop (src/trainPPO.py:172-227) accumulates, step by step on the host, the
reward arrays, the acception quality, the prices of world.acceptedOffers per job kind and, at the
end of an episode, the normalised dwell times of world.verweilzeiten (src/world.py:350-357), and
stores per-episode means in `argsDict` (src/trainPPO.py:229-243).  Batched, the same quantities
are reduced on the device: the step kernels add the per-kind price / dwell statistics into a stats
buffer (msched_bind_stats), msched_result_sums reduces every step's result records over the
environment dimension, and `summary()` divides like the script does -- one small device->host copy
per EPISODE instead of one round trip per step.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


class EpisodeMetrics:
    """Usage:  m = EpisodeMetrics(env.core); m.begin(); ...each step: m.add()...; m.summary()."""

    def __init__(self, core):
        self.core = core
        lay, cfg = core.layout, core.cfg
        self.J = cfg.J
        self.stats = torch.zeros((lay.padded_envs, self.J, 4), dtype=torch.int32, device=core.device)
        self.totals = torch.zeros(lay.result_words + 6, dtype=torch.float64, device=core.device)
        self.stat_totals = torch.zeros((self.J, 4), dtype=torch.int64, device=core.device)
        self.steps = 0
        L.check(core.lib.msched_bind_stats(core.handle, self.stats.data_ptr()))

    def close(self):
        L.check(self.core.lib.msched_bind_stats(self.core.handle, None))

    def begin(self):
        """Start of an episode (the scripts reset their accumulators, src/trainPPO.py:140-157)."""
        self.stats.zero_()
        self.totals.zero_()
        self.stat_totals.zero_()
        self.steps = 0

    def add(self, result=None):
        """After every env.step: reduce this step's result records over the environments."""
        r = self.core.result if result is None else result
        L.check(self.core.lib.msched_result_sums(self.core.handle, r.data_ptr(), self.totals.data_ptr(),
                                                 self.core._stream()))
        self.steps += 1

    def summary(self):
        """The per-episode entries of the reference's argsDict (src/trainPPO.py:200-243), averaged
        over the environments.  One device->host copy."""
        c, lay, cfg = self.core, self.core.layout, self.core.cfg
        self.stat_totals.zero_()
        L.check(c.lib.msched_stats_sums(c.handle, self.stat_totals.data_ptr(), c._stream()))
        tot = self.totals.cpu().numpy()
        st = self.stat_totals.cpu().numpy()
        B, N, T = c.B, c.N, max(self.steps, 1)
        RW = lay.result_words

        def mean_block(off, n):
            return float(tot[off: off + n].sum()) / (T * B * n)

        out = {}
        out["coreChooserRew"] = mean_block(lay.r_offer, N * lay.RL)
        out["priceChooserRew"] = mean_block(lay.r_price, N * lay.RL) if lay.r_price >= 0 else 0.0
        out["acceptorRew"] = mean_block(lay.r_acceptor, N * lay.RC)
        out["auctioneerRew"] = float(tot[lay.r_auctioneer: lay.r_auctioneer + c.C].sum()) / (T * B)
        out["agentRew"] = (tot[lay.r_agent: lay.r_agent + N] / (T * B)).tolist()
        out["acceptionQuality"] = float(tot[RW + 4] / tot[RW + 5]) if tot[RW + 5] > 0 else None
        out["acceptionAmount"] = float(tot[RW + 0]) / (T * B)
        out["acceptedOffers"] = float(tot[RW + 1]) / (T * B)
        out["terminations"] = float(tot[RW + 2]) / (T * B)
        prios = [cfg.prio[k] for k in range(self.J)]
        lens = [cfg.len[k] for k in range(self.J)]
        out["prices"] = [float(st[k, 0]) / float(st[k, 1]) if st[k, 1] > 0 else None for k in range(self.J)]
        out["dwellTimes"] = [float(st[k, 2]) / (float(st[k, 3]) * lens[k]) if st[k, 3] > 0 else None
                             for k in range(self.J)]
        term_rev = sum(cfg.rewardMultiplier * prios[k] * int(st[k, 3]) for k in range(self.J))
        # env.terminationRevenues is only ever credited by getDividedFixedPricesReward (src/Reward.py:193)
        fixed = cfg.rewardVariant == 0
        out["terminationRevenues"] = (term_rev / B / (T * N * c.C)) if fixed else 0.0
        out["tradeRevenues"] = 0.0  # the reference never updates env.tradeRevenues (src/Reward.py:176-177)
        return out
