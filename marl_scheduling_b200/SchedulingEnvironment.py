"""Drop-in batched counterparts of the reference's environment classes.

Class names, constructor arguments, method names, argument ORDER and the 9-tuple returned by
`step` follow reference src/SchedulingEnvironment.py:21-456; every nested Python list of
per-agent / per-unit tensors becomes ONE device tensor with a leading environment dimension:

  reset() -> (acceptorObs [B,N,C,3+2NL], offerObs [B,N,L,2C+2], auctioneerObs [B,C,3+2NL])  int16
  step(offerActions, acceptorActions, auctioneer_action) ->
      (acceptorObs, offerObs, auctioneerObs,
       offerRewards, acceptorRewards, auctioneerReward [B,C], agentReward [B,N],
       (acceptionQualityMean [B] (NaN where the reference returns None), amount [B]), done)

  offerRewards    fixed prices: float32 [B,N,L,1]; free prices: (coreChooser, priceChooser) pair;
                  aggregated: [B,N,1]                                    (src/Reward.py)
  acceptorRewards int32 [B,N,C,1]; aggregated: [B,N,1]
  offerActions    [B,N,L] core indices (C = no offer); free prices: ([B,N,L] cores, [B,N,L]
                  prices) or a [B,N,L,2] tensor                          (src/world.py:406-478)
  acceptorActions [B,N,C] table indices (N*L = reject)                   (src/world.py:391-404)
  auctioneer_action [B,C] from world.auctioneer.getAuctioneerAction(...), or None to run the
                  hard-coded auction inside the step kernel (same rule, one launch fewer)
"""
from __future__ import annotations

import torch

from .batched_env import BatchedSchedulingEnv


class SchedulingEnv(object):
    REWARD = "fix"  # which src/Reward.py function getRewards() stands for

    def __init__(self, world, params, reward=None):
        self.world = world
        self.netZeroOfferReward = params["netZeroOfferReward"]
        self.reward = reward or self.REWARD
        wp = dict(world.params)
        self._auction_external = bool(params.get("externalAuctioneer", True))
        self.core = BatchedSchedulingEnv(
            world.numberOfEnvironments, wp, reward=self.reward,
            auction="external" if self._auction_external else
            ("random" if world.randomAuctioneerTies else "first"),
            spawn="philox", chain_capacity=world.chainCapacity, seed=world.seed,
            env_offset=world.envOffset, net_zero_offer_reward=self.netZeroOfferReward,
            device=world.device)
        world._env = self
        self.tradeRevenues = 0
        self.terminationRevenues = 0
        self._last = None

    # -- reference API -------------------------------------------------------------------------
    def _format_obs(self, o):
        """(acceptorObs, offerObs, auctioneerObs) as the agents of this environment consume them."""
        return o["acceptor"], o["offer"], o["auctioneer"]

    def reset(self):
        """Like the reference (src/SchedulingEnvironment.py:85-109) this resets NOTHING: it only
        re-gathers the observations of the current state."""
        return self._format_obs(self.core.observe())

    def step(self, offerActions, acceptorActions, auctioneer_action=None):
        c = self.core
        price = None
        if c.free:
            if isinstance(offerActions, (tuple, list)) and len(offerActions) == 2:
                offerActions, price = offerActions
            else:
                offerActions = torch.as_tensor(offerActions)
                offerActions, price = offerActions[..., 0], offerActions[..., 1]
        if self._auction_external:
            if auctioneer_action is None:
                auctioneer_action = c.auctioneer_action(self.world.randomAuctioneerTies)
        elif auctioneer_action is not None:
            raise ValueError("this environment runs the auction in-kernel; pass None")
        # one C-ABI call: the transition and the observations of the new state (fused launch)
        r = c.step(offerActions, acceptorActions, auctioneer_action, offer_price=price, observe=True)
        self._last = r
        o = c.obs_views()
        offerRewards, acceptorRewards, auctioneerReward, agentReward = self.getRewards()
        cnt = r["quality_cnt"]
        mean = torch.where(cnt > 0, r["quality_sum"] / cnt.clamp(min=1).double(),
                           torch.full_like(r["quality_sum"], float("nan")))
        done = (c.round % self.world.episodeLength) == 0
        if self.reward == "fix":  # src/Reward.py:193: env.terminationRevenues += generatedReward
            self.terminationRevenues = self.terminationRevenues + self._termination_revenue(r)
        return (*self._format_obs(o), offerRewards, acceptorRewards,
                auctioneerReward, agentReward, (mean, cnt), done)

    def getRewards(self):
        r = self._last
        if self.core.agg:  # np.array([[0] for agent ...]) in src/Reward.py:94-95 -> [B,N,1]
            return r["offer"], r["acceptor"], r["auctioneer"], r["agent"]
        off = r["offer"].unsqueeze(-1)
        if self.core.free:
            off = (off, r["price"].unsqueeze(-1))
        return off, r["acceptor"].unsqueeze(-1), r["auctioneer"], r["agent"]

    def _termination_revenue(self, r):
        """Sum of generatedReward over this step's terminations, per environment (int64 [B]).  In
        getDividedFixedPricesReward (src/Reward.py:186-208) every termination credits generatedReward to its
        owner's agentReward; every liability-chain entry then moves tradedReward from the offerer's agentReward
        to the recipient's, or -- for the one entry the auctioneer accepted -- to auctioneerReward[core] (a core
        terminates at most once per step, so that "=" never overwrites).  Trades between agents cancel, hence
        sum(agentReward) + sum(auctioneerReward) = sum(generatedReward), exactly."""
        return r["agent"].sum(dim=1, dtype=torch.int64) + r["auctioneer"].sum(dim=1, dtype=torch.int64)

    def render(self, mode="human"):
        e = self.core.export_state(0, 1)
        print("___________________________________")
        print("Round:", self.core.round)
        for k in ("core_owner", "core_prio", "core_rem", "core_jobid", "slot_prio", "slot_rem",
                  "slot_jobid", "off_core", "off_recip", "off_price"):
            print(k, e[k][0].tolist())

    def close(self):
        self.core.close()

    def calculateAverageAcceptionQuality(self):
        r = self._last
        return r["quality_sum"] / r["quality_cnt"].clamp(min=1).double(), r["quality_cnt"]


class DividedHardcodedAgents(object):
    """All N DividedHardcodedAgent objects of a world (src/Agent.py:622-641): per (agent, core) a
    HardcodedAcceptor, per (agent, slot) a HardcodedOfferer (src/HardcodedModules.py:16-45, 81-109), evaluated
    for every environment by ONE launch (msched_hardcoded_actions) that writes the env's action record."""

    def __init__(self, world, env):
        self.world, self.env = world, env

    def getActions(self, offerObs=None, acceptorObs=None):
        """The observations are accepted for signature compatibility; the kernel reads the observation
        record the last reset()/step() wrote (the tensors handed in are views into it)."""
        acc, off = self.env.core.hardcoded_actions(random_ties=self.world.randomAuctioneerTies)
        return off, acc  # (offerNetActions, acceptorNetActions), src/Agent.py:641


class HardcodedFixPriceEnvironment(SchedulingEnv):
    """src/SchedulingEnvironment.py:439-456 (getDividedFixedPricesReward): BASELINE config 1, the
    heuristic agents of src/trainHC.py."""
    REWARD = "fix"

    def __init__(self, world, params):
        super().__init__(world, params)
        self.agents = DividedHardcodedAgents(world, self)
        self.world.agents = self.agents

    def getActionForAllAgents(self, acceptorObs=None, offerObs=None):
        """src/SchedulingEnvironment.py:150-172; returns (acceptorActions [B,N,C], offerActions [B,N,L])."""
        off, acc = self.agents.getActions(offerObs, acceptorObs)
        return acc, off

    def saveRewards(self, offerNetRewards, acceptorNetRewards, agentReward):
        ...

    def updateAgents(self):
        ...


class PPOSchedulingEnv(SchedulingEnv):
    """src/SchedulingEnvironment.py:195-210: carries the RL hyper-parameters."""

    def __init__(self, world, params, reward=None):
        super().__init__(world, params, reward)
        self.LR_ACTOR = params["LR_ACTOR"]
        self.LR_CRITIC = params["LR_CRITIC"]
        self.OFFER_GAMMA = params["OFFER_GAMMA"]
        self.ACCEPTOR_GAMMA = params["ACCEPTOR_GAMMA"]
        self.EPS_CLIP = params["EPS_CLIP"]
        self.RAW_K_EPOCHS = params["RAW_K_EPOCHS"]
        self.ACCEPTOR_K_EPOCHS = params["ACCEPTOR_K_EPOCHS"]
        self.OFFER_K_EPOCHS = params["OFFER_K_EPOCHS"]
        self.CENTRALISATION_SAMPLE = params["CENTRALISATION_SAMPLE"]
        self.agents = None

    def _attach(self, agents):
        self.agents = agents
        self.world.agents = agents

    def getActionForAllAgents(self, acceptorObs, offerObs):
        """src/SchedulingEnvironment.py:150-172; note the (acceptor, offer) return order."""
        return self.agents.getActions(offerObs, acceptorObs)

    def saveRewards(self, offerRewards, acceptorRewards, agentReward):
        self.agents.saveRewards(offerRewards, acceptorRewards, agentReward)

    def updateAgents(self):
        self.agents.updateParts()


class PPODividedFixedPriceEnv(PPOSchedulingEnv):
    """src/SchedulingEnvironment.py:274-291."""
    REWARD = "fix"

    def __init__(self, world, params):
        super().__init__(world, params)
        from .agents import DividedFixedPricePPOAgents
        self._attach(DividedFixedPricePPOAgents(world, self))


class PPODividedFreePriceEnv(PPOSchedulingEnv):
    """src/SchedulingEnvironment.py:253-271."""

    def __init__(self, world, params, commercialFreePriceReward):
        super().__init__(world, params, "free_comm" if commercialFreePriceReward else "free_ncomm")
        self.commercialFreePriceReward = commercialFreePriceReward
        from .agents import DividedFreePricePPOAgents
        self._attach(DividedFreePricePPOAgents(world, self))


class GloballySharedParamsDividedFixedPriceEnv(PPOSchedulingEnv):
    """src/SchedulingEnvironment.py:294-329: one acceptor net and one offer net for all units."""
    REWARD = "fix"

    def __init__(self, world, params):
        super().__init__(world, params)
        from .agents import DividedFixedPricePPOAgents
        self._attach(DividedFixedPricePPOAgents(world, self, sharing="global"))


class LocallySharedParamsDividedFixedPriceEnv(PPOSchedulingEnv):
    """src/SchedulingEnvironment.py:332-348: one acceptor net and one offer net per agent."""
    REWARD = "fix"

    def __init__(self, world, params):
        super().__init__(world, params)
        from .agents import DividedFixedPricePPOAgents
        self._attach(DividedFixedPricePPOAgents(world, self, sharing="local"))


class PPOAggregatedFixPriceEnv(PPOSchedulingEnv):
    """src/SchedulingEnvironment.py:213-228 (getAggregatedFixedPricesReward): semi-aggregated agents.
    reset/step return the aggregated observations of src/Agent.py:82-140: acceptor int16
    [B,N,C*(3+2NL)] (float32 in the reference), offer int16 [B,N,2C+2L], auctioneer [B,C,3+2NL]."""
    REWARD = "agg"

    def __init__(self, world, params, agents=True):
        super().__init__(world, params)
        if agents:
            from .agents import AggregatedFixPricePPOAgents
            self._attach(AggregatedFixPricePPOAgents(world, self))

    def _aggregate(self, o):
        B, N, C, L = self.core.B, self.core.N, self.core.C, self.core.Lc
        acc = o["acceptor"].reshape(B, N, -1)
        cores = o["offer"][:, :, 0, : 2 * C]
        slots = o["offer"][:, :, :, 2 * C:].reshape(B, N, 2 * L)
        return acc, torch.cat([cores, slots], dim=2)

    def _format_obs(self, o):
        acc, off = self._aggregate(o)
        return acc, off, o["auctioneer"]

    def aggregatedObservations(self):
        """(acceptor float32 [B,N,C*(3+2NL)], offer int16 [B,N,2C+2L]) of the current state."""
        acc, off = self._aggregate(self.core.observe())
        return acc.float(), off


class PPOFullyAggregatedFixPriceEnv(PPOAggregatedFixPriceEnv):
    """src/SchedulingEnvironment.py:231-250."""

    def __init__(self, world, params, agents=True):
        super().__init__(world, params, agents=False)
        if agents:
            from .agents import FullyAggregatedFixPricePPOAgents
            self._attach(FullyAggregatedFixPricePPOAgents(world, self))

    def fullyAggregatedObservations(self):
        acc, off = self.aggregatedObservations()
        return torch.cat([off.float(), acc], dim=2)  # src/Agent.py:463-467


class DQNSchedulingEnv(SchedulingEnv):
    """src/SchedulingEnvironment.py:351-425: carries the DQN hyper-parameters and the
    update*MemoriesAndOptimize methods (push one transition per environment and unit, then one
    optimize_model step per unit)."""
    REWARD = "fix"

    def __init__(self, world, params):
        super().__init__(world, params)
        self.RUN_END = params["RUN_END"]
        self.RUN_START = params["RUN_START"]
        self.RUN_DECAY = params["RUN_DECAY"]
        self.BATCH_SIZE = params["BATCH_SIZE"]
        self.OFFER_GAMMA = params["OFFER_GAMMA"]
        self.ACCEPTOR_GAMMA = params["ACCEPTOR_GAMMA"]
        self.REPLAY_MEMORY_SIZE = params["REPLAY_MEMORY_SIZE"]
        self.agents = None

    def getActionForAllAgents(self, acceptorObs, offerObs):
        return self.agents.getActions(offerObs, acceptorObs)

    def _push_and_optimize(self, dqn, old_obs, actions, new_obs, rewards):
        B = self.core.B
        dqn.memory.push(old_obs.reshape(B, dqn.units, dqn.n_in), actions.reshape(B, dqn.units),
                        new_obs.reshape(B, dqn.units, dqn.n_in), rewards.reshape(B, dqn.units))
        return dqn.optimize_model(self.BATCH_SIZE)

    def updateOfferMemoriesAndOptimize(self, oldOfferObs, offerActions, newOfferObs, offerNetRewards):
        return self._push_and_optimize(self.agents.offer, oldOfferObs, offerActions, newOfferObs, offerNetRewards)

    def updateAcceptorMemoriesAndOptimize(self, oldAcceptorObs, acceptorActions, newAcceptorObs, acceptorNetRewards):
        return self._push_and_optimize(self.agents.acceptor, oldAcceptorObs, acceptorActions, newAcceptorObs,
                                       acceptorNetRewards)


class DQNDividedFixedPricesEnv(DQNSchedulingEnv):
    """src/SchedulingEnvironment.py:428-436."""

    def __init__(self, world, params):
        super().__init__(world, params)
        from .dqn import DividedFixPriceDQNAgents
        self.agents = DividedFixPriceDQNAgents(world, self)
        self.world.agents = self.agents


def numberToNDimensionalAction(number, base, dimensionality):
    """src/Agent.py:644-666 for tensors: digit k = (number // base**k) % base, index 0 = least
    significant.  Raises ValueError("Illegal Argument") like the reference."""
    number = torch.as_tensor(number)
    if bool(((number < 0) | (number >= base ** dimensionality)).any()):
        raise ValueError("Illegal Argument")
    return torch.stack([(number // (base ** k)) % base for k in range(dimensionality)], dim=-1)
