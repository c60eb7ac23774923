"""World / Auctioneer: host-side mirror of the reference's world objects for the batched path.

Same constructor keys, attribute names and method names as reference src/world.py:210-254 and
src/Auctioneer.py:80-102, with one extra optional key `numberOfEnvironments` (default 1): the
object now stands for that many independent worlds advanced in lock-step on the GPU.  The state
itself lives in device records owned by the environment (BatchedSchedulingEnv); this class only
carries the configuration and forwards `round`, `auctioneer`, etc.
"""
from __future__ import annotations


class Auctioneer(object):
    """reference src/Auctioneer.py:80-102.  getAuctioneerAction returns the hard-coded
    auctioneer's acceptor index per idle core (HardcodedAuctioneerAcceptor,
    src/HardcodedModules.py:48-78) for every environment: int16 tensor [B, C]."""

    def __init__(self, world):
        self.world = world
        self.auctioneerID = 0

    def getAuctioneerAction(self, auctioneerObservationsTensor=None):
        env = self.world._env
        if env is None:
            raise RuntimeError("World is not attached to an environment yet")
        # the observation argument is accepted for signature compatibility; the kernel reads the
        # same information from the device state (the observation is a function of it)
        return env.core.auctioneer_action(random_ties=self.world.randomAuctioneerTies)


class World(object):
    REQUIRED = ("freePrices", "numberOfAgents", "numberOfCores", "possibleJobLengths",
                "possibleJobPriorities", "probabilities", "collectionLength",
                "newJobsPerRoundPerAgent", "episodeLength", "maxVisibleOffers", "rewardMultiplier")

    def __init__(self, params):
        for k in self.REQUIRED:
            if k not in params:
                raise KeyError(k)  # the reference raises KeyError from params[...] too
        self.params = dict(params)
        self.freePrices = params["freePrices"]
        if self.freePrices is False:
            self.listOfFixPrices = params["fixPricesList"]
        self.numberOfAgents = params["numberOfAgents"]
        self.numberOfCores = params["numberOfCores"]
        self.possibleJobLengths = params["possibleJobLengths"]
        self.possibleJobPriorities = params["possibleJobPriorities"]
        self.probabilities = params["probabilities"]
        self.accProbabilities = [sum(self.probabilities[: (i + 1)])
                                 for i in range(len(self.probabilities))]
        self.maxSumToOffer = max(self.possibleJobPriorities)
        self.collectionLength = params["collectionLength"]
        self.maxAmountOfOffers = self.numberOfAgents * self.collectionLength
        self.maxAmountOfOffersToOneAgent = self.numberOfAgents * self.collectionLength
        self.maxAmountOfAcceptionsPerTimeStepPerAgent = min(self.maxAmountOfOffersToOneAgent,
                                                            self.numberOfCores)
        self.newJobsPerRoundPerAgent = params["newJobsPerRoundPerAgent"]
        self.episodeLength = params["episodeLength"]
        self.maxVisibleOffers = params["maxVisibleOffers"]
        self.rewardMultiplier = params["rewardMultiplier"]
        # batched extensions
        self.numberOfEnvironments = int(params.get("numberOfEnvironments", 1))
        self.seed = int(params.get("seed", 0))
        self.envOffset = int(params.get("envOffset", 0))
        self.device = int(params.get("device", 0))
        self.chainCapacity = int(params.get("chainCapacity", 64))
        self.randomAuctioneerTies = bool(params.get("randomAuctioneerTies", True))
        self.agents = None
        self.randomPolicy = False
        self._env = None
        self.auctioneer = Auctioneer(self)

    @property
    def round(self):
        return self._env.core.round if self._env is not None else 0
