"""Reference harness: stub-import lr40/marl-scheduling and record step traces.

TEST INFRASTRUCTURE ONLY.  This module imports the *unmodified* reference from
/root/reference/src (read-only, present only in the build container, never on
the GPU box) and drives it with recorded spawn draws and action streams so that
the C restatement in oracle/msched_oracle.c and the CUDA path can be pinned
against the reference itself.  Nothing in the product package imports it.

The recipe follows SURVEY.md Appendix C: three missing third-party modules
(gym, matplotlib.pyplot, seaborn) are stubbed in sys.modules, the class-level
ID counters are reset by hand before every World(...), and random.random is
monkey-patched to script / record the spawn draws.
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np

# the read-only mount in the build container, else the byte-for-byte copy oracle/make_ref.py left in
# oracle/_ref (git-ignored; ships with the snapshot so that bench.py can time the reference on the GPU box)
REFERENCE_SRC = "/root/reference/src"
if not os.path.isdir(REFERENCE_SRC):
    REFERENCE_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_SRC, "SchedulingEnvironment.py"))

_mods = None


def import_reference():
    """Stub gym/matplotlib/seaborn and import the reference modules once."""
    global _mods
    if _mods is not None:
        return _mods
    gym = types.ModuleType("gym")
    gym.Env = type("Env", (), {"close": lambda self: None})
    for n in ("error", "spaces", "utils"):
        m = types.ModuleType("gym." + n)
        setattr(gym, n, m)
        sys.modules["gym." + n] = m
    sys.modules["gym"] = gym
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    sys.modules["seaborn"] = types.ModuleType("seaborn")
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import SchedulingEnvironment as SE  # noqa: E402
    import world as W  # noqa: E402
    import Agent as A  # noqa: E402
    import Reward as R  # noqa: E402
    import HardcodedModules as H  # noqa: E402
    import PPOmodules as P  # noqa: E402

    _mods = types.SimpleNamespace(SE=SE, W=W, A=A, R=R, H=H, P=P)
    return _mods


def reset_ids():
    m = import_reference()
    m.W.Core.IDCounter = 1
    m.W.Job.IDCounter = 1
    m.A.Agent.IDCounter = 1
    m.W.Offer.offerID = 1


REWARD_MODES = ("fix", "free_comm", "free_ncomm", "agg")


def make_env(dom, reward_mode, agent_kind="divided"):
    """dom: dict(N,C,L,prios,lens,probs,fix,mult,newJobs,episodeLength)."""
    m = import_reference()
    free = reward_mode.startswith("free")
    wp = dict(
        num_episodes=1,
        episodeLength=dom.get("episodeLength", 100),
        numberOfAgents=dom["N"],
        numberOfCores=dom["C"],
        possibleJobPriorities=list(dom["prios"]),
        possibleJobLengths=list(dom["lens"]),
        collectionLength=dom["L"],
        probabilities=list(dom["probs"]),
        newJobsPerRoundPerAgent=dom.get("newJobs", 1),
        rewardMultiplier=dom.get("mult", 1),
        freePrices=free,
        fixPricesList=list(dom["fix"]),
        maxVisibleOffers=4,
    )
    reset_ids()
    world = m.W.World(wp)

    SE, A, R = m.SE, m.A, m.R

    class ScriptedEnv(SE.SchedulingEnv):
        def __init__(self, world, params):
            super().__init__(world, params)
            cls = {"divided": A.DividedAgent, "aggregated": A.AggregatedAgent,
                   "hardcoded": A.DividedHardcodedAgent}[agent_kind]
            self.world.agents = [cls(world) for _ in range(world.numberOfAgents)]
            self.terminationRevenues = 0
            self.tradeRevenues = 0

        def getRewards(self):
            if reward_mode == "fix":
                return R.getDividedFixedPricesReward(self)
            if reward_mode == "free_comm":
                return R.getDividedFreePricesReward(self, True)
            if reward_mode == "free_ncomm":
                return R.getDividedFreePricesReward(self, False)
            if reward_mode == "agg":
                return R.getAggregatedFixedPricesReward(self)
            raise ValueError(reward_mode)

    env = ScriptedEnv(world, dict(netZeroOfferReward=dom.get("netZero", 0.5)))
    return world, env


class SpawnRecorder:
    """Patches random.random; records (agent, k, u) for every spawn draw.

    `source` is a callable returning the next u in [0,1) (scripted KAT sequence
    or the real Mersenne Twister).  The agent is tracked by wrapping
    Agent.fillCollectionRandomly (src/Agent.py:50-70)."""

    def __init__(self, source=None):
        m = import_reference()
        self.m = m
        self._orig_random = random.random
        self._orig_fill = m.A.Agent.fillCollectionRandomly
        self.source = source if source is not None else self._orig_random
        self.cur_agent = None
        self.cur_k = 0
        self.draws = []  # list of (agentIndex, k, u) for the current step
        self.total = 0

    def __enter__(self):
        rec = self

        def patched_random():
            u = rec.source()
            if rec.cur_agent is not None:
                rec.draws.append((rec.cur_agent, rec.cur_k, u))
                rec.cur_k += 1
                rec.total += 1
            return u

        orig_fill = self._orig_fill

        def patched_fill(agent_self, amountOfNewEntries):
            rec.cur_agent = agent_self.agentID - 1
            rec.cur_k = 0
            try:
                return orig_fill(agent_self, amountOfNewEntries)
            finally:
                rec.cur_agent = None

        random.random = patched_random
        self.m.A.Agent.fillCollectionRandomly = patched_fill
        return self

    def __exit__(self, *exc):
        random.random = self._orig_random
        self.m.A.Agent.fillCollectionRandomly = self._orig_fill
        return False

    def take(self):
        d, self.draws = self.draws, []
        return d


def kat_u_source():
    """u_k = ((k*2654435761) mod 2^32)/2^32, k = 1,2,...."""
    k = [0]

    def src():
        k[0] += 1
        return ((k[0] * 2654435761) % (1 << 32)) / float(1 << 32)

    return src


def first_max_auctioneer(obs, NL):
    """RNG-free auctioneer of the KAT protocol: FIRST arg-max of price/time."""
    out = []
    for o in obs:
        o = o.tolist()
        if o[0] == 0:
            out.append(NL)
            continue
        ratios = [(-1 if (p in (-1, -2) or t in (-1, -2)) else p / t)
                  for p, t in zip(o[3::2], o[4::2])]
        mx = max(ratios)
        out.append(ratios.index(mx) if mx > -1 else NL)
    return out


def _job_fields(job):
    empty = bool(job.empty)
    return dict(
        prio=job.priority, rem=job.remainingLength, jobid=job.jobID, kind=job.jobKind,
        wait=int(bool(job.wait)),
        birth=(-1 if empty else job.birthDate),
        init=(-1 if empty else job.initialLength),
    )


def snapshot(world, env, K):
    """Full integer state of one world after a step (reference-shaped)."""
    N, C, L = world.numberOfAgents, world.numberOfCores, world.collectionLength
    s = {}
    s["core_owner"] = np.array([c.ownerID for c in world.cores], np.int32)
    cf = [_job_fields(c.job) for c in world.cores]
    for k in ("prio", "rem", "jobid", "kind", "birth", "init"):
        s["core_" + k] = np.array([f[k] for f in cf], np.int32)
    sf = [[_job_fields(j) for j in ag.collection] for ag in world.agents]
    for k in ("prio", "rem", "jobid", "kind", "wait", "birth", "init"):
        s["slot_" + k] = np.array([[f[k] for f in row] for row in sf], np.int32)
    s["free_slots"] = np.array([ag.collection.numberOfFreeSlots for ag in world.agents], np.int32)
    off = np.zeros((5, N, L), np.int32)  # core, recip, price, time, id ; core 0 = none
    off[1:] = -1
    for f in world.offers:
        i, q = f.offererID - 1, f.queuePosition
        off[:, i, q] = (f.coreID, f.recipientID, f.offeredReward, f.necessaryTime, f.offerID)
    s["off_core"], s["off_recip"], s["off_price"], s["off_time"], s["off_id"] = off
    chain = np.full((C, K, 5), -1, np.int32)
    clen = np.zeros(C, np.int32)
    for c in range(C):
        dq = world.liabilityList[c]
        clen[c] = len(dq)
        if len(dq) > K:
            raise RuntimeError("chain longer than recording capacity %d" % K)
        for k, e in enumerate(dq):  # newest first
            chain[c, k] = (e.offererID, e.recipientID, e.offeredReward, e.necessaryTime, e.round)
    s["chain"], s["chain_len"] = chain, clen
    s["round"] = np.int32(world.round)
    return s


def record_trace(dom, reward_mode, steps, policy, u_source=None, agent_kind="divided",
                 K=24, seed=None):
    """Run the reference `steps` steps and record everything.

    policy(step, world, env, obsA, obsO, obsAuc) -> (acc[N][C], off[N][L] or tuples, auc[C])
    """
    m = import_reference()
    if seed is not None:
        random.seed(seed)
    world, env = make_env(dom, reward_mode, agent_kind)
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL = N * L
    nj = dom.get("newJobs", 1)
    free = reward_mode.startswith("free")
    agg = reward_mode == "agg"
    rows = []
    with SpawnRecorder(u_source) as rec:
        a, o, au = env.reset()
        init = dict(snapshot(world, env, K))
        init.update(_obs_arrays(env, a, o, au, agent_kind, N, C, L))
        for s in range(steps):
            acc, off, auc = policy(s, world, env, a, o, au)
            a, o, au, offR, accR, aucR, agR, qual, done = env.step(off, acc, auc)
            draws = rec.take()
            row = dict(snapshot(world, env, K))
            row["in_acc"] = np.array(acc, np.int32).reshape(N, C)
            if free:
                row["in_offc"] = np.array([[t[0] for t in r] for r in off], np.int32)
                row["in_offp"] = np.array([[t[1] for t in r] for r in off], np.int32)
            else:
                row["in_offc"] = np.array(off, np.int32).reshape(N, L)
                row["in_offp"] = np.zeros((N, L), np.int32)
            row["in_auc"] = np.array(auc, np.int32)
            su = np.full((N, nj), np.nan, np.float64)
            for (ai, k, u) in draws:
                su[ai, k] = u
            row["in_spawn_u"] = su
            if free:
                row["r_offer"] = np.asarray(offR[0], np.float64).reshape(N, L)
                row["r_price"] = np.asarray(offR[1], np.float64).reshape(N, L)
            elif agg:
                row["r_offer"] = np.asarray(offR, np.float64).reshape(N, 1)
                row["r_price"] = np.zeros((N, 1), np.float64)
            else:
                row["r_offer"] = np.asarray(offR, np.float64).reshape(N, L)
                row["r_price"] = np.zeros((N, L), np.float64)
            row["r_acceptor"] = (np.asarray(accR, np.int64).reshape(N, 1) if agg
                                 else np.asarray(accR, np.int64).reshape(N, C))
            row["r_auctioneer"] = np.asarray(aucR, np.int64).reshape(C)
            row["r_agent"] = np.asarray(agR, np.int64).reshape(N)
            row["quality"] = np.float64(np.nan if qual[0] is None else qual[0])
            row["quality_cnt"] = np.int32(qual[1])
            row["done"] = np.int32(bool(done))
            accd = np.full((C, 5), -1, np.int32)
            for k, f in enumerate(world.acceptedOffers):
                accd[k] = (f.offererID, f.recipientID, f.coreID, f.queuePosition, f.offeredReward)
            row["accepted"] = accd
            row["n_accepted"] = np.int32(len(world.acceptedOffers))
            term = np.full((C, 5), -1, np.int32)
            for k, (c, own, jid, gr, rd) in enumerate(world.jobTerminationInfo):
                term[k] = (c.coreID, own, jid, gr, rd)
            row["term"] = term
            row["n_term"] = np.int32(len(world.jobTerminationInfo))
            row["term_revenue"] = np.int64(env.terminationRevenues)
            row.update(_obs_arrays(env, a, o, au, agent_kind, N, C, L))
            rows.append(row)
        total_draws = rec.total
    out = {k: np.stack([r[k] for r in rows]) for k in rows[0]}
    for k, v in init.items():
        out["init_" + k] = np.asarray(v)
    # dwell-time records (src/world.py:350-357), in order of occurrence
    vz = world.verweilzeiten
    out["dwell"] = np.array([[v.Prioritaet, v.Bedienzeit, v.Verweilzeit] for v in vz],
                            np.int32).reshape(-1, 3)
    out["dwell_norm"] = np.array([v.normalisierte_Verweilzeit for v in vz], np.float64)
    out["total_draws"] = np.int32(total_draws)
    out["last_jobid"] = np.int32(m.W.Job.IDCounter - 1)
    out["cum_prob"] = np.array(world.accProbabilities, np.float64)
    return out


def _obs_arrays(env, a, o, au, agent_kind, N, C, L):
    """Dense observations + offer-ID tables as returned by the reference."""
    NL = N * L
    r = {}
    if agent_kind in ("divided", "hardcoded"):
        r["obs_acc"] = np.array([[t.tolist() for t in row] for row in a], np.int32).reshape(
            N, C, 3 + 2 * NL)
        r["obs_off"] = np.array([[t.tolist() for t in row] for row in o], np.int32).reshape(
            N, L, 2 * C + 2)
    else:  # aggregated layouts (src/Agent.py:82-140): float32 [C(3+2NL)], int64 [2C+2L]
        r["obs_acc"] = np.array([t.tolist() for t in a], np.float32).reshape(N, C * (3 + 2 * NL))
        r["obs_off"] = np.array([t.tolist() for t in o], np.int32).reshape(N, 2 * C + 2 * L)
    r["obs_auc"] = np.array([t.tolist() for t in au], np.int32).reshape(C, 3 + 2 * NL)
    r["ids"] = np.array(env.correspondingOfferIDs, np.int32).reshape(N, C, NL)
    r["auc_ids"] = np.array(env.auctioneer_correspondingOfferIDs, np.int32).reshape(C, NL)
    return r


# ---------------------------------------------------------------- policies
def kat_policy(dom, free):
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL = N * L
    P = max(dom["prios"])

    def pol(s, world, env, a, o, au):
        acc = [[(s * 7 + i * 3 + j * 5) % (NL + 1) for j in range(C)] for i in range(N)]
        if free:
            off = [[((s * 5 + i * 2 + q * 3) % (C + 1), (s * 3 + i + q * 2) % (P + 1))
                    for q in range(L)] for i in range(N)]
        else:
            off = [[(s * 5 + i * 2 + q * 3) % (C + 1) for q in range(L)] for i in range(N)]
        return acc, off, first_max_auctioneer(au, NL)

    return pol


def random_policy(dom, free, rng, p_valid=0.6, p_weird=0.03, auctioneer="reference"):
    """Uniform-random actions biased towards valid acceptances so that chains grow.

    rng is a private numpy Generator (does not touch the global `random` stream that
    the reference uses for spawn draws and auctioneer tie-breaks)."""
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL = N * L
    P = max(dom["prios"])

    def pol(s, world, env, a, o, au):
        acc = []
        for i in range(N):
            row = []
            for j in range(C):
                ids = env.correspondingOfferIDs[i][j]
                nvalid = sum(1 for x in ids if x > 0)
                if nvalid and rng.random() < p_valid:
                    row.append(int(rng.integers(0, nvalid)))
                else:
                    row.append(int(rng.integers(0, NL + 1)))
            acc.append(row)
        off = []
        for i in range(N):
            row = []
            for q in range(L):
                c = int(rng.integers(0, C + 1))
                if rng.random() < p_weird:
                    c = int(rng.choice([-1, C + 1, C + 3]))
                if free:
                    p = int(rng.integers(0, P + 1))
                    if rng.random() < p_weird:
                        p = int(rng.choice([-5, -1, -2, P + 3]))
                    row.append((c, p))
                else:
                    row.append(c)
            off.append(row)
        if auctioneer == "reference":
            auc = world.auctioneer.getAuctioneerAction(au)  # random tie-break, recorded
        else:
            auc = first_max_auctioneer(au, NL)
        if rng.random() < 0.1:  # occasionally a scripted / padding index instead
            j = int(rng.integers(0, C))
            auc = list(auc)
            auc[j] = int(rng.integers(0, NL + 1))
        return acc, off, auc

    return pol


def hardcoded_policy():
    """The reference's own heuristic agents + auctioneer (config 1)."""

    def pol(s, world, env, a, o, au):
        acc, off = env.getActionForAllAgents(a, o)
        auc = world.auctioneer.getAuctioneerAction(au)
        return acc, off, auc

    return pol
