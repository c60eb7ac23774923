"""GPU: the hard-coded agents of BASELINE config 1 (src/HardcodedModules.py:16-45, 81-109, src/Agent.py:622-641)
on the device, against the reference's own trace and against the oracle in closed loop."""
import numpy as np
import pytest

from helpers import assert_state_equal, load_golden

pytestmark = pytest.mark.gpu


def test_hardcoded_policy_replays_reference_trace():
    """hardcoded_A is the unmodified reference running DividedHardcodedAgent + its auctioneer (seed 1).  At every
    step the device policy, fed the tie draws that reproduce the reference's random.sample choices, must return
    exactly the recorded actions: same candidate sets, same accept/reject decisions, same indices."""
    import torch
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    from oracle import oracle as O
    tr, meta = load_golden("hardcoded_A")
    N, C, L = meta["N"], meta["C"], meta["L"]
    T = tr["done"].shape[0]
    env = BatchedSchedulingEnv(1, world_params_from_dom(meta, False), reward="fix", auction="external", spawn="u64",
                               chain_capacity=32)
    obs = env.observe()
    obs_a, obs_o = tr["init_obs_acc"], tr["init_obs_off"]
    n_ties = 0
    for t in range(T):
        assert np.array_equal(obs["acceptor"][0].cpu().numpy(), obs_a) and np.array_equal(obs["offer"][0].cpu().numpy(), obs_o)
        cands = {}
        _, _, nca, nco = O.hardcoded_actions(obs_a, obs_o, cands_out=cands)
        ra, ro = tr["in_acc"][t], tr["in_offc"][t]
        u = np.zeros(N * C + N * L, np.float32)
        for (kind, i, k), cl in cands.items():
            rec = int(ra[i, k] if kind == "acc" else ro[i, k])
            assert rec in cl, (t, kind, i, k, rec, cl)       # the reference's random choice is one of OUR candidates
            u[(i * C + k) if kind == "acc" else (N * C + i * L + k)] = (cl.index(rec) + 0.5) / len(cl)
            n_ties += len(cl) > 1
        acc, off, nc = env.hardcoded_actions(u=u[None], want_ncand=True)
        assert np.array_equal(acc[0].cpu().numpy(), ra), t
        assert np.array_equal(off[0].cpu().numpy(), ro), t
        assert np.array_equal(nc[0].cpu().numpy(), np.concatenate([nca.reshape(-1), nco.reshape(-1)])), t
        env.step(off, acc, torch.as_tensor(tr["in_auc"][t][None]), spawn_u=tr["in_spawn_u"][t][None], observe=True)
        obs = env.obs_views()
        assert_state_equal({k: v[0] for k, v in env.export_state().items()}, tr, t)
        obs_a, obs_o = tr["obs_acc"][t], tr["obs_off"][t]
    assert n_ties > 100  # the trace exercises the tie-break
    env.close()


def test_hardcoded_env_closed_loop_matches_oracle():
    """HardcodedFixPriceEnvironment driven like src/trainHC.py (getActionForAllAgents, getAuctioneerAction, step):
    policy, auction and transition all on the device with Philox ties, against the oracle world + the oracle's
    restatement of the heuristic agents fed the same draws; terminationRevenues against the oracle's."""
    from marl_scheduling_b200.SchedulingEnvironment import HardcodedFixPriceEnvironment
    from marl_scheduling_b200.world import World
    from oracle import oracle as O
    dom = dict(N=2, C=3, L=2, prios=[5], lens=[4], probs=[1], fix=[3], mult=2, newJobs=1, episodeLength=50)
    wp = dict(freePrices=False, fixPricesList=[3], numberOfAgents=2, numberOfCores=3, collectionLength=2,
              possibleJobPriorities=[5], possibleJobLengths=[4], probabilities=[1], newJobsPerRoundPerAgent=1,
              rewardMultiplier=2, episodeLength=50, maxVisibleOffers=4)
    B, seed, off0 = 96, 11, 1000
    world = World(dict(wp, numberOfEnvironments=B, seed=seed, envOffset=off0))
    env = HardcodedFixPriceEnvironment(world, dict(netZeroOfferReward=0.5))
    orc = O.Oracle(B, dom, "fix", tie_mode=O.TIE_PHILOX, seed=seed, env_offset=off0)
    N, C, L = 2, 3, 2
    U = N * C + N * L
    accO, offO, aucO = env.reset()
    env.terminationRevenues = 0
    for t in range(40):
        acc, off = env.getActionForAllAgents(accO, offO)
        o_acc, o_off = np.zeros((B, N, C), np.int32), np.zeros((B, N, L), np.int32)
        for b in range(B):
            ob = orc.observe(b)
            u = O.hardcoded_draws(seed, off0 + b, t, U)
            o_acc[b], o_off[b], _, _ = O.hardcoded_actions(ob["obs_acc"], ob["obs_off"], u[: N * C].reshape(N, C),
                                                           u[N * C:].reshape(N, L))
        assert np.array_equal(acc.cpu().numpy(), o_acc), t
        assert np.array_equal(off.cpu().numpy(), o_off), t
        aa = world.auctioneer.getAuctioneerAction(aucO)
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(off, acc, aa)
        orc.step(o_off, o_acc, None)
        assert np.array_equal(aa.cpu().numpy(), orc.auc_out), t
        assert np.array_equal(agR.cpu().numpy(), orc.r_agent), t
        assert np.array_equal(accR[..., 0].cpu().numpy(), orc.r_acceptor), t
    tr_dev = env.terminationRevenues.cpu().numpy()
    assert tr_dev.shape == (B,) and tr_dev.sum() > 0
    assert np.array_equal(tr_dev, np.array([orc.export(b)["term_revenue"] for b in range(B)])), "terminationRevenues"
    e = env.core.export_state()
    for b in (0, 17, 95):
        ob = orc.export(b)
        for k in ("core_owner", "core_rem", "core_jobid", "slot_rem", "slot_jobid", "off_core", "chain_len"):
            assert np.array_equal(np.asarray(e[k][b]), np.asarray(ob[k])), (b, k)
    env.close()
