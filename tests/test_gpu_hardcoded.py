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


@pytest.mark.parametrize("obs_every,ties,dev_round,B", [(False, True, False, 900), (True, True, True, 900), (True, False, False, 900),
                                                        (True, True, True, 65536)])  # the last: BASELINE configs[0]'s domain at the full batch
def test_rollout_hardcoded_one_launch_matches_two_launch_loop(obs_every, ties, dev_round, B):
    """msched_rollout_hardcoded (config 1: T x (step ; hard-coded agents) inside ONE launch, the agents reading the
    observation tile in shared memory) against the loop of msched_step_observe + msched_hardcoded_actions: result
    records of every step, final state, chains, observations and the next action record, bit for bit; two launches."""
    import torch
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    dom = dict(N=2, C=3, L=2, prios=[5], lens=[4], probs=[1], fix=[3], mult=2)
    T = 11
    mk = lambda: BatchedSchedulingEnv(B, world_params_from_dom(dom, False), reward="fix", auction="random", spawn="philox", seed=5)
    a, b = mk(), mk()
    if dev_round:
        a.set_device_round(True); b.set_device_round(True)
    lay, dev = a.layout, a.device
    for e in (a, b):
        e.observe()
        e.hardcoded_actions(random_ties=ties)
    for launch in range(2):
        res1 = torch.zeros((T, lay.padded_envs, lay.result_words), dtype=torch.int32, device=dev)
        obs1 = torch.zeros((T, lay.padded_envs, lay.obs_halfs), dtype=torch.int16, device=dev)
        for t in range(T):
            a.step_observe_records(a.action, res1[t], obs=obs1[t])
            a.hardcoded_actions(obs=obs1[t], random_ties=ties)
        res2 = torch.zeros_like(res1)
        obs2 = torch.zeros_like(obs1) if obs_every else torch.zeros_like(obs1[0])
        b.rollout_hardcoded(res2, obs2, obs_every=obs_every, random_ties=ties)
        torch.cuda.synchronize()
        assert torch.equal(res1[:, :B], res2[:, :B]), launch
        assert torch.equal(a.state[:B], b.state[:B]) and torch.equal(a.chain[:B], b.chain[:B])
        assert torch.equal(obs1[:, :B], obs2[:, :B]) if obs_every else torch.equal(obs1[T - 1, :B], obs2[:B])
        NC, NL = a.N * a.C, a.NL
        assert torch.equal(a.action[:B, lay.a_acceptor:lay.a_acceptor + NC], b.action[:B, lay.a_acceptor:lay.a_acceptor + NC])
        assert torch.equal(a.action[:B, lay.a_offer_core:lay.a_offer_core + NL], b.action[:B, lay.a_offer_core:lay.a_offer_core + NL])
        assert a.round == b.round == (launch + 1) * T
    assert int(res2[:, :B, lay.r_agent:lay.r_agent + a.N].sum()) > 0     # jobs terminate, agents earn
    a.close(); b.close()
