"""GPU: the ends of the size range -- a batch of ONE environment against a recorded reference trace, and calls over
nothing (zero environments / zero units), which must succeed, launch nothing and leave their outputs alone."""
import numpy as np
import pytest

from helpers import assert_state_equal, golden_names, load_golden

pytestmark = pytest.mark.gpu


def _one(exp, b):
    return {k: (v[b] if isinstance(v, np.ndarray) else v) for k, v in exp.items()}


@pytest.mark.parametrize("name", golden_names()[:4])
def test_single_environment_batch_replays_reference_trace(name):
    """B = 1: one live lane in a 128-environment padding unit (the reference's own shape: one World per process,
    src/world.py:210-254)."""
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    tr, meta = load_golden(name)
    free = meta["mode"].startswith("free")
    env = BatchedSchedulingEnv(1, world_params_from_dom(meta, free), reward=meta["mode"],
                               net_zero_offer_reward=meta.get("netZero", 0.5), auction="external", spawn="u64")
    one = lambda a: np.asarray(a)[None].copy()
    for t in range(tr["done"].shape[0]):
        r = env.step(one(tr["in_offc"][t]), one(tr["in_acc"][t]), one(tr["in_auc"][t]),
                     offer_price=one(tr["in_offp"][t]) if free else None, spawn_u=one(tr["in_spawn_u"][t]))
        e = env.export_state()
        assert (e["flags"] == 0).all(), (name, t)
        assert_state_equal(_one(e, 0), tr, t, prefix=name)
        assert np.array_equal(r["agent"].cpu().numpy()[0], tr["r_agent"][t]), (name, t)
        assert np.array_equal(r["acceptor"].cpu().numpy()[0], tr["r_acceptor"][t]), (name, t)
        assert int(r["n_accepted"][0]) == tr["n_accepted"][t] and int(r["n_terminated"][0]) == tr["n_term"][t]
    env.close()


def test_calls_over_nothing_succeed_and_write_nothing():
    """n_envs == 0 / M == 0: MSCHED_OK without a launch (include/msched.h); the output buffers keep their contents."""
    import torch
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200._lib import MschedError
    dev = torch.device("cuda", 0)
    net = policy.MlpGroup.random(15, 16, 7, 6, dev, seed=1)
    x = torch.zeros((128, 6, 16), dtype=torch.int16, device=dev)
    act = torch.full((6,), -7, dtype=torch.int32, device=dev)
    lp = torch.full((6,), -7.0, dtype=torch.float32, device=dev)
    policy.actor_forward(net, x, 16, 6, 0, seed=1, action=act, logprob=lp)
    # the whole-step launch over zero environments
    core = policy.MlpGroup.random(8, 16, 4, 6, dev, seed=2)
    ga = policy.policy_step_group(net, 6, 0, 16, 0, 1, act, lp)
    go = policy.policy_step_group(core, 6, 96, 8, 6, 2, act, lp)
    policy.policy_step(x, 6 * 16 + 6 * 8, 0, 3, ga, go, None, input_bound=8)
    # returns over zero columns; one time step cannot be normalised (std over one value)
    r0 = torch.zeros((5, 0), dtype=torch.float32, device=dev)
    assert policy.returns(r0, 0.9, True).shape == (5, 0)
    r1 = torch.ones((1, 8), dtype=torch.float32, device=dev)
    assert torch.equal(policy.returns(r1, 0.9, False), r1)
    with pytest.raises(MschedError):
        policy.returns(r1, 0.9, True)
    torch.cuda.synchronize()
    assert int(act.min()) == -7 and float(lp.max()) == -7.0


def test_multi_step_entry_points_at_one_and_zero_steps():
    """msched_step_multi / msched_rollout_hardcoded with T = 1 are one msched_step_observe (+ one hard-coded agents
    launch); T = 0 is refused (n_steps out of range), nothing is launched and the round does not move."""
    import torch
    from marl_scheduling_b200._lib import MschedError
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    dom = dict(N=2, C=3, L=2, prios=[5], lens=[4], probs=[1], fix=[3], mult=2)
    B = 333  # ragged: 11 tiles of 32, the last one with 13 live lanes
    mk = lambda: BatchedSchedulingEnv(B, world_params_from_dom(dom, False), reward="fix", auction="random", spawn="philox", seed=9)
    a, b, c = mk(), mk(), mk()
    lay, dev = a.layout, a.device
    for e in (a, b, c):
        e.observe()
        e.hardcoded_actions(random_ties=True)
    for t in range(5):
        res1 = torch.zeros((1, lay.padded_envs, lay.result_words), dtype=torch.int32, device=dev)
        obs1 = torch.zeros((1, lay.padded_envs, lay.obs_halfs), dtype=torch.int16, device=dev)
        acts = a.action.clone()[None]
        a.step_observe_records(a.action, res1[0], obs=obs1[0])
        a.hardcoded_actions(obs=obs1[0], random_ties=True)
        res2, obs2 = torch.zeros_like(res1), torch.zeros_like(obs1)
        b.step_multi_records(acts, res2, obs2, obs_every=True)         # scripted actions, one step per launch
        b.action.copy_(a.action)                                        # (b's agents are a's)
        res3, obs3 = torch.zeros_like(res1), torch.zeros_like(obs1)
        c.rollout_hardcoded(res3, obs3, obs_every=True, random_ties=True)  # step ; agents, one step per launch
        torch.cuda.synchronize()
        for r, o, e in ((res2, obs2, b), (res3, obs3, c)):
            assert torch.equal(res1[:, :B], r[:, :B]) and torch.equal(obs1[:, :B], o[:, :B]), t
            assert torch.equal(a.state[:B], e.state[:B]) and torch.equal(a.chain[:B], e.chain[:B]), t
        assert torch.equal(a.action[:B], c.action[:B]), t
    assert a.round == b.round == c.round == 5
    empty_a = torch.zeros((0, lay.padded_envs, lay.action_halfs), dtype=torch.int16, device=dev)
    empty_r = torch.zeros((0, lay.padded_envs, lay.result_words), dtype=torch.int32, device=dev)
    with pytest.raises(MschedError):
        b.step_multi_records(empty_a, empty_r, None, obs_every=False)
    with pytest.raises(MschedError):
        c.rollout_hardcoded(empty_r, None, obs_every=False, random_ties=True)
    assert b.round == c.round == 5
    for e in (a, b, c):
        e.close()


@pytest.mark.parametrize("B", [2, 100, 130, 4098])
def test_policy_step_small_and_ragged_batches_tc_equals_simt(B, monkeypatch):
    """msched_policy_step on batches smaller than one tile, straddling two, and with a 2-environment last tile: the
    tensor-core kernel (slots without a tile, the environment-pair Philox call shared between two tiles of a slot)
    against the fp32 SIMT kernel with the same draws -- experience rows identical, the same actions up to draws that
    sit on a CDF step."""
    import torch
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    dom = dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1, 2, 3])
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL, P = N * L, max(dom["prios"])
    env = BatchedSchedulingEnv(B, world_params_from_dom(dom, True), reward="free_comm", auction="random", spawn="philox", seed=4)
    dev, lay = env.device, env.layout
    g = torch.Generator(device=dev).manual_seed(2)
    for t in range(10):
        env.acceptor_actions.random_(0, 2, generator=g)
        env.offer_core_actions.random_(0, C + 1, generator=g)
        env.offer_price_actions.random_(0, P + 1, generator=g)
        env.step_observe_records()
    Ua, Uo = N * C, NL
    ga = policy.MlpGroup.random(3 + 2 * NL, 16, NL + 1, Ua, dev, seed=31)
    go = policy.MlpGroup.random(2 * C + 2, 16, C + 1, Uo, dev, seed=32)
    gp = policy.MlpGroup.random(4, 16, P + 1, Uo, dev, seed=33)
    res = {}
    for impl in ("tc", "simt"):
        monkeypatch.setenv("MSCHED_POLICY_STEP_IMPL", impl)
        out = [(torch.full((B, n), -1, dtype=torch.int32, device=dev), torch.zeros((B, n), device=dev)) for n in (Ua, Uo, Uo)]
        xs = [torch.zeros((B, Ua, lay.o_acc_row), dtype=torch.int16, device=dev), torch.zeros((B, Uo, lay.o_off_row), dtype=torch.int16, device=dev),
              torch.zeros((B, Uo, 4), dtype=torch.int16, device=dev)]
        A_ = policy.policy_step_group(ga, Ua, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor, 7, *out[0], x_used=xs[0])
        O_ = policy.policy_step_group(go, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_core, 8, *out[1], x_used=xs[1])
        P_ = policy.policy_step_group(gp, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_price, 9, *out[2], x_used=xs[2])
        policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, C, A_, O_, P_, env_offset=6, step=11, input_bound=16)
        torch.cuda.synchronize()
        res[impl] = (out, xs)
    (to, tx), (so, sx) = res["tc"], res["simt"]
    for k in range(2):
        assert torch.equal(tx[k], sx[k])
        assert int(to[k][0].min()) >= 0                                  # every row was written
        diff = int((to[k][0] != so[k][0]).sum())
        assert diff <= max(1, to[k][0].numel() // 1000), (k, diff)
        same = to[k][0] == so[k][0]
        assert float((to[k][1] - so[k][1])[same].abs().max()) < 1e-4
    same_core = (to[1][0] == so[1][0])
    assert torch.equal(tx[2][same_core], sx[2][same_core])               # the price chooser's inputs follow the sampled core
    env.close()


def test_two_million_environments_last_window_equals_offset_batch():
    """32 x BASELINE's batch (2,097,152 environments, 1.5 GB of records): the last 128 environments walk through the
    states, results and observations of a 128-environment batch created at env_offset = B - 128 (64-bit offsets, draws
    keyed on the global index), and spawned = terminated + present holds for every environment."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "big_batch_check", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "big_batch_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check(1 << 21, 20) > (1 << 21)


@pytest.mark.parametrize("base", [(1 << 32) - 40, (1 << 33) + 6, (1 << 40) + 2])
def test_global_environment_index_beyond_32_bits_matches_oracle(base):
    """env_offset around and beyond 2^32 (a shard deep inside a very large job): the Philox counters carry the global
    environment index in two 32-bit words -- the batch straddling the 2^32 boundary included -- and the step walks
    through the oracle's states, rewards and auction winners."""
    import torch
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    from helpers import STATE_KEYS
    from oracle import oracle as O
    dom = dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1])
    B, seed = 96, 12
    env = BatchedSchedulingEnv(B, world_params_from_dom(dom, True), reward="free_comm", auction="random", spawn="philox",
                               seed=seed, env_offset=base)
    orc = O.Oracle(B, dom, "free_comm", tie_mode=O.TIE_PHILOX, seed=seed, env_offset=base)
    rng = np.random.default_rng(2)
    N, C, L = 2, 3, 3
    for t in range(25):
        acc = np.where(rng.random((B, N, C)) < 0.6, rng.integers(0, 2, (B, N, C)), rng.integers(0, N * L + 1, (B, N, C)))
        offc, offp = rng.integers(0, C + 1, (B, N, L)), rng.integers(0, 9, (B, N, L))
        r = env.step(offc, acc, None, offer_price=offp)
        orc.step(offc, acc, None, offp=offp)
        assert np.array_equal(r["auctioneer_idx"].cpu().numpy(), orc.auc_out), t
        assert np.array_equal(r["agent"].cpu().numpy(), orc.r_agent), t
        assert np.array_equal(r["acceptor"].cpu().numpy(), orc.r_acceptor), t
        assert int(r["flags"].max()) == 0
    e = env.export_state()
    for b in (0, 39, 40, 41, B - 1):
        ob = orc.export(b)
        for k in STATE_KEYS:
            assert np.array_equal(np.asarray(e[k][b]), np.asarray(ob[k])), (b, k)
    env.close()


def test_policy_step_tensor_core_inputs_at_the_bound():
    """The tensor-core kernel takes observations up to |x| = 511 as exact fp16 operands (include/msched.h
    input_bound): an observation record filled with values over the whole range [-511, 511], the extremes included,
    gives the fp32 SIMT kernel's probabilities and the same experience rows."""
    import os
    import torch
    from marl_scheduling_b200 import policy
    dev = torch.device("cuda", 0)
    B, C, N, L = 300, 3, 2, 3
    NL, Ua, Uo = N * L, N * C, N * L
    row_a, row_o = 16, 8                      # the config-3 layout: acceptor rows 15 values behind one pad, offer rows 8
    stride = Ua * row_a + Uo * row_o
    g = torch.Generator(device=dev).manual_seed(5)
    obs = torch.randint(-511, 512, (B, stride), dtype=torch.int16, device=dev, generator=g)
    obs[0, :] = 511
    obs[1, :] = -511
    ga = policy.MlpGroup.random(15, 16, 7, Ua, dev, seed=1)
    go = policy.MlpGroup.random(8, 16, 4, Uo, dev, seed=2)
    for grp, n_in in ((ga, 15), (go, 8)):     # small first-layer weights keep the Tanh inputs out of saturation
        grp.weights[:, :16 * n_in].mul_(0.004)
    res = {}
    old = os.environ.get("MSCHED_POLICY_STEP_IMPL")
    try:
        for impl in ("tc", "simt"):
            os.environ["MSCHED_POLICY_STEP_IMPL"] = impl
            pa = torch.zeros((B, Ua, 7), device=dev)
            po = torch.zeros((B, Uo, 4), device=dev)
            xa = torch.zeros((B, Ua, row_a), dtype=torch.int16, device=dev)
            xo = torch.zeros((B, Uo, row_o), dtype=torch.int16, device=dev)
            A_ = policy.policy_step_group(ga, Ua, 1, row_a, 0, 1, x_used=xa, probs=pa)
            O_ = policy.policy_step_group(go, Uo, Ua * row_a, row_o, Ua, 2, x_used=xo, probs=po)
            policy.policy_step(obs, stride, B, C, A_, O_, None, step=1, input_bound=511)
            torch.cuda.synchronize()
            res[impl] = (pa, po, xa, xo)
    finally:
        os.environ.pop("MSCHED_POLICY_STEP_IMPL", None)
        if old is not None:
            os.environ["MSCHED_POLICY_STEP_IMPL"] = old
    t, s = res["tc"], res["simt"]
    assert torch.equal(t[2], s[2]) and torch.equal(t[3], s[3])
    assert torch.equal(t[3][:2].reshape(2, -1), obs[:2, Ua * row_a:].reshape(2, -1))     # the extremes came through
    for k in (0, 1):
        assert float((t[k] - s[k]).abs().max()) < 2e-5, k
        assert float((t[k].sum(-1) - 1).abs().max()) < 1e-5
    assert float(t[0].std()) > 1e-3                                   # not a saturated / uniform comparison


@pytest.mark.parametrize("B", [None, 1, 3])
def test_drop_in_loop_at_the_reference_batch_of_one(B):
    """The reference's own shape: ONE world per process (numberOfEnvironments absent = 1).  The free-price PPO loop of
    src/trainPPOExperiment4-2.py -- getActionForAllAgents, step, saveRewards, updateAgents -- through the drop-in
    classes with 1 (and 3) environments: the one-launch policy step on a 1-environment batch, a gradient over T x 1
    samples per net, returns over a handful of columns."""
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    wp = dict(freePrices=True, fixPricesList=[1], numberOfAgents=2, numberOfCores=3, collectionLength=3,
              possibleJobPriorities=[2, 4, 8], possibleJobLengths=[5, 5, 5], probabilities=[1 / 3] * 3,
              newJobsPerRoundPerAgent=1, rewardMultiplier=1, episodeLength=10, maxVisibleOffers=4)
    rl = dict(netZeroOfferReward=0.5, LR_ACTOR=0.003, LR_CRITIC=0.01, OFFER_GAMMA=0.5, ACCEPTOR_GAMMA=0.8733,
              EPS_CLIP=0.2, RAW_K_EPOCHS=2, ACCEPTOR_K_EPOCHS=2, OFFER_K_EPOCHS=2, CENTRALISATION_SAMPLE=2)
    if B is not None:
        wp["numberOfEnvironments"] = B
    n = 1 if B is None else B
    world = World(wp)
    env = SE.PPODividedFreePriceEnv(world, rl, True)
    accO, offO, aucO = env.reset()
    total = 0.0
    for t in range(20):
        acceptorActions, (core, price) = env.getActionForAllAgents(accO, offO)
        assert acceptorActions.shape == (n, 2, 3) and core.shape == (n, 2, 3) and price.shape == (n, 2, 3)
        assert bool(((price == -5) == (core == 0)).all())
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step((core, price), acceptorActions, None)
        assert bool(done) == ((t + 1) % 10 == 0) if not torch.is_tensor(done) else bool(done.all()) == ((t + 1) % 10 == 0)
        env.saveRewards(offR, accR, agR)
        total += float(torch.as_tensor(agR).sum())
    before = env.agents.acceptor.actor.detach().clone()
    env.updateAgents()
    after = env.agents.acceptor.actor.detach()
    assert torch.isfinite(after).all() and torch.isfinite(env.agents.price.actor).all()
    assert not torch.equal(before, after)
    env.close()


def test_dqn_drop_in_loop_with_one_world():
    """The DQN loop of src/trainDQN.py:171-186 with ONE world: optimize_model returns nothing until the replay memory
    holds a batch (src/DQNmodules.py:98-99), then finite losses; one transition per step and unit is stored."""
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    wp = dict(freePrices=False, fixPricesList=[2, 7], numberOfAgents=2, numberOfCores=2, collectionLength=3,
              possibleJobPriorities=[3, 10], possibleJobLengths=[6, 3], probabilities=[0.8, 0.2],
              newJobsPerRoundPerAgent=1, rewardMultiplier=1, episodeLength=10, maxVisibleOffers=4, seed=2)
    params = dict(netZeroOfferReward=0.5, RUN_END=0.05, RUN_START=0.9, RUN_DECAY=200, BATCH_SIZE=8,
                  OFFER_GAMMA=0.5, ACCEPTOR_GAMMA=0.87, REPLAY_MEMORY_SIZE=64)
    world = World(wp)
    env = SE.DQNDividedFixedPricesEnv(world, params)
    accO, offO, aucO = env.reset()
    got = []
    for t in range(20):
        oldA, oldO = accO, offO
        acceptorActions, offerActions = env.getActionForAllAgents(accO, offO)
        assert acceptorActions.shape[0] == 1 and offerActions.shape[0] == 1
        aa = world.auctioneer.getAuctioneerAction(aucO)
        actA, actO = acceptorActions.clone(), offerActions.clone()
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(offerActions, acceptorActions, aa)
        l1 = env.updateOfferMemoriesAndOptimize(oldO, actO, offO, offR[..., 0])
        l2 = env.updateAcceptorMemoriesAndOptimize(oldA, actA, accO, accR[..., 0])
        got.append(l1 is not None and l2 is not None)
        if got[-1]:
            assert bool(torch.isfinite(l1).all()) and bool(torch.isfinite(l2).all())
    assert got[:7] == [False] * 7 and all(got[7:])      # a batch of 8 exists from the 8th transition on
    assert len(env.agents.acceptor.memory) == 20
    env.close()
