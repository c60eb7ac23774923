"""GPU: the drop-in classes (World / SchedulingEnv / Auctioneer / PPO envs) keep the reference's
call shapes and agree with the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WP = dict(freePrices=False, fixPricesList=[2, 7], numberOfAgents=4, numberOfCores=4, collectionLength=3,
          possibleJobPriorities=[3, 10], possibleJobLengths=[6, 3], probabilities=[0.8, 0.2],
          newJobsPerRoundPerAgent=1, rewardMultiplier=1, episodeLength=10, maxVisibleOffers=4)
RL = dict(netZeroOfferReward=0.5, LR_ACTOR=0.003, LR_CRITIC=0.01, OFFER_GAMMA=0.5, ACCEPTOR_GAMMA=0.8733,
          EPS_CLIP=0.2, RAW_K_EPOCHS=2, ACCEPTOR_K_EPOCHS=2, OFFER_K_EPOCHS=2, CENTRALISATION_SAMPLE=2)
DOM = dict(N=4, C=4, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7], episodeLength=10)


def test_hardcoded_env_loop_matches_oracle():
    """trainHC-shaped loop: auctioneer action from world.auctioneer, 9-tuple from env.step."""
    import torch
    from marl_scheduling_b200.SchedulingEnvironment import HardcodedFixPriceEnvironment
    from marl_scheduling_b200.world import World
    from oracle import oracle as O
    B = 300
    world = World(dict(WP, numberOfEnvironments=B, seed=5, chainCapacity=32))
    env = HardcodedFixPriceEnvironment(world, dict(netZeroOfferReward=0.5))
    orc = O.Oracle(B, DOM, "fix", tie_mode=O.TIE_PHILOX, seed=5)
    accO, offO, aucO = env.reset()
    assert accO.shape == (B, 4, 4, 27) and offO.shape == (B, 4, 3, 10) and aucO.shape == (B, 4, 27)
    rng = np.random.default_rng(0)
    for t in range(25):
        offc = rng.integers(0, 5, (B, 4, 3))
        acc = rng.integers(0, 3, (B, 4, 4))
        aa = world.auctioneer.getAuctioneerAction(aucO)
        out = env.step(torch.as_tensor(offc), torch.as_tensor(acc), aa)
        assert len(out) == 9
        accO, offO, aucO, offR, accR, aucR, agR, (qm, qn), done = out
        orc.step(offc, acc, None)
        assert np.array_equal(aa.cpu().numpy(), orc.auc_out), t
        assert offR.shape == (B, 4, 3, 1) and accR.shape == (B, 4, 4, 1)
        assert np.array_equal(offR[..., 0].cpu().numpy().astype(np.float64), orc.r_offer)
        assert np.array_equal(accR[..., 0].cpu().numpy(), orc.r_acceptor)
        assert np.array_equal(aucR.cpu().numpy(), orc.r_auctioneer)
        assert np.array_equal(agR.cpu().numpy(), orc.r_agent)
        assert done == bool(orc.done[0]) == ((t + 1) % 10 == 0)
        assert world.round == t + 1
        o = orc.observe(7)
        assert np.array_equal(accO[7].cpu().numpy(), o["obs_acc"])
        assert np.array_equal(aucO[7].cpu().numpy(), o["obs_auc"])
        qn_ = qn.cpu().numpy()
        assert np.array_equal(qn_, orc.quality_cnt)
        assert np.isnan(qm.cpu().numpy()[qn_ == 0]).all()
    env.close()


@pytest.mark.parametrize("cls", ["divided", "global", "local"])
def test_ppo_fixed_price_env_rollout_and_update(cls):
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    B = 64
    world = World(dict(WP, numberOfEnvironments=B, seed=1))
    env = {"divided": SE.PPODividedFixedPriceEnv, "global": SE.GloballySharedParamsDividedFixedPriceEnv,
           "local": SE.LocallySharedParamsDividedFixedPriceEnv}[cls](world, RL)
    accO, offO, aucO = env.reset()
    for t in range(12):
        acceptorActions, offerActions = env.getActionForAllAgents(accO, offO)
        assert acceptorActions.shape == (B, 4, 4) and offerActions.shape == (B, 4, 3)
        assert int(acceptorActions.max()) <= 12 and int(offerActions.max()) <= 4
        aa = world.auctioneer.getAuctioneerAction(aucO)
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(offerActions, acceptorActions, aa)
        env.saveRewards(offR, accR, agR)
    ag = env.agents
    # log-probs stored by the CUDA actor kernel == torch forward of policy_old on the stored obs
    ppo = ag.acceptor
    X = torch.stack(ppo.buf_x).float()
    T = X.shape[0]
    nets = torch.tensor([(u // ppo.unit_div) % ppo.n_nets for u in range(ppo.units)], device=X.device)
    flat = ppo.policy_old.weights[nets]
    logits = ppo._forward(flat, X.permute(2, 0, 1, 3).reshape(ppo.units, T * B, -1), ppo.A)
    lp = torch.log_softmax(logits, -1).gather(-1, torch.stack(ppo.buf_a).long().permute(2, 0, 1).reshape(
        ppo.units, T * B, 1)).squeeze(-1)
    lp_k = torch.stack(ppo.buf_lp).permute(2, 0, 1).reshape(ppo.units, T * B)
    assert torch.allclose(lp, lp_k, rtol=1e-4, atol=1e-4)
    before = ppo.actor.detach().clone()
    env.updateAgents()
    assert not torch.equal(before, ppo.actor.detach())
    assert torch.equal(ppo.policy_old.weights, ppo.actor.detach())
    assert ppo.buf_r == [] and torch.isfinite(ppo.actor).all()
    env.close()


def test_ppo_free_price_env_rollout_and_update():
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    B = 48
    wp = dict(WP, freePrices=True, numberOfAgents=2, numberOfCores=3, collectionLength=3,
              possibleJobPriorities=[2, 4, 8], possibleJobLengths=[5, 5, 5], probabilities=[1 / 3] * 3,
              numberOfEnvironments=B)
    world = World(wp)
    env = SE.PPODividedFreePriceEnv(world, RL, True)
    accO, offO, aucO = env.reset()
    for t in range(8):
        acceptorActions, offerActions = env.getActionForAllAgents(accO, offO)
        core, price = offerActions
        assert core.shape == (B, 2, 3) and price.shape == (B, 2, 3)
        assert bool(((price == -5) == (core == 0)).all())  # quirk Q1
        assert int(price.max()) <= 8
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(offerActions, acceptorActions, None)
        assert isinstance(offR, tuple) and offR[0].shape == (B, 2, 3, 1) and offR[1].shape == (B, 2, 3, 1)
        env.saveRewards(offR, accR, agR)
    env.updateAgents()
    assert torch.isfinite(env.agents.price.actor).all()
    env.close()


def test_aggregated_env_views_match_oracle():
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    from oracle import oracle as O
    B = 40
    dom = dict(N=2, C=2, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7])
    wp = dict(WP, numberOfAgents=2, numberOfCores=2, numberOfEnvironments=B, seed=3)
    world = World(wp)
    env = SE.PPOFullyAggregatedFixPriceEnv(world, RL, agents=False)
    orc = O.Oracle(B, dom, "agg", tie_mode=O.TIE_PHILOX, seed=3)
    rng = np.random.default_rng(1)
    for t in range(15):
        offc, acc = rng.integers(0, 3, (B, 2, 3)), rng.integers(0, 3, (B, 2, 2))
        out = env.step(torch.as_tensor(offc), torch.as_tensor(acc), None)
        orc.step(offc, acc, None)
        assert out[3].shape == (B, 2, 1) and out[4].shape == (B, 2, 1)
        assert np.array_equal(out[4][..., 0].cpu().numpy(), orc.r_acceptor[..., 0])
        assert np.array_equal(out[3][..., 0].cpu().numpy().astype(np.float64), orc.r_offer[..., 0])
    accA, offA = env.aggregatedObservations()
    full = env.fullyAggregatedObservations()
    o = orc.observe(5)
    assert np.array_equal(accA[5].cpu().numpy(), o["obs_acc"].reshape(2, -1).astype(np.float32))
    cores = o["obs_off"][:, 0, :4]
    slots = o["obs_off"][:, :, 4:].reshape(2, 6)
    assert np.array_equal(offA[5].cpu().numpy(), np.concatenate([cores, slots], 1))
    assert full.shape == (B, 2, 10 + 2 * 15)
    env.close()


@pytest.mark.parametrize("kind", ["semi", "fully"])
def test_aggregated_ppo_env_rollout_and_update(kind):
    """Semi- / fully-aggregated agents (src/Agent.py:359-492) on the README 2-agent domain: the
    aggregated heads (49 / 27 / 1,323 actions) run in the tensor-core actor kernel, the action
    number decodes to per-core / per-slot actions, the log-probs match a torch forward."""
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    B = 40
    wp = dict(WP, numberOfAgents=2, numberOfCores=2, numberOfEnvironments=B, seed=4)
    world = World(wp)
    env = (SE.PPOAggregatedFixPriceEnv if kind == "semi" else SE.PPOFullyAggregatedFixPriceEnv)(world, RL)
    accO, offO, aucO = env.reset()
    assert accO.shape == (B, 2, 2 * 15) and offO.shape == (B, 2, 4 + 6) and aucO.shape == (B, 2, 15)
    for t in range(10):
        acceptorActions, offerActions = env.getActionForAllAgents(accO, offO)
        assert acceptorActions.shape == (B, 2, 2) and offerActions.shape == (B, 2, 3)
        assert int(acceptorActions.max()) <= 6 and int(offerActions.max()) <= 2
        aa = world.auctioneer.getAuctioneerAction(aucO)
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(offerActions, acceptorActions, aa)
        assert offR.shape == (B, 2, 1) and accR.shape == (B, 2, 1)
        env.saveRewards(offR, accR, agR)
        assert int(q[1].max()) >= 0
    ppo = env.agents.acceptor if kind == "semi" else env.agents.unit
    X = torch.stack(ppo.buf_x).float()
    T = X.shape[0]
    logits = ppo._forward(ppo.policy_old.weights, X.permute(2, 0, 1, 3).reshape(ppo.units, T * B, -1), ppo.A)
    lp = torch.log_softmax(logits, -1).gather(-1, torch.stack(ppo.buf_a).long().permute(2, 0, 1).reshape(
        ppo.units, T * B, 1)).squeeze(-1)
    lp_k = torch.stack(ppo.buf_lp).permute(2, 0, 1).reshape(ppo.units, T * B)
    assert torch.allclose(lp, lp_k, rtol=1e-4, atol=1e-4)
    before = ppo.actor.detach().clone()
    env.updateAgents()
    assert not torch.equal(before, ppo.actor.detach()) and torch.isfinite(ppo.actor).all()
    env.close()


def test_ppo_checkpoint_roundtrip(tmp_path):
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.results import load_checkpoint, save_checkpoint
    from marl_scheduling_b200.world import World
    B = 16
    world = World(dict(WP, numberOfEnvironments=B, seed=1))
    env = SE.PPODividedFixedPriceEnv(world, RL)
    accO, offO, aucO = env.reset()
    for t in range(6):
        a, o = env.getActionForAllAgents(accO, offO)
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(
            o, a, world.auctioneer.getAuctioneerAction(aucO))
        env.saveRewards(offR, accR, agR)
    env.updateAgents()
    path = str(tmp_path / "ckpt.pt")
    save_checkpoint(path, world)
    saved = world.agents.acceptor.actor.detach().clone()
    with torch.no_grad():
        world.agents.acceptor.actor.add_(1.0)
        world.agents.acceptor.policy_old.weights.zero_()
    load_checkpoint(path, world)
    assert torch.equal(world.agents.acceptor.actor.detach(), saved)
    assert torch.equal(world.agents.acceptor.policy_old.weights, saved)
    env.close()


@pytest.mark.parametrize("arch", ["divided", "free", "semi", "fully"])
def test_save_rewards_routes_like_the_reference(arch):
    """Row a13: what env.saveRewards leaves in every unit's reward buffer, per architecture, against the oracle's
    rewards for the same actions routed by the reference's rules -- divided: unit (i,k) gets [i][k][0]
    (src/Agent.py:531-536); free prices: core chooser / price chooser / acceptor planes (:610-619); semi-aggregated:
    the acceptor unit gets agentReward[i], the offer unit offerRewards[i][0] (:388-390,
    src/SchedulingEnvironment.py:223-225); fully aggregated: agentReward[i] + offerRewards[i][0] (:490-492,
    src/SchedulingEnvironment.py:243-247)."""
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    from oracle import oracle as O
    B, T = 48, 14
    if arch == "free":
        dom = dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1], episodeLength=10)
        wp = dict(WP, freePrices=True, numberOfAgents=2, numberOfCores=3, collectionLength=3,
                  possibleJobPriorities=[2, 4, 8], possibleJobLengths=[5, 5, 5], probabilities=[1 / 3] * 3)
        mode = "free_comm"
    elif arch == "divided":
        dom, wp, mode = DOM, WP, "fix"
    else:
        dom = dict(N=2, C=2, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7], episodeLength=10)
        wp = dict(WP, numberOfAgents=2, numberOfCores=2)
        mode = "agg"
    world = World(dict(wp, numberOfEnvironments=B, seed=6))
    env = {"divided": lambda: SE.PPODividedFixedPriceEnv(world, RL), "free": lambda: SE.PPODividedFreePriceEnv(world, RL, True),
           "semi": lambda: SE.PPOAggregatedFixPriceEnv(world, RL), "fully": lambda: SE.PPOFullyAggregatedFixPriceEnv(world, RL)}[arch]()
    orc = O.Oracle(B, dom, mode, tie_mode=O.TIE_PHILOX, seed=6)
    accO, offO, aucO = env.reset()
    exp = {}
    for t in range(T):
        acceptorActions, offerActions = env.getActionForAllAgents(accO, offO)
        if arch == "free":
            offc, offp = (x.cpu().numpy() for x in offerActions)
        else:
            offc, offp = offerActions.cpu().numpy(), None
        acc = acceptorActions.cpu().numpy()
        aa = world.auctioneer.getAuctioneerAction(aucO)
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(offerActions, acceptorActions, aa)
        env.saveRewards(offR, accR, agR)
        orc.step(offc, acc, None, offp=offp)
        if arch == "divided":
            cur = dict(acceptor=orc.r_acceptor.reshape(B, -1), offer=orc.r_offer.reshape(B, -1))
        elif arch == "free":
            cur = dict(acceptor=orc.r_acceptor.reshape(B, -1), core=orc.r_offer.reshape(B, -1), price=orc.r_price.reshape(B, -1))
        elif arch == "semi":
            cur = dict(acceptor=orc.r_agent, offer=orc.r_offer[..., 0])
        else:
            cur = dict(unit=orc.r_agent + orc.r_offer[..., 0])
        for k, v in cur.items():
            exp.setdefault(k, []).append(np.asarray(v, np.float64).copy())
    n_nonzero = 0
    for name, rows in exp.items():
        ppo = getattr(env.agents, name)
        got = torch.stack(ppo.buf_r).cpu().numpy().astype(np.float64)
        want = np.stack(rows)
        assert got.shape == want.shape == (T, B, ppo.units), (name, got.shape, want.shape)
        assert np.array_equal(got, want), name
        n_nonzero += int((want != 0).sum())
    assert n_nonzero > 20
    env.close()
