"""GPU: batched DQN interface: epsilon-greedy Q-forward kernel vs a float64
evaluation of the same nets, replay memory and one optimize_model step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_dqn_select_matches_reference_semantics():
    import torch
    from marl_scheduling_b200.dqn import BatchedDQN
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(4)
    n_envs, units, nin, A = 517, 5, 27, 13
    dqn = BatchedDQN(nin, A, units, run_start=0.9, run_end=0.05, run_decay=200.0, gamma=0.9, memory_size=64,
                     device=dev, seed=3)
    xs = rng.integers(-2, 12, (n_envs, units * nin)).astype(np.int16)
    u = rng.random((n_envs * units, 2)).astype(np.float32)
    rnd = 150
    act, q = dqn.selectAction(torch.as_tensor(xs).to(dev), nin, 0, n_envs, rnd, seed=1, u=u, want_q=True)
    act, q = act.cpu().numpy(), q.cpu().numpy()
    eps = 0.05 + (0.9 - 0.05) * np.exp(-rnd / 200.0)
    w = dqn.policy.detach().cpu().numpy().astype(np.float64)
    uu = u.reshape(n_envs, units, 2)
    for n in range(units):
        o, ws = 0, []
        for sz in (16 * nin, 16, A * 16, A):
            ws.append(w[n, o:o + sz]); o += sz
        x = xs[:, n * nin:(n + 1) * nin].astype(np.float64)
        qr = np.tanh(x @ ws[0].reshape(16, nin).T + ws[1]) @ ws[2].reshape(A, 16).T + ws[3]
        np.testing.assert_allclose(q[:, n], qr, rtol=2e-5, atol=2e-6)
        greedy = qr.argmax(1)
        explore = ~(uu[:, n, 0] > np.float32(eps))
        rand_a = np.minimum((uu[:, n, 1] * np.float32(A)).astype(np.int64), A - 1)
        srt = np.sort(qr, 1)
        sure = (srt[:, -1] - srt[:, -2]) > 1e-4
        exp = np.where(explore, rand_a, greedy)
        assert np.array_equal(act[:, n][sure | explore], exp[sure | explore])
    # random policy / epsilon 1: uniform actions
    a2 = dqn.selectAction(torch.as_tensor(xs).to(dev), nin, 0, n_envs, rnd, seed=5, random_policy=True).cpu().numpy()
    freq = np.bincount(a2.reshape(-1), minlength=A) / a2.size
    assert np.abs(freq - 1 / A).max() < 0.03


def test_dqn_env_loop_learns_something():
    import torch
    from marl_scheduling_b200 import SchedulingEnvironment as SE
    from marl_scheduling_b200.world import World
    B = 64
    wp = dict(freePrices=False, fixPricesList=[2, 7], numberOfAgents=2, numberOfCores=2, collectionLength=3,
              possibleJobPriorities=[3, 10], possibleJobLengths=[6, 3], probabilities=[0.8, 0.2],
              newJobsPerRoundPerAgent=1, rewardMultiplier=1, episodeLength=10, maxVisibleOffers=4,
              numberOfEnvironments=B, seed=2)
    params = dict(netZeroOfferReward=0.5, RUN_END=0.05, RUN_START=0.9, RUN_DECAY=200, BATCH_SIZE=32,
                  OFFER_GAMMA=0.5, ACCEPTOR_GAMMA=0.87, REPLAY_MEMORY_SIZE=4096)
    world = World(wp)
    env = SE.DQNDividedFixedPricesEnv(world, params)
    accO, offO, aucO = env.reset()
    before = env.agents.acceptor.policy.detach().clone()
    for t in range(12):
        oldA, oldO = accO.clone(), offO.clone()
        acceptorActions, offerActions = env.getActionForAllAgents(accO, offO)
        assert int(acceptorActions.max()) <= 6 and int(offerActions.max()) <= 2
        aa = world.auctioneer.getAuctioneerAction(aucO)
        actA, actO = acceptorActions.clone(), offerActions.clone()
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(offerActions, acceptorActions, aa)
        l1 = env.updateOfferMemoriesAndOptimize(oldO, actO, offO, offR[..., 0])
        l2 = env.updateAcceptorMemoriesAndOptimize(oldA, actA, accO, accR[..., 0])
        assert l1 is not None and l2 is not None and np.isfinite(l1) and np.isfinite(l2)
    env.agents.updateTargetNets()
    assert not torch.equal(before, env.agents.acceptor.policy.detach())
    assert torch.equal(env.agents.acceptor.target, env.agents.acceptor.policy.detach())
    assert len(env.agents.acceptor.memory) == 12 * B
    env.close()
