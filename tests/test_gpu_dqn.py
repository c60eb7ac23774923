"""GPU: batched DQN interface: epsilon-greedy Q-forward kernel and the optimize_model gradient
kernel against fixtures recorded from the UNMODIFIED src/DQNmodules.py (tests/golden/dqn.npz,
oracle/gen_dqn_golden.py) and against the oracle's float64 restatement."""
import os

import numpy as np
import pytest

from helpers import GOLDEN

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


@pytest.mark.parametrize("tag", ["acc_cfg2", "off_cfg2", "acc_cfg3"])
def test_dqn_select_and_optimize_match_reference_fixtures(tag):
    """DQNEntity.forward / selectAction with the recorded Python draws, then three optimize_model steps on the
    recorded batches: Q-values, actions and the weights after every step equal the reference's."""
    import torch
    from marl_scheduling_b200.dqn import BatchedDQN
    z = np.load(os.path.join(GOLDEN, "dqn.npz"))
    dev = torch.device("cuda", 0)
    nin, A = int(z[tag + ".n_in"]), int(z[tag + ".n_actions"])
    dqn = BatchedDQN(nin, A, 1, run_start=0.9, run_end=0.05, run_decay=200.0, gamma=float(z[tag + ".gamma"]),
                     memory_size=16, device=dev, seed=0)
    w0 = torch.as_tensor(z[tag + ".w0"]).to(dev).view(1, -1)
    dqn.policy.data.copy_(w0)
    dqn.target.copy_(w0)
    x = torch.as_tensor(z[tag + ".x"]).to(dev)
    M = x.shape[0]
    u = np.stack([z[tag + ".sample"], (z[tag + ".randrange"] + 0.5) / A], 1).astype(np.float32)
    assert abs(dqn.epsilon(150) - float(z[tag + ".eps"])) < 1e-12
    act, q = dqn.selectAction(x, nin, 0, M, 150, seed=1, u=u, want_q=True)
    np.testing.assert_allclose(q.cpu().numpy()[:, 0], z[tag + ".q"], rtol=2e-5, atol=2e-6)
    srt = np.sort(z[tag + ".q"], 1)
    sure = ((srt[:, -1] - srt[:, -2]) > 1e-4) | ~(z[tag + ".sample"] > float(z[tag + ".eps"]))
    assert sure.mean() > 0.95 and np.array_equal(act.cpu().numpy()[:, 0][sure], z[tag + ".action"][sure])
    S, S2 = (torch.as_tensor(z[tag + k]).to(dev).unsqueeze(1) for k in (".S", ".S2"))
    Aa, Rr = (torch.as_tensor(z[tag + k]).to(dev).unsqueeze(1) for k in (".A", ".R"))
    for k, idx in enumerate(z[tag + ".idx"]):
        i = torch.as_tensor(idx.astype(np.int64)).to(dev)
        dqn.optimize_batch(S[i], Aa[i], S2[i], Rr[i].float())
        np.testing.assert_allclose(dqn.policy.detach().cpu().numpy()[0], z[tag + ".w_after"][k], rtol=2e-4, atol=2e-6)
    assert torch.equal(dqn.target, w0)


def test_dqn_gradient_kernel_matches_oracle_for_many_units():
    """Several nets per launch, a batch that is not a multiple of the 128-transition chunk, large TD errors (the
    linear branch of the Huber loss) and the clamp of the gradient to [-1, 1]; bit-reproducible."""
    import torch
    from marl_scheduling_b200.dqn import BatchedDQN
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(9)
    units, nin, A, batch = 5, 10, 5, 300
    dqn = BatchedDQN(nin, A, units, run_start=0.9, run_end=0.05, run_decay=200.0, gamma=0.5, memory_size=16, device=dev, seed=4)
    dqn.policy.data[:, :16 * nin].mul_(0.01)   # keep tanh out of saturation so that large inputs give large gradients
    dqn.target.mul_(0.5)
    S = rng.integers(-20, 120, (batch, units, nin)).astype(np.int16)   # large inputs: some gradients exceed the clamp
    S2 = rng.integers(-2, 12, (batch, units, nin)).astype(np.int16)
    Aa = rng.integers(0, A, (batch, units)).astype(np.int32)
    Rr = rng.integers(-8, 30, (batch, units)).astype(np.float32)
    w, wt = dqn.policy.detach().cpu().numpy().astype(np.float64), dqn.target.cpu().numpy().astype(np.float64)
    args = [torch.as_tensor(t).to(dev) for t in (S, Aa, S2, Rr)]
    loss = dqn.optimize_batch(*args).cpu().numpy()
    g1 = dqn._grad.clone()
    clamped = 0
    for n in range(units):
        g, l = O.dqn_grad(w[n], wt[n], S[:, n], Aa[:, n], S2[:, n], Rr[:, n], 0.5, A)
        np.testing.assert_allclose(g1[n].cpu().numpy(), g, rtol=2e-4, atol=2e-6)
        assert abs(loss[n] - l) < 1e-4 * max(1.0, abs(l))
        clamped += int((np.abs(g) == 1.0).sum())
    assert clamped > 0  # the clamp is exercised
    dqn.policy.data.copy_(torch.as_tensor(w).float().to(dev))
    dqn.optimize_batch(*args)
    assert torch.equal(g1, dqn._grad)


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
    changed = 0
    for t in range(12):
        oldA, oldO = accO, offO   # the reference's `old = new` idiom: step() returns views that stay valid one more step
        acceptorActions, offerActions = env.getActionForAllAgents(accO, offO)
        assert int(acceptorActions.max()) <= 6 and int(offerActions.max()) <= 2
        aa = world.auctioneer.getAuctioneerAction(aucO)
        actA, actO = acceptorActions.clone(), offerActions.clone()
        accO, offO, aucO, offR, accR, aucR, agR, q, done = env.step(offerActions, acceptorActions, aa)
        l1 = env.updateOfferMemoriesAndOptimize(oldO, actO, offO, offR[..., 0])
        l2 = env.updateAcceptorMemoriesAndOptimize(oldA, actA, accO, accR[..., 0])
        assert l1 is not None and l2 is not None
        assert bool(torch.isfinite(l1).all()) and bool(torch.isfinite(l2).all())
        assert oldA.data_ptr() != accO.data_ptr() and oldO.data_ptr() != offO.data_ptr()   # no aliasing of old and new
        changed += int(not torch.equal(oldA, accO)) + int(not torch.equal(oldO, offO))
    assert changed > 6   # (the first acceptor observations are all "empty core, no offers": equal but distinct buffers)
    env.agents.updateTargetNets()
    assert not torch.equal(before, env.agents.acceptor.policy.detach())
    assert torch.equal(env.agents.acceptor.target, env.agents.acceptor.policy.detach())
    assert len(env.agents.acceptor.memory) == 12 * B
    env.close()
