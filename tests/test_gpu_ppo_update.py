"""GPU: PPO.update on the device -- msched_ppo_grad / msched_adam_step against a float64
torch autograd evaluation of the reference's loss (src/PPOmodules.py:139-174), torch.optim.Adam, and the
weights the unmodified reference PPO.update produced (tests/golden/ppo_update.npz, oracle/gen_ppo_golden.py)."""
import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
H = 16


def _unpack(flat, n_in, A):
    o, out = 0, []
    for shape in ((H, n_in), (H,), (H, H), (H,), (A, H), (A,)):
        n = int(np.prod(shape))
        out.append(flat[..., o:o + n].reshape(*flat.shape[:-1], *shape))
        o += n
    return out


def _mlp(flat, x, n_in, A):
    import torch
    W1, b1, W2, b2, W3, b3 = _unpack(flat, n_in, A)
    h = torch.tanh(x @ W1.T + b1)
    h = torch.tanh(h @ W2.T + b2)
    return h @ W3.T + b3


def _reference_loss_grads(aw, cw, x, act, lp_old, G, eps_clip):
    """float64 autograd of loss.mean() for one net over its samples."""
    import torch
    from torch.distributions import Categorical
    n_in, A = x.shape[-1], None
    aw = aw.double().clone().requires_grad_(True)
    cw = cw.double().clone().requires_grad_(True)
    A = (aw.numel() - (H * n_in + H + H * H + H)) // (H + 1)
    probs = torch.softmax(_mlp(aw, x.double(), n_in, A), -1)
    dist = Categorical(probs)
    logp, ent = dist.log_prob(act.long()), dist.entropy()
    v = _mlp(cw, x.double(), n_in, 1).squeeze(-1)
    ratios = torch.exp(logp - lp_old.double())
    adv = G.double() - v.detach()
    s1 = ratios * adv
    s2 = torch.clamp(ratios, 1 - eps_clip, 1 + eps_clip) * adv
    loss = -torch.min(s1, s2) + 0.5 * torch.nn.functional.mse_loss(v, G.double()) - 0.01 * ent
    loss.mean().backward()
    stats = [(-torch.min(s1, s2)).mean().item(), ((v - G.double()) ** 2).mean().item(), ent.mean().item(), float(x.shape[0])]
    return aw.grad, cw.grad, stats


@pytest.mark.parametrize("n_in,A,units,n_nets,unit_div", [(15, 7, 6, 6, 1), (8, 4, 6, 6, 1), (4, 9, 6, 6, 1),
                                                          (27, 13, 8, 1, 1), (10, 5, 12, 4, 3), (33, 16, 2, 2, 1),
                                                          (64, 3, 2, 1, 1)])
def test_ppo_grad_matches_float64_autograd(n_in, A, units, n_nets, unit_div):
    import torch
    from marl_scheduling_b200 import policy
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(n_in * 100 + A)
    TB = 1237
    grpA = policy.MlpGroup.random(n_in, H, A, n_nets, dev, seed=3)
    grpC = policy.MlpGroup.random(n_in, H, 1, n_nets, dev, seed=4)
    # wide strides: the observation rows sit inside a larger record
    rec = torch.randint(-2, 9, (TB, units * (n_in + 3) + 5), generator=g, dtype=torch.int16).to(dev)
    x = rec.as_strided((TB, units, n_in), (rec.stride(0), n_in + 3, 1), 2)
    act = torch.randint(0, A, (TB, units), generator=g, dtype=torch.int32).to(dev)
    # old log-probs around the current ones so that both clip branches and the unclipped one occur
    lp_old = (torch.log(torch.tensor(1.0 / A)) + 0.4 * torch.randn(TB, units, generator=g)).to(dev)
    G = torch.randn(TB, units, generator=g).to(dev)
    per_net = {n: [u for u in range(units) if (u // unit_div) % n_nets == n] for n in range(n_nets)}
    net_ids = torch.arange(n_nets, dtype=torch.int32, device=dev)
    unit_ids = torch.tensor([per_net[n] for n in range(n_nets)], dtype=torch.int32, device=dev)
    ga = torch.full_like(grpA.weights, 7.0)
    gc = torch.full_like(grpC.weights, 7.0)
    stats, ws = policy.ppo_grad(grpA.weights, grpC.weights, n_in, A, x, act, lp_old, G, net_ids, unit_ids, ga, gc)
    ga2, gc2 = torch.zeros_like(ga), torch.zeros_like(gc)
    policy.ppo_grad(grpA.weights, grpC.weights, n_in, A, x, act, lp_old, G, net_ids, unit_ids, ga2, gc2, workspace=ws)
    torch.cuda.synchronize()
    assert torch.equal(ga, ga2) and torch.equal(gc, gc2)  # fixed-order reduction: bit-reproducible
    for n in range(n_nets):
        us = per_net[n]
        xs = x[:, us].reshape(-1, n_in).cpu()
        ra, rc, rstats = _reference_loss_grads(grpA.weights[n].cpu(), grpC.weights[n].cpu(), xs, act[:, us].reshape(-1).cpu(),
                                               lp_old[:, us].reshape(-1).cpu(), G[:, us].reshape(-1).cpu(), 0.2)
        for got, ref in ((ga[n], ra), (gc[n], rc)):
            got, ref = got.cpu().double().numpy(), ref.numpy()
            scale = np.abs(ref).max()
            np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5 * scale)
        np.testing.assert_allclose(stats[n].cpu().numpy(), rstats, rtol=2e-5, atol=1e-6)


def test_ppo_grad_unit_subset_leaves_other_nets_untouched():
    """The CENTRALISATION_SAMPLE rule (src/SchedulingEnvironment.py:314-329) updates a subset of units / nets."""
    import torch
    from marl_scheduling_b200 import policy
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    TB, units, n_in, A = 300, 6, 8, 4
    grpA = policy.MlpGroup.random(n_in, H, A, 3, dev, seed=3)
    grpC = policy.MlpGroup.random(n_in, H, 1, 3, dev, seed=4)
    x = torch.randint(-1, 6, (TB, units, n_in), generator=g, dtype=torch.int16).to(dev)
    act = torch.randint(0, A, (TB, units), generator=g, dtype=torch.int32).to(dev)
    lp_old = torch.full((TB, units), float(np.log(1.0 / A)), device=dev)
    G = torch.randn(TB, units, generator=g).to(dev)
    ga = torch.full_like(grpA.weights, 7.0)
    gc = torch.full_like(grpC.weights, 7.0)
    net_ids = torch.tensor([2, 0], dtype=torch.int32, device=dev)
    unit_ids = torch.tensor([[5], [3]], dtype=torch.int32, device=dev)   # net 2 <- unit 5, net 0 <- unit 3
    policy.ppo_grad(grpA.weights, grpC.weights, n_in, A, x, act, lp_old, G, net_ids, unit_ids, ga, gc)
    assert (ga[1] == 7).all() and (gc[1] == 7).all()
    for n, u in ((2, 5), (0, 3)):
        ra, rc, _ = _reference_loss_grads(grpA.weights[n].cpu(), grpC.weights[n].cpu(), x[:, u].cpu(), act[:, u].cpu(),
                                          lp_old[:, u].cpu(), G[:, u].cpu(), 0.2)
        np.testing.assert_allclose(ga[n].cpu().numpy(), ra.numpy(), rtol=1e-4, atol=1e-5 * float(ra.abs().max()))
        np.testing.assert_allclose(gc[n].cpu().numpy(), rc.numpy(), rtol=1e-4, atol=1e-5 * float(rc.abs().max()))


def test_adam_step_matches_torch_adam():
    import torch
    from marl_scheduling_b200 import policy
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(9)
    p0 = torch.randn(5000, generator=g)
    ref = torch.nn.Parameter(p0.clone().to(dev))
    opt = torch.optim.Adam([ref], lr=3e-4)
    p, m, v = p0.clone().to(dev), torch.zeros(5000, device=dev), torch.zeros(5000, device=dev)
    for step in range(1, 8):
        gr = (torch.randn(5000, generator=g) * 10 ** float(torch.randint(-6, 2, (1,), generator=g))).to(dev)
        ref.grad = gr.clone()
        opt.step()
        policy.adam_step(p, gr, m, v, 3e-4, step)
        np.testing.assert_allclose(p.cpu().numpy(), ref.detach().cpu().numpy(), rtol=0, atol=2e-7)


def test_update_reproduces_the_reference_ppo_update():
    """The unmodified reference PPO.update (one world: B = 1) on a recorded buffer: returns, K epochs of
    gradient + Adam through the kernels must land on the reference's weights."""
    import torch
    from marl_scheduling_b200 import policy
    dev = torch.device("cuda", 0)
    z = np.load(__import__("os").path.join(GOLDEN, "ppo_update.npz"))
    for tag in [t[:-len(".states")] for t in z.files if t.endswith(".states")]:
        n_in, A, K = int(z[tag + ".n_in"]), int(z[tag + ".A"]), int(z[tag + ".K"])
        gamma, eps_clip = float(z[tag + ".gamma"]), float(z[tag + ".eps_clip"])
        lr_a, lr_c = float(z[tag + ".lr_actor"]), float(z[tag + ".lr_critic"])
        x = torch.as_tensor(z[tag + ".states"].astype(np.int16)).to(dev).view(-1, 1, n_in)
        T = x.shape[0]
        act = torch.as_tensor(z[tag + ".actions"].astype(np.int32)).to(dev).view(T, 1)
        lp_old = torch.as_tensor(z[tag + ".logprobs"].astype(np.float32)).to(dev).view(T, 1)
        r = torch.as_tensor(z[tag + ".rewards"].astype(np.float32)).to(dev).view(T, 1)
        aw = torch.as_tensor(z[tag + ".actor0"]).to(dev).view(1, -1).clone()
        cw = torch.as_tensor(z[tag + ".critic0"]).to(dev).view(1, -1).clone()
        G = policy.returns(r, gamma, normalise=True)
        ids = torch.zeros((1, 1), dtype=torch.int32, device=dev)
        ga, gc = torch.zeros_like(aw), torch.zeros_like(cw)
        ma, va, mc, vc = (torch.zeros_like(t) for t in (aw, aw, cw, cw))
        ws = None
        for k in range(1, K + 1):
            _, ws = policy.ppo_grad(aw, cw, n_in, A, x, act, lp_old, G, ids.view(-1), ids, ga, gc, eps_clip=eps_clip, workspace=ws)
            policy.adam_step(aw, ga, ma, va, lr_a, k)
            policy.adam_step(cw, gc, mc, vc, lr_c, k)
        for got, ref, start, lr in ((aw, z[tag + ".actor1"], z[tag + ".actor0"], lr_a),
                                    (cw, z[tag + ".critic1"], z[tag + ".critic0"], lr_c)):
            got = got.cpu().numpy().reshape(-1)
            diff = np.abs(got - ref.reshape(-1))
            # Adam divides by sqrt(v): an entry whose gradient nearly cancels may take a different-sized step
            # of at most lr per epoch; everything else must agree to float32 round-off
            assert diff.max() <= 2.0 * lr * K, (tag, diff.max())
            assert np.mean(diff > 1e-5) < 0.02, (tag, np.mean(diff > 1e-5), diff.max())
            # and the update must have moved the weights by much more than the disagreement
            moved = np.abs(ref.reshape(-1) - start.reshape(-1))
            assert np.median(diff) < 1e-2 * np.median(moved) + 1e-7


@pytest.mark.parametrize("n_nets,unit_div,subset", [(6, 1, None), (1, 1, None), (2, 3, None), (1, 1, [1, 4])])
def test_batched_ppo_update_kernels_match_the_autograd_path(monkeypatch, n_nets, unit_div, subset):
    """BatchedPPO.update through msched_ppo_grad + msched_adam_step against the PyTorch autograd + torch.optim.Adam
    version of the same update (divided, globally shared, locally shared nets, and a unit subset)."""
    import torch
    from marl_scheduling_b200.agents import BatchedPPO
    dev = torch.device("cuda", 0)
    B, U, n_in, A, T = 37, 6, 15, 7, 40
    kw = dict(lr_actor=3e-4, lr_critic=1e-3, gamma=0.9, eps_clip=0.2, k_epochs=4, device=dev, seed=5)
    g = torch.Generator().manual_seed(1)
    xs = [torch.randint(-1, 9, (B, U, n_in), generator=g, dtype=torch.int16).to(dev) for _ in range(T)]
    rs = [torch.randint(-5, 9, (B, U), generator=g).float().to(dev) for _ in range(T)]
    out = {}
    for mode in ("kernel", "autograd"):
        monkeypatch.setenv("MSCHED_PPO_UPDATE", mode)
        ppo = BatchedPPO(n_in, A, 16, n_nets, U, unit_div, **kw)
        assert ppo.use_kernels == (mode == "kernel")
        for t in range(T):
            ppo.selectAction(xs[t], n_in, U * n_in, B, seed=3)
            ppo.saveReward(rs[t])
        a0 = ppo.actor.detach().clone()
        mse = ppo.update(subset)
        out[mode] = (ppo.actor.detach().cpu().numpy(), ppo.critic.detach().cpu().numpy(), mse,
                     (ppo.actor.detach() - a0).abs().max().item())
    (ak, ck, mk, moved), (aa, ca, ma, _) = out["kernel"], out["autograd"]
    assert moved > 5e-4
    assert abs(mk - ma) < 1e-4 * max(1.0, abs(ma))
    for got, ref in ((ak, aa), (ck, ca)):
        diff = np.abs(got - ref)
        assert diff.max() <= 2 * 1e-3 * 4
        assert np.mean(diff > 1e-5) < 0.02, (np.mean(diff > 1e-5), diff.max())
