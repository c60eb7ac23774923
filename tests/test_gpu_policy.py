"""GPU: actor forward / categorical log-prob / returns kernels against the reference's torch
outputs (tests/golden/torch_vectors.npz) and the CPU oracle.  Floating point: 1e-5 relative."""
import os

import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tc", "simt"])
def actor_impl(request, monkeypatch):
    """The per-group actor kernels: tcgen05 tensor cores (3xTF32) and fp32 SIMT (shapes one does not serve fall back
    to the default choice)."""
    monkeypatch.setenv("MSCHED_ACTOR_IMPL", request.param)
    return request.param

TAGS = dict(acc_cfg3=(15, 7, 16), off_cfg3=(8, 4, 16), price_cfg3=(4, 9, 16), acc_cfg2=(27, 13, 16),
            off_cfg2=(10, 5, 16), aggoff=(12, 64, 32))


def test_actor_forward_matches_torch_reference(actor_impl):
    import torch
    from marl_scheduling_b200 import policy
    z = np.load(os.path.join(GOLDEN, "torch_vectors.npz"))
    dev = torch.device("cuda", 0)
    for tag, (nin, A, h) in TAGS.items():
        sd = {k[len(tag) + 1:]: z[k] for k in z.files if k.startswith(tag + ".actor")}
        grp = policy.MlpGroup.from_state_dicts(nin, h, A, [sd], dev)
        x = torch.as_tensor(z[tag + ".x"]).to(torch.int16).to(dev)
        M = x.shape[0]
        # u chosen so that the inverse CDF lands on the torch-sampled action
        probs_ref = z[tag + ".probs"]
        cdf = np.cumsum(probs_ref, 1)
        a_ref = z[tag + ".action"]
        lo = np.where(a_ref > 0, cdf[np.arange(M), a_ref - 1], 0.0)
        hi = cdf[np.arange(M), a_ref]
        u = ((lo + hi) / 2 / cdf[:, -1]).astype(np.float32)
        act, lp, pr = policy.actor_forward(grp, x, nin, 1, M, u=u, want_probs=True)
        np.testing.assert_allclose(pr.cpu().numpy(), probs_ref, rtol=2e-5, atol=1e-7)
        sure = (hi - lo) > 1e-4
        assert np.array_equal(act.cpu().numpy()[sure], a_ref[sure])
        np.testing.assert_allclose(lp.cpu().numpy()[sure], z[tag + ".logprob"][sure], rtol=1e-4, atol=2e-5)


def test_actor_forward_grouped_nets_and_strides_match_oracle(actor_impl):
    import torch
    from marl_scheduling_b200 import policy
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    n_envs, units, nin, h, A, stride, env_stride = 777, 6, 15, 16, 7, 15, 119
    grp = policy.MlpGroup.random(nin, h, A, units, dev, seed=4)
    xs = rng.integers(-2, 12, (n_envs, env_stride)).astype(np.int16)
    u = rng.random(n_envs * units).astype(np.float32)
    act, lp, pr = policy.actor_forward(grp, torch.as_tensor(xs).to(dev), stride, units, n_envs,
                                       env_stride=env_stride, u=u, want_probs=True)
    act, lp, pr = act.cpu().numpy(), lp.cpu().numpy(), pr.cpu().numpy()
    w = grp.weights.cpu().numpy()
    for n in range(units):
        o, ws = 0, []
        for sz in (h * nin, h, h * h, h, A * h, A):
            ws.append(w[n, o:o + sz]); o += sz
        x = xs[:, n * stride: n * stride + nin].astype(np.float32)
        p, a, l = O.mlp_forward(x, ws[0].reshape(h, nin), ws[1], ws[2].reshape(h, h), ws[3],
                                ws[4].reshape(A, h), ws[5], u=u[n::units])
        np.testing.assert_allclose(pr[n::units], p, rtol=2e-5, atol=1e-7)
        cdf = np.cumsum(p, 1)
        thr = (u[n::units] * p.sum(1))[:, None]
        margin = np.abs(cdf - thr).min(1)
        sure = margin > 1e-5
        assert np.array_equal(act[n::units][sure], a[sure])
        np.testing.assert_allclose(lp[n::units][sure], l[sure], rtol=1e-4, atol=2e-5)


def test_actor_sampling_is_distributionally_correct(actor_impl):
    import torch
    from marl_scheduling_b200 import policy
    dev = torch.device("cuda", 0)
    grp = policy.MlpGroup.random(8, 16, 4, 1, dev, seed=5)
    x = torch.zeros((200000, 8), dtype=torch.int16, device=dev)
    act, lp, pr = policy.actor_forward(grp, x, 8, 1, 200000, seed=11, step=3, want_probs=True)
    p = pr[0].cpu().numpy()
    freq = np.bincount(act.cpu().numpy(), minlength=4) / 200000
    assert np.abs(freq - p).max() < 5e-3
    act2, _, _ = policy.actor_forward(grp, x, 8, 1, 200000, seed=11, step=4)
    assert (act2 != act).any()


def test_returns_match_reference_and_oracle():
    import torch
    from marl_scheduling_b200 import policy
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    z = np.load(os.path.join(GOLDEN, "torch_vectors.npz"))
    for tag in ("ret_a", "ret_b", "ret_c", "ret_d"):
        r = torch.as_tensor(z[tag + ".r"], dtype=torch.float32)[:, None].to(dev)
        g = float(z[tag + ".gamma"])
        raw = policy.returns(r, g, normalise=False).cpu().numpy()[:, 0]
        assert np.array_equal(raw, z[tag + ".raw"].astype(np.float32))
        nrm = policy.returns(r, g, normalise=True).cpu().numpy()[:, 0]
        np.testing.assert_allclose(nrm, z[tag + ".norm"], rtol=1e-5, atol=1e-6)
    rng = np.random.default_rng(1)
    r = rng.integers(-9, 15, (200, 5000)).astype(np.float32)
    out = policy.returns(torch.as_tensor(r).to(dev), 0.8733333333333333).cpu().numpy()
    np.testing.assert_allclose(out, O.returns(r.astype(np.float64), 0.8733333333333333), rtol=1e-5, atol=1e-5)
    # scan property: raw returns satisfy G_t - gamma*G_{t+1} = r_t
    raw = policy.returns(torch.as_tensor(r).to(dev), 0.5, normalise=False).cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(raw[:-1] - 0.5 * raw[1:], r[:-1], rtol=0, atol=2e-5)
    # the TMA-tiled kernel (16-byte aligned rows, tile in shared memory) and the streaming kernel (any M, any T)
    # do the same arithmetic: bit-equal results
    for T, M in ((200, 5000), (57, 4096), (3, 132), (500, 260), (1000, 100), (1600, 36)):   # tiles of 128 / 64 / 32 columns
        r = rng.integers(-9, 15, (T, M + 1)).astype(np.float32)
        for norm in (True, False):
            a = policy.returns(torch.as_tensor(np.ascontiguousarray(r[:, :M])).to(dev), 0.9, normalise=norm)   # tiled
            b = policy.returns(torch.as_tensor(r).to(dev), 0.9, normalise=norm)[:, :M]                         # M + 1: streaming
            assert torch.equal(a, b), (T, M, norm)
    big = rng.integers(-9, 15, (1700, 256)).astype(np.float32)   # no tile width fits shared memory: streaming kernel
    np.testing.assert_allclose(policy.returns(torch.as_tensor(big).to(dev), 0.95).cpu().numpy(),
                               O.returns(big.astype(np.float64), 0.95), rtol=1e-5, atol=1e-5)


def test_price_chooser_gather_and_action_record(actor_impl):
    """FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332): the price chooser sees
    [obs[2a], obs[2a+1], obs[-2], obs[-1]] of the core chooser's action a, the dummy [-5]*4 and a
    reported price of -5 for a == 0; the kernel also writes the reported action into the action
    record."""
    import torch
    from marl_scheduling_b200 import policy
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(8)
    n_envs, units, C, A, h = 333, 6, 3, 9, 16
    row = 2 * C + 2
    env_stride = units * row + 5
    xs = rng.integers(-1, 9, (n_envs, env_stride)).astype(np.int16)
    core = rng.integers(0, C + 1, (n_envs, units)).astype(np.int32)
    u = rng.random(n_envs * units).astype(np.float32)
    grp = policy.MlpGroup.random(4, h, A, units, dev, seed=12)
    rec = torch.full((n_envs, 32), 77, dtype=torch.int16, device=dev)
    x_used = torch.zeros((n_envs * units, 4), dtype=torch.int16, device=dev)
    act, lp, pr = policy.actor_forward(grp, torch.as_tensor(xs).to(dev), row, units, n_envs,
                                       env_stride=env_stride, u=u, want_probs=True,
                                       action_rec=rec[:, 10:], action_rec_stride=32,
                                       gather_core=torch.as_tensor(core.reshape(-1)).to(dev), n_cores=C,
                                       x_used=x_used)
    rows = xs[:, : units * row].reshape(n_envs, units, row)
    a = core[..., None]
    exp = np.concatenate([np.take_along_axis(rows, np.concatenate([2 * a, 2 * a + 1], -1), 2), rows[..., -2:]], -1)
    exp = np.where(a == 0, -5, exp).astype(np.int16)
    assert np.array_equal(x_used.cpu().numpy().reshape(n_envs, units, 4), exp)
    act, pr = act.cpu().numpy().reshape(n_envs, units), pr.cpu().numpy().reshape(n_envs, units, A)
    w = grp.weights.cpu().numpy()
    for n in range(units):
        o, ws = 0, []
        for sz in (h * 4, h, h * h, h, A * h, A):
            ws.append(w[n, o:o + sz]); o += sz
        p, _, _ = O.mlp_forward(exp[:, n].astype(np.float32), ws[0].reshape(h, 4), ws[1], ws[2].reshape(h, h),
                                ws[3], ws[4].reshape(A, h), ws[5])
        np.testing.assert_allclose(pr[:, n], p, rtol=2e-5, atol=1e-7)
    r = rec.cpu().numpy()
    assert (r[:, :10] == 77).all() and (r[:, 10 + units:] == 77).all()
    assert np.array_equal(r[:, 10:10 + units], np.where(core == 0, -5, act))


@pytest.mark.parametrize("shape", [(45, 32, 343, 2), (12, 32, 64, 2), (40, 64, 1323, 1), (30, 16, 130, 3)])
def test_aggregated_head_matches_fp32_reference(shape):
    """Aggregated action heads (src/PPOmodules.py:177-232): the tensor-core kernel tiles the last
    layer over the actions and never materialises the logits; probabilities, the inverse-CDF
    sample and the log-prob are compared with a float64 evaluation of the same net."""
    import torch
    from marl_scheduling_b200 import policy
    nin, h, A, units = shape
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(21)
    n_envs = 333
    grp = policy.MlpGroup.random(nin, h, A, units, dev, seed=31)
    xs = rng.integers(-2, 12, (n_envs, units * nin)).astype(np.int16)
    u = rng.random(n_envs * units).astype(np.float32)
    act, lp, pr = policy.actor_forward(grp, torch.as_tensor(xs).to(dev), nin, units, n_envs, u=u, want_probs=True)
    act, lp, pr = act.cpu().numpy(), lp.cpu().numpy(), pr.cpu().numpy()
    w = grp.weights.cpu().numpy().astype(np.float64)
    for n in range(units):
        o, ws = 0, []
        for sz in (h * nin, h, h * h, h, A * h, A):
            ws.append(w[n, o:o + sz]); o += sz
        x = xs[:, n * nin:(n + 1) * nin].astype(np.float64)
        h1 = np.tanh(x @ ws[0].reshape(h, nin).T + ws[1])
        h2 = np.tanh(h1 @ ws[2].reshape(h, h).T + ws[3])
        lg = h2 @ ws[4].reshape(A, h).T + ws[5]
        p = np.exp(lg - lg.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
        np.testing.assert_allclose(pr[n::units], p, rtol=5e-5, atol=1e-9)
        cdf = np.cumsum(p, 1)
        thr = u[n::units].astype(np.float64)[:, None]
        a_ref = (cdf > thr).argmax(1)
        margin = np.abs(cdf - thr).min(1)
        sure = margin > 1e-5
        assert sure.mean() > 0.9
        assert np.array_equal(act[n::units][sure], a_ref[sure])
        np.testing.assert_allclose(lp[n::units][sure], np.log(p[np.arange(n_envs), a_ref])[sure], rtol=1e-4, atol=5e-5)


def test_aggregated_head_sampling_distribution():
    import torch
    from marl_scheduling_b200 import policy
    dev = torch.device("cuda", 0)
    grp = policy.MlpGroup.random(12, 32, 200, 1, dev, seed=5)
    x = torch.ones((100000, 12), dtype=torch.int16, device=dev)
    act, lp, pr = policy.actor_forward(grp, x, 12, 1, 100000, seed=3, step=1, want_probs=True)
    p = pr[0].cpu().numpy()
    freq = np.bincount(act.cpu().numpy(), minlength=200) / 100000
    assert np.abs(freq - p).max() < 5e-3
    assert abs(float(lp.exp().mean().cpu()) - float((p * p).sum())) < 2e-3   # E[p(a)] = sum p^2
