"""GPU: msched_policy_step -- every PPO unit of a rollout step in one launch -- against the oracle MLP
(src/PPOmodules.py:32-39,53-63,114-125,312-332), on real observation records of the env kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG3 = dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1, 2, 3])
CFG2 = dict(N=4, C=4, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7])
CFG1 = dict(N=2, C=3, L=2, prios=[5], lens=[4], probs=[1], fix=[3], mult=2)


def _unpack(w, n, nin, A, h=16):
    o, ws = 0, []
    for sz in (h * nin, h, h * h, h, A * h, A):
        ws.append(w[n, o:o + sz]); o += sz
    return ws[0].reshape(h, nin), ws[1], ws[2].reshape(h, h), ws[3], ws[4].reshape(A, h), ws[5]


def _check_group(O, w, x, u, act, lp, pr, nin, A, nets_of_unit):
    """x [B,U,nin] float inputs, u [B,U], kernel outputs act/lp [B,U], pr [B,U,A]."""
    B, U = act.shape
    for n in range(U):
        p, a, l = O.mlp_forward(x[:, n], *_unpack(w, nets_of_unit(n), nin, A), u=u[:, n].copy())
        np.testing.assert_allclose(pr[:, n], p, rtol=2e-5, atol=1e-7)
        cdf = np.cumsum(p, 1)
        margin = np.abs(cdf - (u[:, n] * p.sum(1))[:, None]).min(1)
        sure = margin > 1e-5
        assert sure.mean() > 0.9
        assert np.array_equal(act[:, n][sure], a[sure])
        np.testing.assert_allclose(lp[:, n][sure], l[sure], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("key", ["cfg3_free", "cfg2_fix", "cfg2_shared", "cfg3_fix", "cfg1_fix", "cfg3_narrow_free", "cfg2_narrow_fix"])
def test_policy_step_matches_oracle_mlp(key, impl, monkeypatch):
    """Both kernels behind msched_policy_step -- fp32 SIMT and tcgen05 tensor cores (3xTF32) -- against the oracle.
    The "narrow" keys shrink the action counts below the BASELINE shapes' (one acceptor / core / price action fewer):
    the tensor-core instantiations built for the exact BASELINE counts do not apply and the padded ones run."""
    monkeypatch.setenv("MSCHED_POLICY_STEP_IMPL", impl)
    import torch
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    from oracle import oracle as O
    dom = {"cfg3": CFG3, "cfg2": CFG2, "cfg1": CFG1}[key[:4]]
    free = key.endswith("free")
    shared = key.endswith("shared")
    narrow = 1 if "narrow" in key else 0
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL, P = N * L, max(dom["prios"])
    B = 777
    env = BatchedSchedulingEnv(B, world_params_from_dom(dom, free), reward="free_comm" if free else "fix",
                               auction="random", spawn="philox", seed=3)
    dev, lay = env.device, env.layout
    rng = np.random.default_rng(1)
    for t in range(12):  # a state with jobs on cores, pending offers and agent-owned cores
        env.step(rng.integers(0, C + 1, (B, N, L)), rng.integers(0, 2, (B, N, C)), None,
                 offer_price=rng.integers(0, P + 1, (B, N, L)) if free else None, observe=True)
    obs = env.obs_views()
    Ua, Uo = N * C, NL
    nin_a, A_a, nin_o, A_o = 3 + 2 * NL, NL + 1 - narrow, 2 * C + 2, C + 1
    P -= narrow
    na, no = (1, 1) if shared else (Ua, Uo)
    ga = policy.MlpGroup.random(nin_a, 16, A_a, na, dev, seed=11)
    go = policy.MlpGroup.random(nin_o, 16, A_o, no, dev, seed=12)
    gp = policy.MlpGroup.random(4, 16, P + 1, Uo, dev, seed=13) if free else None
    assert policy.policy_step_supported(ga, go, gp)
    f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=dev)
    ua, uo, up = (torch.rand((B, n), device=dev, generator=torch.Generator(device=dev).manual_seed(s)) for n, s in ((Ua, 1), (Uo, 2), (Uo, 3)))
    out = {k: (i32(B, n), f32(B, n), f32(B, n, A)) for k, n, A in (("a", Ua, A_a), ("o", Uo, A_o), ("p", Uo, P + 1))}
    xa = torch.zeros((B, Ua, lay.o_acc_row), dtype=torch.int16, device=dev)
    xo = torch.zeros((B, Uo, lay.o_off_row), dtype=torch.int16, device=dev)
    xp = torch.zeros((B, Uo, 4), dtype=torch.int16, device=dev)
    env.action.zero_()

    def groups(use_u):
        A_ = policy.policy_step_group(ga, Ua, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor, 101, *out["a"][:2], x_used=xa,
                                      u=ua if use_u else None, probs=out["a"][2])
        O_ = policy.policy_step_group(go, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_core, 102, *out["o"][:2], x_used=xo,
                                      u=uo if use_u else None, probs=out["o"][2])
        P_ = policy.policy_step_group(gp, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_price, 103, *out["p"][:2], x_used=xp,
                                      u=up if use_u else None, probs=out["p"][2]) if free else None
        return A_, O_, P_
    A_, O_, P_ = groups(True)
    policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, C, A_, O_, P_, action_rec=env.action,
                       action_rec_stride=lay.action_halfs, env_offset=0, step=5, input_bound=16)
    torch.cuda.synchronize()
    acc_x = obs["acceptor"].reshape(B, Ua, nin_a).cpu().numpy()
    off_x = obs["offer"].reshape(B, Uo, nin_o).cpu().numpy()
    # the experience-buffer rows are the observation rows (acceptor rows carry one leading pad value)
    assert np.array_equal(xa[:, :, 1:1 + nin_a].cpu().numpy(), acc_x)
    assert np.array_equal(xo[:, :, :nin_o].cpu().numpy(), off_x)
    a_act, a_lp, a_pr = (t.cpu().numpy() for t in out["a"])
    o_act, o_lp, o_pr = (t.cpu().numpy() for t in out["o"])
    _check_group(O, ga.weights.cpu().numpy(), acc_x.astype(np.float32), ua.cpu().numpy(), a_act, a_lp, a_pr, nin_a, A_a,
                 (lambda n: 0) if shared else (lambda n: n))
    _check_group(O, go.weights.cpu().numpy(), off_x.astype(np.float32), uo.cpu().numpy(), o_act, o_lp, o_pr, nin_o, A_o,
                 (lambda n: 0) if shared else (lambda n: n))
    assert np.array_equal(env.acceptor_actions.reshape(B, Ua).cpu().numpy(), a_act)
    assert np.array_equal(env.offer_core_actions.reshape(B, Uo).cpu().numpy(), o_act)
    assert a_act.max() <= NL and o_act.max() <= C and len(np.unique(a_act)) > 2
    if free:
        p_act, p_lp, p_pr = (t.cpu().numpy() for t in out["p"])
        a = o_act[..., None]
        exp = np.concatenate([np.take_along_axis(off_x, np.concatenate([2 * a, 2 * a + 1], -1), 2), off_x[..., -2:]], -1)
        exp = np.where(a == 0, -5, exp).astype(np.int16)   # quirk Q1: dummy input for core action 0
        assert np.array_equal(xp.cpu().numpy(), exp)
        _check_group(O, gp.weights.cpu().numpy(), exp.astype(np.float32), up.cpu().numpy(), p_act, p_lp, p_pr, 4, P + 1, lambda n: n)
        assert np.array_equal(env.offer_price_actions.reshape(B, Uo).cpu().numpy(), np.where(o_act == 0, -5, p_act))
    # ---- the Philox draws: feeding the contract's u as overrides reproduces the un-overridden launch ----
    keep = {k: tuple(t.clone() for t in v[:2]) for k, v in out.items()}
    A_, O_, P_ = groups(False)
    off_env = (1 << 33) + 4242  # beyond 32 bits: the pair index fills both counter words
    policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, C, A_, O_, P_, env_offset=off_env, step=9, input_bound=16)
    torch.cuda.synchronize()
    drawn = {k: tuple(t.clone() for t in v[:2]) for k, v in out.items()}
    for b in (0, 1, 2, 333, 776):
        for unit in (0, Uo - 1):
            pair = (off_env + b) >> 1
            for seed, ut in ((101, ua), (102, uo)):
                if unit >= ut.shape[1]:
                    continue
                x = O.philox([pair & 0xFFFFFFFF, pair >> 32, 9, (4 << 28) | unit], [seed, 0])
                ut[b, unit] = float(np.float32(int(x[b & 1]) >> 8) * np.float32(2.0 ** -24))
                if seed == 102 and free:
                    up[b, unit] = float(np.float32(int(x[2 + (b & 1)]) >> 8) * np.float32(2.0 ** -24))
    A_, O_, P_ = groups(True)
    policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, C, A_, O_, P_, env_offset=off_env, step=9, input_bound=16)
    torch.cuda.synchronize()
    for b in (0, 1, 2, 333, 776):
        for unit in (0, Uo - 1):
            for k in ("a", "o") + (("p",) if free else ()):
                if unit < out[k][0].shape[1]:
                    assert int(out[k][0][b, unit]) == int(drawn[k][0][b, unit]), (k, b, unit)
                    assert float(out[k][1][b, unit]) == float(drawn[k][1][b, unit]), (k, b, unit)
    assert not torch.equal(drawn["a"][0], keep["a"][0])  # other draws, other actions
    env.close()


def test_policy_step_refuses_other_shapes():
    import torch
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200._lib import MschedError
    dev = torch.device("cuda", 0)
    ga = policy.MlpGroup.random(19, 16, 9, 2, dev, seed=1)   # N2 C2 L4: no instantiation
    go = policy.MlpGroup.random(6, 16, 3, 2, dev, seed=2)
    assert not policy.policy_step_supported(ga, go)
    obs = torch.zeros((8, 128), dtype=torch.int16, device=dev)
    A_ = policy.policy_step_group(ga, 2, 1, 20, 0, 1)
    O_ = policy.policy_step_group(go, 2, 64, 6, 4, 2)
    with pytest.raises(MschedError):
        policy.policy_step(obs, 128, 8, 2, A_, O_)


def test_policy_step_full_size_windows_match_oracle_and_simt(monkeypatch):
    """BASELINE size (65,536 envs, the grid the bench launches): three 128-env windows of the big batch against the
    oracle MLP (probabilities, sampled actions, log-probs, experience rows), and over ALL rows the tensor-core kernel
    against the fp32 SIMT kernel run with the same draws (same actions wherever the draw is not within 1e-5 of a CDF
    step, log-probs within 1e-4)."""
    import torch
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    from oracle import oracle as O
    dom = CFG3
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL, P, B = N * L, max(dom["prios"]), 65536
    env = BatchedSchedulingEnv(B, world_params_from_dom(dom, True), reward="free_comm", auction="random", spawn="philox", seed=4)
    dev, lay = env.device, env.layout
    g = torch.Generator(device=dev).manual_seed(2)
    for t in range(15):
        env.acceptor_actions.random_(0, 2, generator=g)
        env.offer_core_actions.random_(0, C + 1, generator=g)
        env.offer_price_actions.random_(0, P + 1, generator=g)
        env.step_observe_records()
    obs = env.obs_views()
    Ua, Uo = N * C, NL
    nin_a, A_a, nin_o, A_o = 3 + 2 * NL, NL + 1, 2 * C + 2, C + 1
    ga = policy.MlpGroup.random(nin_a, 16, A_a, Ua, dev, seed=31)
    go = policy.MlpGroup.random(nin_o, 16, A_o, Uo, dev, seed=32)
    gp = policy.MlpGroup.random(4, 16, P + 1, Uo, dev, seed=33)
    res = {}
    for impl in ("tc", "simt"):
        monkeypatch.setenv("MSCHED_POLICY_STEP_IMPL", impl)
        out = [(torch.zeros((B, n), dtype=torch.int32, device=dev), torch.zeros((B, n), device=dev),
                torch.zeros((B, n, A), device=dev)) for n, A in ((Ua, A_a), (Uo, A_o), (Uo, P + 1))]
        xs = [torch.zeros((B, Ua, lay.o_acc_row), dtype=torch.int16, device=dev), torch.zeros((B, Uo, lay.o_off_row), dtype=torch.int16, device=dev),
              torch.zeros((B, Uo, 4), dtype=torch.int16, device=dev)]
        A_ = policy.policy_step_group(ga, Ua, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor, 7, *out[0][:2], x_used=xs[0], probs=out[0][2])
        O_ = policy.policy_step_group(go, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_core, 8, *out[1][:2], x_used=xs[1], probs=out[1][2])
        P_ = policy.policy_step_group(gp, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_price, 9, *out[2][:2], x_used=xs[2], probs=out[2][2])
        policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, C, A_, O_, P_, env_offset=0, step=3, input_bound=16)
        torch.cuda.synchronize()
        res[impl] = (out, xs)
    (to, tx), (so, sx) = res["tc"], res["simt"]
    for k in range(3):
        assert torch.equal(tx[k], sx[k])                               # experience rows: bit-identical
        pt, ps = to[k][2], so[k][2]
        if k < 2:                                                      # (the price chooser's input depends on the sampled core)
            assert float((pt - ps).abs().max()) < 2e-5
        same = to[k][0] == so[k][0]
        assert float(same.float().mean()) > 0.9995
        if k < 2:
            assert float((to[k][1] - so[k][1])[same].abs().max()) < 1e-4
    acc_x = obs["acceptor"].reshape(B, Ua, nin_a)
    off_x = obs["offer"].reshape(B, Uo, nin_o)
    for w0 in (0, 31872, B - 128):                                     # windows: first, an odd tile in the middle, last
        sl = slice(w0, w0 + 128)
        for (x, grp, out, nin, A) in ((acc_x, ga, to[0], nin_a, A_a), (off_x, go, to[1], nin_o, A_o)):
            wts = grp.weights.cpu().numpy()
            xw = x[sl].cpu().numpy().astype(np.float32)
            for n in range(x.shape[1]):
                p, _, _ = O.mlp_forward(xw[:, n], *_unpack(wts, n, nin, A))
                np.testing.assert_allclose(out[2][sl, n].cpu().numpy(), p, rtol=2e-5, atol=1e-7)
                pa = np.take_along_axis(p / p.sum(1, keepdims=True), out[0][sl, n].cpu().numpy()[:, None].astype(np.int64), 1)[:, 0]
                np.testing.assert_allclose(out[1][sl, n].cpu().numpy(), np.log(np.clip(pa, 1.1920929e-07, 1 - 1.1920929e-07)), rtol=1e-4, atol=2e-5)
    env.close()


def test_policy_step_large_inputs_take_the_simt_kernel(monkeypatch):
    """input_bound > 511 (or unknown) must not reach the fp16 operand path: forcing `tc` is an error, the default call
    falls back to the fp32 SIMT kernel and agrees with the oracle on rows holding values beyond +-511."""
    import torch
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200._lib import MschedError
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    B, N, C, L = 300, 2, 3, 3
    NL, Ua, Uo = N * L, N * C, N * L
    nin_a, A_a, nin_o, A_o = 3 + 2 * NL, NL + 1, 2 * C + 2, C + 1
    RA, RO = nin_a + 1, nin_o
    OH = Ua * RA + Uo * RO + 2
    rng = np.random.default_rng(3)
    obs = torch.as_tensor(rng.integers(-900, 900, (B, OH)).astype(np.int16)).to(dev)
    ga = policy.MlpGroup.random(nin_a, 16, A_a, Ua, dev, seed=41)
    go = policy.MlpGroup.random(nin_o, 16, A_o, Uo, dev, seed=42)
    ga.weights.mul_(0.02); go.weights.mul_(0.02)   # keep the pre-activations of 900-sized inputs in range
    out = [(torch.zeros((B, n), dtype=torch.int32, device=dev), torch.zeros((B, n), device=dev), torch.zeros((B, n, A), device=dev))
           for n, A in ((Ua, A_a), (Uo, A_o))]
    A_ = policy.policy_step_group(ga, Ua, 1, RA, 0, 1, *out[0][:2], probs=out[0][2])
    O_ = policy.policy_step_group(go, Uo, Ua * RA, RO, Ua, 2, *out[1][:2], probs=out[1][2])
    monkeypatch.setenv("MSCHED_POLICY_STEP_IMPL", "tc")
    with pytest.raises(MschedError):
        policy.policy_step(obs, OH, B, C, A_, O_, input_bound=900)
    monkeypatch.delenv("MSCHED_POLICY_STEP_IMPL")
    policy.policy_step(obs, OH, B, C, A_, O_, input_bound=900)
    torch.cuda.synchronize()
    x = obs.cpu().numpy()
    acc_x = x[:, :Ua * RA].reshape(B, Ua, RA)[:, :, 1:].astype(np.float32)
    off_x = x[:, Ua * RA:Ua * RA + Uo * RO].reshape(B, Uo, RO).astype(np.float32)
    for (xw, grp, o, nin, A) in ((acc_x, ga, out[0], nin_a, A_a), (off_x, go, out[1], nin_o, A_o)):
        wts = grp.weights.cpu().numpy()
        for n in (0, xw.shape[1] - 1):
            p, _, _ = O.mlp_forward(xw[:, n], *_unpack(wts, n, nin, A))
            np.testing.assert_allclose(o[2][:, n].cpu().numpy(), p, rtol=5e-5, atol=1e-7)


@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_policy_step_two_env_shards_reproduce_the_single_shard_actions(impl, monkeypatch):
    """Sampling draws are keyed on the GLOBAL environment index (env_offset): the policy step of two env shards, each
    launched with its own offset as under data parallelism, gives the same actions, log-probs and experience rows as the
    single launch over the whole batch (exploration noise must not repeat across ranks, results must
    not depend on the number of GPUs)."""
    monkeypatch.setenv("MSCHED_POLICY_STEP_IMPL", impl)
    import torch
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    dom = CFG3
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL, P, B = N * L, max(dom["prios"]), 512
    env = BatchedSchedulingEnv(B, world_params_from_dom(dom, True), reward="free_comm", auction="random", spawn="philox", seed=8)
    dev, lay = env.device, env.layout
    g = torch.Generator(device=dev).manual_seed(3)
    for t in range(10):
        env.acceptor_actions.random_(0, 2, generator=g)
        env.offer_core_actions.random_(0, C + 1, generator=g)
        env.offer_price_actions.random_(0, P + 1, generator=g)
        env.step_observe_records()
    Ua, Uo = N * C, NL
    ga = policy.MlpGroup.random(3 + 2 * NL, 16, NL + 1, Ua, dev, seed=51)
    go = policy.MlpGroup.random(2 * C + 2, 16, C + 1, Uo, dev, seed=52)
    gp = policy.MlpGroup.random(4, 16, P + 1, Uo, dev, seed=53)
    obs = env._obs_buffer()

    def run(obs_part, n, off):
        out = [(torch.zeros((n, u), dtype=torch.int32, device=dev), torch.zeros((n, u), device=dev)) for u in (Ua, Uo, Uo)]
        A_ = policy.policy_step_group(ga, Ua, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor, 11, *out[0])
        O_ = policy.policy_step_group(go, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_core, 12, *out[1])
        P_ = policy.policy_step_group(gp, Uo, lay.o_offer, lay.o_off_row, lay.a_offer_price, 13, *out[2])
        policy.policy_step(obs_part, lay.obs_halfs, n, C, A_, O_, P_, env_offset=off, step=4, input_bound=16)
        torch.cuda.synchronize()
        return out
    whole = run(obs, B, 1000)
    lo, hi = run(obs[:256], 256, 1000), run(obs[256:], 256, 1256)
    for k in range(3):
        assert torch.equal(whole[k][0], torch.cat([lo[k][0], hi[k][0]]))
        assert torch.equal(whole[k][1], torch.cat([lo[k][1], hi[k][1]]))
    assert not torch.equal(lo[0][0], hi[0][0])   # (different environments, different draws)
    env.close()
