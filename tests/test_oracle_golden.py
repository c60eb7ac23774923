"""CPU: the C restatement (oracle/) replayed against traces of the unmodified reference.

The golden traces were produced by oracle/gen_golden.py from /root/reference;
this is what pins the oracle.  Everything integer is compared bit-exactly; float rewards exactly
(they are small dyadic rationals); the acception-quality mean within 1e-12 relative."""
import numpy as np
import pytest

from helpers import assert_state_equal, golden_names, load_golden
from oracle import oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_oracle_replays_reference_trace(name):
    tr, meta = load_golden(name)
    T = tr["done"].shape[0]
    free = meta["mode"].startswith("free")
    orc = O.Oracle(1, meta, meta["mode"], chain_cap=32)
    assert_state_equal(orc.export(0), tr, None, prefix=name)
    obs = orc.observe(0)
    for k in ("obs_acc", "obs_off", "obs_auc", "ids", "auc_ids"):
        assert np.array_equal(obs[k], tr["init_" + k].astype(np.int32)), (name, "init", k)
    for t in range(T):
        orc.step(tr["in_offc"][t][None], tr["in_acc"][t][None], tr["in_auc"][t][None],
                 offp=tr["in_offp"][t][None] if free else None, spawn_u=tr["in_spawn_u"][t][None])
        e = orc.export(0)
        assert e["flags"] == 0, (name, t, e["flags"])
        assert_state_equal(e, tr, t, prefix=name)
        assert e["round"] == int(tr["round"][t])
        assert orc.n_accepted[0] == tr["n_accepted"][t] and orc.n_term[0] == tr["n_term"][t]
        na, nt = int(tr["n_accepted"][t]), int(tr["n_term"][t])
        assert np.array_equal(e["accepted"][:na], tr["accepted"][t][:na])
        assert np.array_equal(e["term"][:nt, :5], tr["term"][t][:nt])
        assert np.array_equal(orc.r_offer[0], tr["r_offer"][t]), (name, t, "r_offer")
        assert np.array_equal(orc.r_price[0], tr["r_price"][t]), (name, t, "r_price")
        assert np.array_equal(orc.r_acceptor[0], tr["r_acceptor"][t]), (name, t, "r_acceptor")
        assert np.array_equal(orc.r_auctioneer[0], tr["r_auctioneer"][t]), (name, t)
        assert np.array_equal(orc.r_agent[0], tr["r_agent"][t]), (name, t, "r_agent")
        assert orc.done[0] == tr["done"][t]
        assert orc.quality_cnt[0] == tr["quality_cnt"][t]
        if tr["quality_cnt"][t] > 0:
            mean = orc.quality_sum[0] / orc.quality_cnt[0]
            assert mean == pytest.approx(float(tr["quality"][t]), rel=1e-12, abs=1e-12)
        obs = orc.observe(0)
        if meta["agent_kind"] != "aggregated":
            for k in ("obs_acc", "obs_off"):
                assert np.array_equal(obs[k], tr[k][t].astype(np.int32)), (name, t, k)
        for k in ("obs_auc", "ids", "auc_ids"):
            assert np.array_equal(obs[k], tr[k][t].astype(np.int32)), (name, t, k)
    if "term_revenue" in tr and meta["mode"] == "fix":
        assert e["term_revenue"] == int(tr["term_revenue"][-1])


def test_oracle_semi_aggregated_observation_layout():
    """AggregatedAgent layouts (src/Agent.py:82-140) are concatenations of the divided blocks."""
    tr, meta = load_golden("aggobs_E")
    N, C, L = meta["N"], meta["C"], meta["L"]
    orc = O.Oracle(1, meta, meta["mode"])
    for t in range(tr["done"].shape[0]):
        orc.step(tr["in_offc"][t][None], tr["in_acc"][t][None], tr["in_auc"][t][None],
                 spawn_u=tr["in_spawn_u"][t][None])
        obs = orc.observe(0)
        acc = obs["obs_acc"].reshape(N, -1).astype(np.float32)
        assert np.array_equal(acc, tr["obs_acc"][t])
        cores = obs["obs_off"][:, 0, : 2 * C]
        slots = obs["obs_off"][:, :, 2 * C:].reshape(N, 2 * L)
        assert np.array_equal(np.concatenate([cores, slots], 1), tr["obs_off"][t].astype(np.int32))


def test_oracle_hardcoded_auctioneer_matches_reference_choice_when_unique():
    """With in-oracle auction (auc=None, first-arg-max) the choice equals the recorded reference
    auctioneer action whenever the arg-max is unique (ties are random in the reference)."""
    tr, meta = load_golden("kat_B")  # KAT protocol = first arg-max: must match always
    orc = O.Oracle(1, meta, meta["mode"])
    for t in range(tr["done"].shape[0]):
        orc.step(tr["in_offc"][t][None], tr["in_acc"][t][None], None,
                 spawn_u=tr["in_spawn_u"][t][None])
        assert np.array_equal(orc.auc_out[0], tr["in_auc"][t]), t
        assert_state_equal(orc.export(0), tr, t)


def test_oracle_returns_and_mlp_match_torch_vectors():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "torch_vectors.npz"))
    for tag in ("ret_a", "ret_b", "ret_c", "ret_d"):
        r = z[tag + ".r"]
        raw = O.returns(r[:, None], float(z[tag + ".gamma"]), normalise=False)[:, 0]
        assert np.array_equal(raw, z[tag + ".raw"].astype(np.float32))
        nrm = O.returns(r[:, None], float(z[tag + ".gamma"]), normalise=True)[:, 0]
        np.testing.assert_allclose(nrm, z[tag + ".norm"], rtol=1e-5, atol=1e-6)
    # SURVEY App. C arithmetic vector
    nrm = O.returns(np.array([[1.0], [0.0], [2.0]]), 0.5)[:, 0]
    np.testing.assert_allclose(nrm, [0.0, -0.9999998, 0.9999998], atol=1e-6)
    for tag in ("acc_cfg3", "off_cfg3", "price_cfg3", "acc_cfg2", "off_cfg2", "aggoff"):
        w = [z[f"{tag}.actor.{i}.{p}"] for i in (0, 2, 4) for p in ("weight", "bias")]
        probs, _, _ = O.mlp_forward(z[tag + ".x"], *w)
        np.testing.assert_allclose(probs, z[tag + ".probs"], rtol=2e-5, atol=1e-7)
        wc = [z[f"{tag}.critic.{i}.{p}"] for i in (0, 2, 4) for p in ("weight", "bias")]
        val, _, _ = O.mlp_forward(z[tag + ".x"], *wc, softmax=False)
        np.testing.assert_allclose(val[:, 0], z[tag + ".value"], rtol=2e-5, atol=2e-6)
        # log-prob of the torch-sampled action via Categorical semantics
        p = probs[np.arange(len(probs)), z[tag + ".action"]] / probs.sum(1)
        np.testing.assert_allclose(np.log(p), z[tag + ".logprob"], rtol=1e-4, atol=1e-5)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, exp in kat:
        assert O.philox(ctr, key).tolist() == exp


def test_arithmetic_vectors():
    """round-half-even on float64 products."""
    import math
    for p, t, d, exp in [(3, 6, 1, 0), (3, 6, 3, 2), (10, 3, 2, 7), (7, 3, 3, 7)]:
        assert round(p / t * d) == exp
    assert [round(0.5), round(1.5), round(2.5)] == [0, 2, 2]
    assert sum([1 / 6] * 6) == 1.0 or math.isclose(sum([1 / 6] * 6), 1.0)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_episode_statistics_match_reference_records(name):
    """The per-kind episode statistics (accepted prices, dwell times) against what the train scripts
    read off the reference's world.acceptedOffers and world.verweilzeiten (src/trainPPO.py:172-227)."""
    tr, meta = load_golden(name)
    T = tr["done"].shape[0]
    free = meta["mode"].startswith("free")
    prios, lens = list(meta["prios"]), list(meta["lens"])
    J = len(prios)
    orc = O.Oracle(1, meta, meta["mode"], chain_cap=32)
    exp = np.zeros((J, 4), np.int64)
    kind_before = tr["init_slot_kind"]
    for t in range(T):
        for k in range(int(tr["n_accepted"][t])):
            offerer, _, _, slot, price = (int(v) for v in tr["accepted"][t][k])
            kind = int(kind_before[offerer - 1][slot])   # offer.jobKind: the offered job sat in that slot
            exp[kind, 0] += price
            exp[kind, 1] += 1
        kind_before = tr["slot_kind"][t]
        orc.step(tr["in_offc"][t][None], tr["in_acc"][t][None], tr["in_auc"][t][None],
                 offp=tr["in_offp"][t][None] if free else None, spawn_u=tr["in_spawn_u"][t][None])
    if len(set(zip(prios, lens))) == J:   # a Verweilzeit record names its kind by (priority, length)
        n_rec = int(tr["n_term"].sum())   # (a trace may record fewer steps than the reference ran)
        for (prio, ln, dwell), norm in zip(tr["dwell"][:n_rec], tr["dwell_norm"][:n_rec]):
            kind = list(zip(prios, lens)).index((int(prio), int(ln)))
            exp[kind, 2] += int(dwell) - 1
            exp[kind, 3] += 1
            assert norm == (int(dwell) - 1) / int(ln)
        assert np.array_equal(orc.stats()[0].astype(np.int64), exp), name
    else:
        assert np.array_equal(orc.stats()[0][:, :2].astype(np.int64), exp[:, :2]), name
    assert int(orc.stats()[0][:, 3].sum()) == int(tr["n_term"].sum())


def test_ppo_update_oracle_reproduces_the_reference_update():
    from helpers import GOLDEN
    """oracle.ppo_update (hand-written float64 backward + Adam) against the weights the unmodified reference
    PPO.update produced (tests/golden/ppo_update.npz, oracle/gen_ppo_golden.py)."""
    from oracle import oracle as O
    z = np.load(__import__("os").path.join(GOLDEN, "ppo_update.npz"))
    tags = [t[:-len(".states")] for t in z.files if t.endswith(".states")]
    assert len(tags) == 3
    for tag in tags:
        K = int(z[tag + ".K"])
        a, c = O.ppo_update(z[tag + ".actor0"], z[tag + ".critic0"], z[tag + ".states"], z[tag + ".actions"].astype(np.int64),
                            z[tag + ".logprobs"].astype(np.float64), z[tag + ".rewards"], float(z[tag + ".gamma"]),
                            float(z[tag + ".eps_clip"]), K, float(z[tag + ".lr_actor"]), float(z[tag + ".lr_critic"]))
        for got, ref, lr in ((a, z[tag + ".actor1"], float(z[tag + ".lr_actor"])), (c, z[tag + ".critic1"], float(z[tag + ".lr_critic"]))):
            diff = np.abs(got - ref.astype(np.float64))
            assert diff.max() <= 2.0 * lr * K, (tag, diff.max())
            assert np.mean(diff > 2e-6) < 0.02, (tag, np.mean(diff > 2e-6), diff.max())


def test_oracle_dqn_matches_reference_fixtures():
    """tests/golden/dqn.npz comes from the UNMODIFIED src/DQNmodules.py (oracle/gen_dqn_golden.py): Q-values,
    epsilon-greedy choices with the recorded draws, and the weights after three optimize_model steps."""
    import os
    z = np.load(os.path.join(__import__("helpers").GOLDEN, "dqn.npz"))
    for tag in ("acc_cfg2", "off_cfg2", "acc_cfg3"):
        A = int(z[tag + ".n_actions"])
        w0 = z[tag + ".w0"]
        q, _ = O.dqn_forward(w0, z[tag + ".x"], A)
        np.testing.assert_allclose(q, z[tag + ".q"], rtol=2e-5, atol=2e-6)
        act = O.dqn_select(w0, z[tag + ".x"], A, z[tag + ".sample"], z[tag + ".randrange"], float(z[tag + ".eps"]))
        assert np.array_equal(act, z[tag + ".action"])
        assert 0.2 < (z[tag + ".sample"] > float(z[tag + ".eps"])).mean() < 0.8   # both branches exercised
        w = O.dqn_optimize(w0, w0, z[tag + ".S"], z[tag + ".A"], z[tag + ".S2"], z[tag + ".R"], z[tag + ".idx"],
                           float(z[tag + ".gamma"]), A)
        np.testing.assert_allclose(w, z[tag + ".w_after"], rtol=2e-4, atol=2e-6)
        assert np.abs(w[-1] - w0).max() > 1e-3
