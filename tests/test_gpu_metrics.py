"""GPU: device-side episode aggregates against the CPU oracle and numpy."""
import numpy as np
import pytest

from test_gpu_step_parity import DOMS, _env, _has_fused, random_actions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", ["lane", "coop", "warp", "fused1", "fused4"])
@pytest.mark.parametrize("key", ["cfg3", "cfg2", "agg", "odd"])
def test_episode_statistics_and_sums(key, impl):
    from marl_scheduling_b200.metrics import EpisodeMetrics
    from oracle import oracle as O
    dom, mode = DOMS[key]
    if impl.startswith("fused") and not _has_fused(dom):
        pytest.skip("no compile-time kernel for this domain")
    free = mode.startswith("free")
    B, T = 1500, 50
    env = _env(B, dict(dom, mode=mode), impl=impl, auction="random", spawn="philox", seed=5)
    orc = O.Oracle(B, dom, mode, tie_mode=O.TIE_PHILOX, seed=5)
    m = EpisodeMetrics(env)
    m.begin()
    rng = np.random.default_rng(8)
    acc = {k: 0.0 for k in ("offer", "price", "acceptor", "auctioneer", "amount", "qsum", "qn")}
    agent = np.zeros(dom["N"])
    for t in range(T):
        offc, a, offp = random_actions(rng, B, dom, free)
        env.step(offc, a, None, offer_price=offp)
        m.add()
        orc.step(offc, a, None, offp=offp)
        acc["offer"] += orc.r_offer.sum(); acc["acceptor"] += orc.r_acceptor.sum()
        if free:
            acc["price"] += orc.r_price.sum()
        acc["auctioneer"] += orc.r_auctioneer.sum(); agent += orc.r_agent.sum(0)
        acc["amount"] += orc.quality_cnt.sum()
        has = orc.quality_cnt > 0
        acc["qsum"] += (orc.quality_sum[has] / orc.quality_cnt[has]).sum(); acc["qn"] += has.sum()
    assert np.array_equal(m.stats[:B].cpu().numpy(), orc.stats()), (key, impl)
    s = m.summary()
    N, C, L = dom["N"], dom["C"], dom["L"]
    nOff = orc.r_offer[0].size
    nAcc = orc.r_acceptor[0].size
    assert s["coreChooserRew"] == pytest.approx(acc["offer"] / (T * B * nOff), rel=1e-12)
    assert s["acceptorRew"] == pytest.approx(acc["acceptor"] / (T * B * nAcc), rel=1e-12)
    if free:
        assert s["priceChooserRew"] == pytest.approx(acc["price"] / (T * B * nOff), rel=1e-12)
    assert s["auctioneerRew"] == pytest.approx(acc["auctioneer"] / (T * B), rel=1e-12)
    np.testing.assert_allclose(s["agentRew"], agent / (T * B), rtol=1e-12)
    assert s["acceptionAmount"] == pytest.approx(acc["amount"] / (T * B), rel=1e-12)
    assert s["acceptionQuality"] == pytest.approx(acc["qsum"] / acc["qn"], rel=1e-9)
    st = orc.stats().astype(np.int64).sum(0)
    for k in range(len(dom["prios"])):
        if st[k, 1]:
            assert s["prices"][k] == pytest.approx(st[k, 0] / st[k, 1], rel=1e-12)
        if st[k, 3]:
            assert s["dwellTimes"][k] == pytest.approx(st[k, 2] / (st[k, 3] * dom["lens"][k]), rel=1e-12)
    if mode == "fix":
        rev = sum(dom.get("mult", 1) * dom["prios"][k] * st[k, 3] for k in range(len(dom["prios"])))
        assert s["terminationRevenues"] == pytest.approx(rev / B / (T * N * C), rel=1e-12)
    m.close()
    env.close()
