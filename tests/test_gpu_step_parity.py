"""GPU: the CUDA step / observe path (through the C-ABI) against the reference traces and the
CPU oracle.  Integer state, rewards and observations are compared bit-exactly."""
import numpy as np
import pytest

from helpers import STATE_KEYS, assert_state_equal, golden_names, load_golden

pytestmark = pytest.mark.gpu


def _env(B, meta, impl=None, **kw):
    """impl: None (library default), "lane" (one lane per env), "coop" (G lanes per env), "warp" (one warp per env) or
    "fusedR" (compile-time-domain kernel with R = 1, 2 or 4 role warps per 32-env tile)."""
    import os
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom
    free = meta["mode"].startswith("free")
    old = {k: os.environ.pop(k, None) for k in ("MSCHED_STEP_IMPL", "MSCHED_ROLES")}
    if impl and impl.startswith("fused"):
        os.environ["MSCHED_STEP_IMPL"] = "fused"
        os.environ["MSCHED_ROLES"] = impl[5:]
    elif impl:
        os.environ["MSCHED_STEP_IMPL"] = impl
    try:
        env = BatchedSchedulingEnv(B, world_params_from_dom(meta, free), reward=meta["mode"],
                                   net_zero_offer_reward=meta.get("netZero", 0.5), **kw)
        if impl:
            info = env.info()
            assert info["step_impl"] == impl[:5].rstrip("124") and (
                not impl.startswith("fused") or info["threads_per_cta"] == 32 * int(impl[5:])), info
        return env
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


FUSED_DOMAINS = {(2, 3, 3), (4, 4, 3), (2, 3, 2), (2, 2, 3)}  # compile-time kernels in msched_abi.cu


def _has_fused(dom):
    return (dom["N"], dom["C"], dom["L"]) in FUSED_DOMAINS


def _one(exp, b):
    return {k: (v[b] if isinstance(v, np.ndarray) else v) for k, v in exp.items()}


@pytest.mark.parametrize("impl", ["lane", "coop", "warp", "fused1", "fused2", "fused4"])
@pytest.mark.parametrize("name", golden_names())
def test_cuda_replays_reference_trace(name, impl):
    """Recorded reference trace, replicated into B envs that straddle two tiles."""
    tr, meta = load_golden(name)
    if impl.startswith("fused") and not _has_fused(meta):
        pytest.skip("no compile-time kernel for this domain")
    fused = impl.startswith("fused")
    T = tr["done"].shape[0]
    free = meta["mode"].startswith("free")
    agg = meta["mode"] == "agg"
    B = 130
    env = _env(B, meta, impl=impl, auction="external", spawn="u64")
    e0 = env.export_state()
    for b in (0, 1, 129):
        assert_state_equal(_one(e0, b), tr, None, prefix=name)
    rep = lambda a: np.broadcast_to(np.asarray(a)[None], (B,) + np.asarray(a).shape).copy()
    for t in range(T):
        r = env.step(rep(tr["in_offc"][t]), rep(tr["in_acc"][t]), rep(tr["in_auc"][t]),
                     offer_price=rep(tr["in_offp"][t]) if free else None,
                     spawn_u=rep(tr["in_spawn_u"][t]), observe=fused)
        e = env.export_state()
        assert (e["flags"] == 0).all(), (name, t)
        for b in (0, 64, 129):
            assert_state_equal(_one(e, b), tr, t, prefix=f"{name}[env {b}]")
        r = {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}
        for b in (0, 63, 64, 129):
            assert np.array_equal(r["offer"][b].astype(np.float64), tr["r_offer"][t]), (name, t)
            if free:
                assert np.array_equal(r["price"][b].astype(np.float64), tr["r_price"][t]), (name, t)
            assert np.array_equal(r["acceptor"][b], tr["r_acceptor"][t]), (name, t, "acceptor")
            assert np.array_equal(r["auctioneer"][b], tr["r_auctioneer"][t]), (name, t)
            assert np.array_equal(r["agent"][b], tr["r_agent"][t]), (name, t, "agent")
            assert r["done"][b] == tr["done"][t]
            assert r["n_accepted"][b] == tr["n_accepted"][t]
            assert r["n_terminated"][b] == tr["n_term"][t]
            assert r["quality_cnt"][b] == tr["quality_cnt"][t]
            if tr["quality_cnt"][t] > 0:
                assert r["quality_sum"][b] / r["quality_cnt"][b] == pytest.approx(
                    float(tr["quality"][t]), rel=1e-12, abs=1e-12)
        if t % 7 == 0 or t == T - 1:
            of = {k: v.cpu().numpy() for k, v in env.obs_views().items()} if fused else None
            o = {k: v.cpu().numpy() for k, v in env.observe(with_ids=True).items()}
            if fused:  # the observations the step launch wrote == the stand-alone observe kernel's
                for k in of:
                    assert np.array_equal(of[k], o[k]), (name, t, k)
            for b in (0, 129):
                if meta["agent_kind"] != "aggregated":
                    assert np.array_equal(o["acceptor"][b], tr["obs_acc"][t]), (name, t)
                    assert np.array_equal(o["offer"][b], tr["obs_off"][t]), (name, t)
                assert np.array_equal(o["auctioneer"][b], tr["obs_auc"][t]), (name, t)
                assert np.array_equal(o["ids"][b], tr["ids"][t]), (name, t)
                assert np.array_equal(o["auctioneer_ids"][b], tr["auc_ids"][t]), (name, t)
    env.close()


DOMS = {
    "cfg3": (dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1]), "free_comm"),
    "cfg2": (dict(N=4, C=4, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7]), "fix"),
    "cfg1": (dict(N=2, C=3, L=2, prios=[5], lens=[4], probs=[1], fix=[3], mult=2), "fix"),
    "agg": (dict(N=2, C=2, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7]), "agg"),
    "F": (dict(N=8, C=8, L=4, prios=[3, 10, 6], lens=[6, 3, 2], probs=[0.5, 0.2, 0.3], fix=[2, 7, 4],
               newJobs=2), "free_ncomm"),
    "odd": (dict(N=3, C=5, L=2, prios=[1, 7, 4, 9], lens=[1, 2, 7, 3], probs=[0.1, 0.4, 0.3, 0.2],
                 fix=[1, 5, 2, 6], mult=3), "fix"),
}


def random_actions(rng, B, dom, free, p_valid_hint=None):
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL = N * L
    P = max(dom["prios"])
    acc = rng.integers(0, NL + 1, (B, N, C))
    # bias towards low indices so that real offers get accepted often
    low = rng.random((B, N, C)) < 0.6
    acc = np.where(low, rng.integers(0, 2, (B, N, C)), acc)
    offc = rng.integers(0, C + 1, (B, N, L))
    offp = rng.integers(0, P + 1, (B, N, L)) if free else None
    return offc, acc, offp


DOMS["cfg5"] = (dict(N=32, C=64, L=8, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7]), "fix")


@pytest.mark.parametrize("impl", ["lane", "coop", "warp", "fused1", "fused2", "fused4"])
@pytest.mark.parametrize("key", list(DOMS))
@pytest.mark.parametrize("auction", ["first", "random"])
def test_cuda_matches_oracle_random_batch(key, auction, impl):
    """Thousands of envs, device Philox spawn + in-kernel auction, vs the CPU oracle."""
    from oracle import oracle as O
    dom, mode = DOMS[key]
    if impl.startswith("fused") and not _has_fused(dom):
        pytest.skip("no compile-time kernel for this domain")
    fused = impl.startswith("fused")
    free = mode.startswith("free")
    B, T = (1000, 40) if key == "F" else (3000, 60)
    if key == "cfg5":
        if impl == "lane":
            pytest.skip("config 5 exceeds the lane-per-env kernel's shared-memory budget")
        B, T = (500, 40) if impl == "warp" else (96, 25)
    seed = 1234
    env = _env(B, dict(dom, mode=mode), impl=impl, auction=auction, spawn="philox", seed=seed, env_offset=77)
    orc = O.Oracle(B, dom, mode, tie_mode=O.TIE_PHILOX if auction == "random" else O.TIE_FIRST,
                   seed=seed, env_offset=77)
    rng = np.random.default_rng(5)
    for t in range(T):
        offc, acc, offp = random_actions(rng, B, dom, free)
        r = env.step(offc, acc, None, offer_price=offp, observe=fused)
        orc.step(offc, acc, None, offp=offp)
        r = {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}
        assert np.array_equal(r["auctioneer_idx"], orc.auc_out), (key, t, "auction winners")
        assert np.array_equal(r["offer"].astype(np.float64), orc.r_offer), (key, t)
        if free:
            assert np.array_equal(r["price"].astype(np.float64), orc.r_price), (key, t)
        assert np.array_equal(r["acceptor"], orc.r_acceptor), (key, t)
        assert np.array_equal(r["auctioneer"], orc.r_auctioneer), (key, t)
        assert np.array_equal(r["agent"], orc.r_agent), (key, t)
        assert np.array_equal(r["n_accepted"], orc.n_accepted), (key, t)
        assert np.array_equal(r["n_terminated"], orc.n_term), (key, t)
        assert np.array_equal(r["flags"].astype(np.uint32), orc.flags), (key, t)
        np.testing.assert_allclose(r["quality_sum"], orc.quality_sum, rtol=1e-12, atol=1e-12)
        if t % 10 == 9 or t == T - 1:
            e = env.export_state()
            for b in rng.integers(0, B, 25):
                ob = orc.export(int(b))
                for k in STATE_KEYS:
                    assert np.array_equal(np.asarray(e[k][b]), np.asarray(ob[k])), (key, t, b, k)
                assert np.array_equal(e["chain"][b], ob["chain"]), (key, t, b, "chain")
                assert e["job_counter"][b] == ob["job_counter"]
            co = {k: v.cpu().numpy() for k, v in env.observe_compact().items()}
            for b in rng.integers(0, B, 10):   # the compact observation record against the oracle's state
                oc = O.compact_observation(orc.export(int(b)))
                for k in oc:
                    assert np.array_equal(co[k][b].reshape(oc[k].shape), oc[k]), (key, t, b, k)
            if key == "cfg5":
                continue  # dense observations are 2.2 MB per env there; the compact record is what config 5 uses
            of = {k: v.cpu().numpy() for k, v in env.obs_views().items()} if fused else None
            o = {k: v.cpu().numpy() for k, v in env.observe(with_ids=True).items()}
            if fused:
                for k in of:
                    assert np.array_equal(of[k], o[k]), (key, t, k)
            for b in rng.integers(0, B, 10):
                oo = orc.observe(int(b))
                assert np.array_equal(o["acceptor"][b], oo["obs_acc"])
                assert np.array_equal(o["offer"][b], oo["obs_off"])
                assert np.array_equal(o["auctioneer"][b], oo["obs_auc"])
                assert np.array_equal(o["ids"][b], oo["ids"])
                assert np.array_equal(o["auctioneer_ids"][b], oo["auc_ids"])
    assert int(orc.n_accepted.sum()) >= 0
    env.close()


def test_fault_flags():
    """Out-of-range acceptor index -> ACTION_RANGE flag (reference: AssertionError);
    chain capacity 1 -> CHAIN_OVERFLOW flag."""
    dom, mode = DOMS["cfg2"]
    env = _env(4, dict(dom, mode=mode), auction="first", spawn="philox", chain_capacity=1)
    N, C, L = dom["N"], dom["C"], dom["L"]
    acc = np.zeros((4, N, C), np.int64)
    acc[1, 0, 0] = N * L + 1
    acc[2, 1, 2] = -1
    r = env.step(np.zeros((4, N, L), np.int64), acc, None)
    f = r["flags"].cpu().numpy()
    assert f[0] == 0 and f[1] & 4 and f[2] & 4 and f[3] == 0
    for t in range(40):
        r = env.step(np.zeros((4, N, L), np.int64), np.zeros((4, N, C), np.int64), None)
    assert (r["flags"].cpu().numpy()[0] & 1) == 1
    env.close()


@pytest.mark.parametrize("B,pinned", [(500, True), (512, True), (1024, False), (384, True)])
def test_step_host_matches_device_step(B, pinned):
    """msched_step_host: the staged path (padded batch or pageable host memory: chunked H2D / kernel / D2H) and the
    zero-copy path (pinned buffers, no padding: the kernel's bulk copies read and write host memory directly)."""
    import torch
    dom, mode = DOMS["cfg3"]
    a = _env(B, dict(dom, mode=mode), auction="first", spawn="philox", seed=3)
    b = _env(B, dict(dom, mode=mode), auction="first", spawn="philox", seed=3)
    rng = np.random.default_rng(0)
    lay = a.layout
    ah = torch.zeros((B, lay.action_halfs), dtype=torch.int16)
    rh = torch.zeros((B, lay.result_words), dtype=torch.int32)
    if pinned:
        ah, rh = ah.pin_memory(), rh.pin_memory()
    for t in range(20):
        offc, acc, offp = random_actions(rng, B, dom, True)
        a.step(offc, acc, None, offer_price=offp)
        ah.copy_(a.action[:B].cpu())
        b.step_host(ah, rh, observe=(t % 2 == 0))
        assert torch.equal(rh, a.result[:B].cpu())
        if t % 2 == 0:
            oa, ob = a.observe(), b.obs_views()
            for k in oa:
                assert torch.equal(oa[k], ob[k]), (t, k)
    assert a.round == b.round == 20
    a.close(); b.close()


@pytest.mark.parametrize("key", ["cfg3", "cfg2", "agg", "cfg1"])
def test_step_host_compact_record_equals_full_record(key):
    """msched_step_host_compact: the same step, the result in int16 / half planes (56 B instead of 116 B per env in
    config 3).  Every field decodes to exactly what the full record holds (quality_sum to float32); priorities
    beyond the exact half range or a netZeroOfferReward that is no half are refused."""
    import torch
    from marl_scheduling_b200._lib import MschedError
    dom, mode = DOMS[key]
    free = mode.startswith("free")
    B = 384
    a = _env(B, dict(dom, mode=mode), auction="random", spawn="philox", seed=3)
    b = _env(B, dict(dom, mode=mode), auction="random", spawn="philox", seed=3)
    cl = b.compact_result_layout()
    lay = a.layout
    assert cl.c_flags == cl.c_counts and cl.words * 4 < lay.result_words * 4   # flags ride in the counts word
    ah = torch.zeros((B, lay.action_halfs), dtype=torch.int16).pin_memory()
    ch = torch.zeros((B, cl.words), dtype=torch.int32).pin_memory()
    rng = np.random.default_rng(0)
    seen = 0
    for t in range(40):
        offc, acc, offp = random_actions(rng, B, dom, free)
        ra = a.step(offc, acc, None, offer_price=offp)
        ah.copy_(a.action[:B].cpu())
        b.step_host_compact(ah, ch, observe=(t % 3 == 0))
        rb = b.compact_rewards(ch)
        for k in ("offer", "price", "acceptor", "auctioneer", "agent", "quality_cnt", "n_accepted", "n_terminated", "done", "flags"):
            if ra[k] is None:
                assert rb[k] is None
                continue
            assert torch.equal(ra[k].cpu().to(torch.float64), rb[k].to(torch.float64)), (key, t, k)
        assert torch.allclose(ra["quality_sum"].cpu().float(), rb["quality_sum"], rtol=1e-6, atol=1e-6)
        seen += int((ra["acceptor"] != 0).sum()) + int((ra["offer"] != 0).sum())
        if t % 3 == 0:
            oa, ob = a.observe(), b.obs_views()
            for k in oa:
                assert torch.equal(oa[k], ob[k]), (t, k)
    assert seen > 100 and a.round == b.round == 40
    ea, eb = a.export_state(), b.export_state()
    for k in STATE_KEYS:
        assert np.array_equal(ea[k], eb[k]), k
    a.close(); b.close()
    with pytest.raises(MschedError):
        _env(128, dict(dom, mode=mode, prios=[5000] * len(dom["prios"]))).compact_result_layout()
    if free:
        with pytest.raises(MschedError):
            _env(128, dict(dom, mode=mode, netZero=0.3)).compact_result_layout()


@pytest.mark.parametrize("full_record", [False, True])
def test_step_host_full_batch_in_waves_matches_device_step(full_record):
    """The host-buffer step at BASELINE's 65,536 environments: the zero-copy launch is held to a few resident CTAs per SM
    and runs its 2,048 tiles in waves (csrc/msched_abi.cu launch_step) -- results, states and observations equal the
    device step's, for the compact and for the full result record."""
    import torch
    dom, mode = DOMS["cfg3"]
    B = 65536
    a = _env(B, dict(dom, mode=mode), auction="random", spawn="philox", seed=8)
    b = _env(B, dict(dom, mode=mode), auction="random", spawn="philox", seed=8)
    lay = a.layout
    cl = b.compact_result_layout()
    ah = torch.zeros((B, lay.action_halfs), dtype=torch.int16).pin_memory()
    rh = torch.zeros((B, lay.result_words if full_record else cl.words), dtype=torch.int32).pin_memory()
    rng = np.random.default_rng(4)
    for t in range(6):
        offc, acc, offp = random_actions(rng, B, dom, True)
        ra = a.step(offc, acc, None, offer_price=offp)
        ah.copy_(a.action[:B].cpu())
        if full_record:
            b.step_host(ah, rh, observe=True)
            assert torch.equal(rh.to(a.result.device), a.result[:B]), t
        else:
            b.step_host_compact(ah, rh, observe=True)
            rb = b.compact_rewards(rh)
            for k in ("offer", "price", "acceptor", "auctioneer", "agent", "n_accepted", "n_terminated", "done", "flags"):
                assert torch.equal(ra[k].cpu().to(torch.float64), rb[k].to(torch.float64)), (t, k)
        oa, ob = a.observe(), b.obs_views()
        for k in oa:
            assert torch.equal(oa[k], ob[k]), (t, k)
    assert torch.equal(a.state[:B], b.state[:B]) and torch.equal(a.chain[:B], b.chain[:B])
    a.close(); b.close()


@pytest.mark.parametrize("impl", ["fused1", "fused4", "lane"])
def test_cuda_graph_replay_matches_eager_steps(impl):
    """Device-side round counter (msched_set_round_mode): a captured step + observations can be
    replayed from a CUDA graph and walks through the same states as eager launches (the round feeds
    the Philox spawn draws, birth dates, chain rounds and the done flag)."""
    import torch
    dom, mode = DOMS["cfg3"]
    B, T = 700, 45
    a = _env(B, dict(dom, mode=mode), impl=impl, auction="random", spawn="philox", seed=11)
    b = _env(B, dict(dom, mode=mode), impl=impl, auction="random", spawn="philox", seed=11)
    b.set_device_round(True)
    rng = np.random.default_rng(2)
    offc, acc, offp = random_actions(rng, B, dom, True)
    for e in (a, b):
        e.set_actions(offc, acc, None, offer_price=offp)
    # warm-up launch outside the capture, mirrored eagerly
    a.step_observe_records()
    b.step_observe_records()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            b.step_observe_records()
    torch.cuda.current_stream().wait_stream(s)
    a.step_observe_records()          # the capture itself did not run anything: replay once for step 2
    g.replay()
    for t in range(T):
        a.step_observe_records()
        g.replay()
        if t % 11 == 0 or t == T - 1:
            torch.cuda.synchronize()
            assert torch.equal(a.state[:B], b.state[:B]), (impl, t)
            assert torch.equal(a.result[:B], b.result[:B]), (impl, t)
            assert torch.equal(a._obs[:B], b._obs[:B]), (impl, t)
    assert a.round == b.round == T + 2
    done = (b.result[:B, b.layout.r_counts] >> 24) & 1
    assert int(done.max()) == int(((T + 2) % 100) == 0)
    a.close(); b.close()


EDGE = {
    # one environment, one agent, one core, one slot
    "tiny": (dict(N=1, C=1, L=1, prios=[3], lens=[2], probs=[1], fix=[2]), "fix", 1),
    # 16 job kinds (the maximum), every agent refills its whole collection per round
    "kinds16": (dict(N=2, C=3, L=3, prios=list(range(1, 17)), lens=[1 + (k % 7) for k in range(16)],
                     probs=[1 / 16] * 16, fix=list(range(16)), newJobs=3), "fix", 257),
    # N*L = 254: the largest offer table of the lane-per-env kernels; job length 255; price -2 / -1 sentinels
    "wide": (dict(N=127, C=2, L=2, prios=[5, 9], lens=[255, 1], probs=[0.5, 0.5], fix=[-2, -1]), "fix", 40),
    # 64 cores (the maximum), cooperative kernel only
    "cores64": (dict(N=3, C=64, L=4, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7], mult=2), "agg", 70),
    # free prices incl. the -5 of quirk Q1 and negative prices (payments flip direction, Q2)
    "negprice": (dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1]), "free_ncomm", 300),
}


@pytest.mark.parametrize("key", list(EDGE))
def test_edge_domains_match_oracle(key):
    """Edge cases: single env, maximum kinds / table width / core count, sentinel and negative prices,
    batches that are not a multiple of the 32-env tile."""
    from oracle import oracle as O
    dom, mode, B = EDGE[key]
    free = mode.startswith("free")
    impls = ["coop", "warp"] if key == "cores64" else (["lane", "coop", "warp"] + (["fused1", "fused2"] if _has_fused(dom) else []))
    from marl_scheduling_b200 import MschedError
    ran = 0
    for impl in impls:
        try:
            env = _env(B, dict(dom, mode=mode), impl=impl, auction="random", spawn="philox", seed=77, chain_capacity=255)
        except MschedError as e:   # e.g. 32 records of the widest domain exceed the lane kernel's shared memory
            assert "not available for this domain" in str(e), e
            continue
        ran += 1
        orc = O.Oracle(B, dom, mode, chain_cap=255, tie_mode=O.TIE_PHILOX, seed=77)
        rng = np.random.default_rng(3)
        N, C, L = dom["N"], dom["C"], dom["L"]
        for t in range(30):
            offc, acc, offp = random_actions(rng, B, dom, free)
            if key == "negprice":
                offp = rng.integers(-6, 9, (B, N, L))
            if key == "wide":   # make acceptances likely despite the 255-entry tables
                acc = rng.integers(0, 3, (B, N, C))
            r = env.step(offc, acc, None, offer_price=offp)
            orc.step(offc, acc, None, offp=offp)
            r = {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}
            assert np.array_equal(r["auctioneer_idx"], orc.auc_out), (key, impl, t)
            assert np.array_equal(r["offer"].astype(np.float64), orc.r_offer), (key, impl, t)
            assert np.array_equal(r["acceptor"], orc.r_acceptor), (key, impl, t)
            assert np.array_equal(r["auctioneer"], orc.r_auctioneer), (key, impl, t)
            assert np.array_equal(r["agent"], orc.r_agent), (key, impl, t)
            assert np.array_equal(r["flags"].astype(np.uint32), orc.flags), (key, impl, t)
            if free:
                assert np.array_equal(r["price"].astype(np.float64), orc.r_price), (key, impl, t)
        e = env.export_state()
        for b in sorted({0, B // 2, B - 1}):
            ob = orc.export(b)
            for k in STATE_KEYS:
                assert np.array_equal(np.asarray(e[k][b]), np.asarray(ob[k])), (key, impl, b, k)
            assert np.array_equal(e["chain"][b], ob["chain"]), (key, impl, b)
        if key != "cores64" and key != "wide":
            o = {k: v.cpu().numpy() for k, v in env.observe().items()}
            oo = orc.observe(B - 1)
            assert np.array_equal(o["acceptor"][B - 1], oo["obs_acc"])
            assert np.array_equal(o["offer"][B - 1], oo["obs_off"])
        env.close()
    assert ran >= 1


@pytest.mark.parametrize("key,B", [("cfg3", 65536), ("cfg2", 65536)])
def test_full_size_batch_windows_match_oracle_and_jobs_are_conserved(key, B):
    """BASELINE sizes (65,536 envs per GPU, the kernels the bench launches): (1) because every device draw is
    keyed on the GLOBAL env index, any window of the big batch must equal an oracle instance started at that env
    offset -- three 64-env windows (first tile, middle, last tile) are compared bit-exactly every step;
    (2) size-independent invariants over ALL envs: jobs spawned = jobs terminated + jobs present, at most C
    acceptances per step, no sticky fault flags."""
    import torch
    from oracle import oracle as O
    dom, mode = DOMS[key]
    free = mode.startswith("free")
    N, C, L = dom["N"], dom["C"], dom["L"]
    T, W, seed, base = 48, 64, 99, 1 << 20
    env = _env(B, dict(dom, mode=mode), auction="random", spawn="philox", seed=seed, env_offset=base)
    assert env.info()["step_impl"] == "fused"
    wins = [0, 30016, B - W]
    orcs = [O.Oracle(W, dom, mode, tie_mode=O.TIE_PHILOX, seed=seed, env_offset=base + w0) for w0 in wins]
    rng = np.random.default_rng(17)
    terminated = torch.zeros(B, dtype=torch.int64, device=env.device)
    for t in range(T):
        offc, acc, offp = random_actions(rng, B, dom, free)
        r = env.step(offc, acc, None, offer_price=offp, observe=True)
        terminated += r["n_terminated"].long()
        assert int(r["n_accepted"].max()) <= C and int(r["flags"].max()) == 0, t
        for w0, orc in zip(wins, orcs):
            sl = slice(w0, w0 + W)
            orc.step(offc[sl], acc[sl], None, offp=None if offp is None else offp[sl])
            assert np.array_equal(r["auctioneer_idx"][sl].cpu().numpy(), orc.auc_out), (t, w0)
            assert np.array_equal(r["agent"][sl].cpu().numpy(), orc.r_agent), (t, w0)
            assert np.array_equal(r["acceptor"][sl].cpu().numpy(), orc.r_acceptor), (t, w0)
            assert np.array_equal(r["auctioneer"][sl].cpu().numpy(), orc.r_auctioneer), (t, w0)
            assert np.array_equal(r["offer"][sl].cpu().numpy().astype(np.float64), orc.r_offer), (t, w0)
    e = env.export_state()
    obs = {k: v.cpu().numpy() for k, v in env.obs_views().items()}
    for w0, orc in zip(wins, orcs):
        for b in (0, W // 2, W - 1):
            ob = orc.export(b)
            for k in STATE_KEYS:
                assert np.array_equal(np.asarray(e[k][w0 + b]), np.asarray(ob[k])), (w0, b, k)
            assert np.array_equal(e["chain"][w0 + b], ob["chain"]), (w0, b)
            oo = orc.observe(b)
            assert np.array_equal(obs["acceptor"][w0 + b], oo["obs_acc"])
            assert np.array_equal(obs["offer"][w0 + b], oo["obs_off"])
            assert np.array_equal(obs["auctioneer"][w0 + b], oo["obs_auc"])
    # conservation over all 65,536 envs: every job ever created (IDs 1 .. counter-1) is on a core, in a slot, or done
    present = ((np.asarray(e["core_jobid"]).reshape(B, -1) > 0).sum(1)
               + (np.asarray(e["slot_jobid"]).reshape(B, -1) > 0).sum(1))
    spawned = np.asarray(e["job_counter"]).astype(np.int64) - 1
    assert np.array_equal(spawned, terminated.cpu().numpy() + present)
    assert spawned.min() > 0 and len(np.unique(spawned)) > 3      # the envs really diverged
    env.close()


def test_config5_full_size_windows_match_oracle_and_jobs_are_conserved():
    """BASELINE configs[4] at its full size (N32 C64 L8, 262,144 envs, the kernel the library picks for it): the
    size-independent properties of test_full_size_batch_windows_... -- (1) three windows of the big batch equal
    oracle instances started at the same global env offsets, step by step (auction winners, every reward plane,
    counts) and in their final state (cores, slots, offers, chains); (2) over ALL envs: jobs spawned = jobs
    terminated + jobs present, at most C acceptances per step, no fault flags.  The action records are drawn on the
    device (a host copy of one step's actions would be 4 GB)."""
    import torch
    from oracle import oracle as O
    dom, mode = DOMS["cfg5"]
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL = N * L
    B, T, W, seed, base = 262144, 14, 8, 7, 1 << 22
    env = _env(B, dict(dom, mode=mode), auction="random", spawn="philox", seed=seed, env_offset=base, chain_capacity=16)
    lay = env.layout
    wins = [0, 131072 + 40, B - W]
    orcs = [O.Oracle(W, dom, mode, chain_cap=16, tie_mode=O.TIE_PHILOX, seed=seed, env_offset=base + w0) for w0 in wins]
    gen = torch.Generator(device=env.device).manual_seed(3)
    terminated = torch.zeros(B, dtype=torch.int64, device=env.device)
    acc_v, off_v = env.acceptor_actions, env.offer_core_actions
    for t in range(T):
        # acceptor indices biased towards the first table entries so that agents really accept offers
        acc_v.random_(0, NL + 1, generator=gen)
        low = torch.rand((B, N, C), device=env.device, generator=gen) < 0.6
        acc_v[low] = 0
        off_v.random_(0, C + 1, generator=gen)
        env.step_compact_records()   # ONE launch: transition + compact observations of the new state
        r = env.rewards()
        terminated += r["n_terminated"].long()
        assert int(r["n_accepted"].max()) <= C and int(r["flags"].max()) == 0, t
        for w0, orc in zip(wins, orcs):
            sl = slice(w0, w0 + W)
            orc.step(off_v[sl].cpu().numpy(), acc_v[sl].cpu().numpy(), None)
            assert np.array_equal(r["auctioneer_idx"][sl].cpu().numpy(), orc.auc_out), (t, w0)
            assert np.array_equal(r["agent"][sl].cpu().numpy(), orc.r_agent), (t, w0)
            assert np.array_equal(r["acceptor"][sl].cpu().numpy(), orc.r_acceptor), (t, w0)
            assert np.array_equal(r["auctioneer"][sl].cpu().numpy(), orc.r_auctioneer), (t, w0)
            assert np.array_equal(r["offer"][sl].cpu().numpy().astype(np.float64), orc.r_offer), (t, w0)
            assert np.array_equal(r["n_accepted"][sl].cpu().numpy(), orc.n_accepted), (t, w0)
            assert np.array_equal(r["n_terminated"][sl].cpu().numpy(), orc.n_term), (t, w0)
    assert env.info()["step_impl"] == "warp"
    n_acc = 0
    co = {k: v for k, v in env.compact_views().items()}
    for w0, orc in zip(wins, orcs):
        e = env.export_state(w0, W)
        for b in range(W):
            ob = orc.export(b)
            oc = O.compact_observation(ob)
            for k in oc:   # the compact observations the step launch wrote
                assert np.array_equal(co[k][w0 + b].cpu().numpy().reshape(oc[k].shape), oc[k]), (w0, b, k)
            for k in STATE_KEYS:
                assert np.array_equal(np.asarray(e[k][b]), np.asarray(ob[k])), (w0, b, k)
            assert np.array_equal(e["chain"][b], ob["chain"]), (w0, b)
            n_acc += int(ob["chain_len"].sum())
    assert n_acc > 20  # liability chains are in play
    # conservation over all 262,144 envs, read from the state records (csrc/msched_common.cuh: word 0 = next jobID,
    # core c jobID at 2+3c+1, slot s jobID at S_SLOT+4s+1; -1 = empty)
    st = env.state[:B]
    s_slot = 2 + 3 * C + (C + 3) // 4
    present = (st[:, 3:2 + 3 * C:3] > 0).sum(1) + (st[:, s_slot + 1::4][:, :NL] > 0).sum(1)
    spawned = st[:, 0].long() - 1
    assert torch.equal(spawned, terminated + present)
    assert int(spawned.min()) > 0 and int(terminated.sum()) > B
    env.close()


@pytest.mark.parametrize("key,auction,obs_every,dev_round,B", [
    ("cfg3", "random", False, False, 1000), ("cfg3", "random", True, True, 1000), ("cfg3", "first", True, False, 1000),
    ("cfg2", "random", True, False, 1000), ("cfg2", "random", False, True, 1000), ("cfg1", "random", False, False, 1000),
    ("cfg3", "random", True, True, 65536), ("cfg3", "random", False, False, 65536)])  # BASELINE's full batch
def test_step_multi_matches_single_steps(key, auction, obs_every, dev_round, B):
    """msched_step_multi (T steps per launch, a CTA keeps its 32 environments; without obs_every the state tile stays
    in shared memory between steps) walks bit for bit through the states, result records and observations of T
    msched_step_observe calls -- host round and device round counter, two launches in a row."""
    import torch
    dom, mode = DOMS[key]
    free = mode.startswith("free")
    T = 9
    a = _env(B, dict(dom, mode=mode), auction=auction, spawn="philox", seed=21)
    b = _env(B, dict(dom, mode=mode), auction=auction, spawn="philox", seed=21)
    if dev_round:
        a.set_device_round(True); b.set_device_round(True)
    lay, dev = a.layout, a.device
    rng = np.random.default_rng(6)
    for launch in range(2):
        acts = torch.zeros((T, lay.padded_envs, lay.action_halfs), dtype=torch.int16, device=dev)
        for t in range(T):
            offc, acc, offp = random_actions(rng, B, dom, free)
            a.set_actions(offc, acc, None, offer_price=offp)
            acts[t].copy_(a.action)
        res1 = torch.zeros((T, lay.padded_envs, lay.result_words), dtype=torch.int32, device=dev)
        obs1 = torch.zeros((T, lay.padded_envs, lay.obs_halfs), dtype=torch.int16, device=dev)
        for t in range(T):
            a.step_observe_records(acts[t], res1[t], obs=obs1[t])
        res2 = torch.zeros_like(res1)
        obs2 = torch.zeros_like(obs1) if obs_every else torch.zeros_like(obs1[0])
        b.step_multi_records(acts, res2, obs2, obs_every=obs_every)
        torch.cuda.synchronize()
        assert torch.equal(res1[:, :B], res2[:, :B]), (key, launch)
        assert torch.equal(a.state[:B], b.state[:B]) and torch.equal(a.chain[:B], b.chain[:B])
        if obs_every:
            assert torch.equal(obs1[:, :B], obs2[:, :B])
        else:
            assert torch.equal(obs1[T - 1, :B], obs2[:B])
        assert a.round == b.round == (launch + 1) * T
    assert int(res2[:, :B, lay.r_flags].max()) == 0
    a.close(); b.close()
