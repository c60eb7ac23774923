#!/usr/bin/env python
"""bench.py -- headline benchmark: agent-steps/sec of the batched env step + auction.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3|cfg2]

One "step" = one pass of the hot path over one batch of environments: ONE launch of the fused
kernel (acceptances, in-kernel auction + core allocation, progress/completion, offer creation,
Philox spawn, rewards, dense observations of the new state), i.e. one SchedulingEnv.step for
every environment.  Workload at N=1: BASELINE.json configs[2] ("cfg3": N=2 C=3 L=3, three job kinds,
free prices, commercial reward, hard-coded auctioneer) at 65,536 environments -- the
configuration the metric is quoted on.  Each extra GPU adds its own 65,536-env shard (weak
scaling, no data-path collective).

Timing: inputs larger than L2 -- several independent 65,536-env shards are visited round-robin, so
a shard's records come from HBM; launches are back to back on the launching stream, timed with
CUDA events in blocks (fresh action records are drawn, untimed, between blocks).  `--l2 flush` is
the per-launch protocol (256 MiB write before every step).  `value` = units processed / sum of
the block times (max over ranks).  `e2e` goes through msched_step_host: pinned host action
records -> H2D -> step -> D2H of the result records -> sync, wall-clock timed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[2]; reference src/trainPPOExperiment4-2.py:42-56,103
    "cfg3": dict(dom=dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1],
                          mult=1, newJobs=1, episodeLength=100, netZero=0.5),
                 mode="free_comm", envs=65536,
                 desc="N2 C3 L3, 3 job kinds, free prices, commercial reward, hard-coded auctioneer"),
    # BASELINE.json configs[1]; reference README.md:49-61, src/trainPPOCopy.py:41-54
    "cfg2": dict(dom=dict(N=4, C=4, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7],
                          mult=1, newJobs=1, episodeLength=100),
                 mode="fix", envs=4096,
                 desc="N4 C4 L3, 2 job kinds, fixed prices, divided reward, hard-coded auctioneer"),
    "cfg4": dict(dom=dict(N=4, C=4, L=3, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7],
                          mult=1, newJobs=1, episodeLength=100),
                 mode="fix", envs=65536,
                 desc="N4 C4 L3 (cfg2 domain) at 65,536 envs per GPU"),
    # BASELINE.json configs[0]: src/trainHC.py:19-30 domain, hard-coded agents (the reference's own CPU-runnable case)
    "cfg1": dict(dom=dict(N=2, C=3, L=2, prios=[5], lens=[4], probs=[1], fix=[3], mult=2, newJobs=1, episodeLength=50),
                 mode="fix", envs=65536, policy="hardcoded",
                 desc="N2 C3 L2, 1 job kind, fixed prices, hard-coded agents (trainHC)"),
    # BASELINE.json configs[4]: large domain, cooperative kernel (one warp per env), no dense obs
    "cfg5": dict(dom=dict(N=32, C=64, L=8, prios=[3, 10], lens=[6, 3], probs=[0.8, 0.2], fix=[2, 7],
                          mult=1, newJobs=1, episodeLength=100),
                 mode="fix", envs=262144, obs="compact", ring=2,
                 desc="N32 C64 L8, 2 job kinds, fixed prices, 262,144 envs, compact observations (state record)"),
}


def algorithmic_bytes(dom, mode):
    """canonical bytes per env-step (state read+write, actions, rewards)
    and of the dense observation record."""
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL = N * L
    free = mode.startswith("free")
    nj = dom.get("newJobs", 1)
    S = 16 * C + 20 * NL + 4 * C + 8
    w = 1 if NL + 1 <= 255 else 2
    a = N * C * w + NL * (2 if free else 1) + C * w + N * nj
    r = 4 * N * C + 4 * NL * (2 if free else 1) + 4 * C + 4 * N
    o = 2 * (N * C * (3 + 2 * NL) + NL * (2 * C + 2) + C * (3 + 2 * NL))
    oc = 2 * (2 * C + 2 * NL + 3 * NL)  # compact observations
    return dict(step=2 * S + a + r, obs=o, obs_compact=oc, state=S)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profile(cfg_name, kernel="step_kernel"):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture,
    or None if that kernel has not been captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(cfg_name, {}).get(kernel + "_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
             0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
             0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def sample(self):
        nv = self.nv
        mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.samples.append(mhz)
        for bit, name in self.NAMES.items():
            if mask & bit and name != "gpu_idle":
                self.reasons.add(name)

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.sample()
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.ok and not self.samples:
            try:
                self.sample()
            except Exception:
                pass
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "window": "state warm-up + warm-up + timed steps (the same launches throughout)"}


def refresh_actions(env, rec, gen):
    """Uniform random action record: acceptor idx ~ U{0..NL}, offer core ~ U{0..C},
    price ~ U{0..maxPrio}.  Fresh draws every step: a short periodic ring of records lets some
    environments fall into resonant cycles whose liability chains grow without bound."""
    lay, B = env.layout, env.B
    P = max(env.cfg.prio[k] for k in range(env.cfg.J))
    rec[:B, lay.a_acceptor: lay.a_acceptor + env.N * env.C].random_(0, env.NL + 1, generator=gen)
    rec[:B, lay.a_offer_core: lay.a_offer_core + env.NL].random_(0, env.C + 1, generator=gen)
    if lay.a_offer_price >= 0:
        rec[:B, lay.a_offer_price: lay.a_offer_price + env.NL].random_(0, P + 1, generator=gen)


def make_actions(torch, env, ring, seed):
    gen = torch.Generator(device=env.device).manual_seed(seed)
    recs = []
    for _ in range(ring):
        rec = torch.zeros((env.layout.padded_envs, env.layout.action_halfs), dtype=torch.int16,
                          device=env.device)
        refresh_actions(env, rec, gen)
        recs.append(rec)
    return recs, gen


def cpu_baseline_sample(cfg, seconds=12.0, threads=1, envs=2048):
    """Times the CPU port (oracle/) on a bounded sample of the same workload.  Checker code used
    as the reported baseline only; see DESIGN.md section "CPU baseline"."""
    import numpy as np
    from oracle import oracle as O
    dom, mode = cfg["dom"], cfg["mode"]
    free = mode.startswith("free")
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL, P = N * L, max(dom["prios"])
    B = envs * threads
    orc = O.Oracle(B, dom, mode, tie_mode=O.TIE_PHILOX, seed=0)
    rng = np.random.default_rng(0)
    ring = [(rng.integers(0, C + 1, (B, N, L)).astype(np.int32),
             rng.integers(0, NL + 1, (B, N, C)).astype(np.int32),
             rng.integers(0, P + 1, (B, N, L)).astype(np.int32) if free else None) for _ in range(4)]

    def run(steps):
        if threads == 1:
            for s in range(steps):
                offc, acc, offp = ring[s % 4]
                orc.step(offc, acc, None, offp=offp)
            return
        bounds = [(t * envs, (t + 1) * envs) for t in range(threads)]

        def work(b0, b1):
            for s in range(steps):
                offc, acc, offp = ring[s % 4]
                orc.step(offc, acc, None, offp=offp, b0=b0, b1=b1)
        ths = [threading.Thread(target=work, args=b) for b in bounds]
        [t.start() for t in ths]
        [t.join() for t in ths]

    run(100)  # warm: steady-state occupancy
    steps_done, t0 = 0, time.perf_counter()
    chunk = 50
    while True:
        run(chunk)
        steps_done += chunk
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return dict(value=B * N * steps_done / el, env_steps=B * steps_done, seconds=el, envs=B,
                steps=steps_done)


def python_reference_sample(cfg, envs=64, steps=20, warm_steps=200, seconds=None, seed=0):
    """Times the UNMODIFIED Python reference (oracle/_ref, or /root/reference in the build container):
    `envs` independent reference worlds, each advanced by SchedulingEnv.step
    (src/SchedulingEnvironment.py:32-83) + Auctioneer.getAuctioneerAction (src/Auctioneer.py:95-102) per
    step, one process, one thread -- the reference has no batched or multi-core mode.  Actions are the
    bench workload's (uniform random indices, drawn untimed).  `warm_steps` untimed steps bring every
    world to steady state; then `steps` timed steps (or, with `seconds`, whole steps until that much
    time has passed).  Returns None when the reference files are absent."""
    from oracle import ref_harness as RH
    if not RH.reference_available():
        return None
    import numpy as np
    import torch
    torch.set_num_threads(1)
    dom, mode = cfg["dom"], cfg["mode"]
    free = mode.startswith("free")
    N, C, L = dom["N"], dom["C"], dom["L"]
    NL, P = N * L, max(dom["prios"])
    rng = np.random.default_rng(seed)
    worlds = [RH.make_env(dom, mode) for _ in range(envs)]
    obs = [env.reset() for _, env in worlds]

    def draw():
        acc = rng.integers(0, NL + 1, (envs, N, C)).tolist()
        offc = rng.integers(0, C + 1, (envs, N, L)).tolist()
        if free:
            offp = rng.integers(0, P + 1, (envs, N, L)).tolist()
            off = [[[(offc[e][i][q], offp[e][i][q]) for q in range(L)] for i in range(N)] for e in range(envs)]
        else:
            off = offc
        return acc, off

    def one_step(acts):
        acc, off = acts
        for e, (world, env) in enumerate(worlds):
            auc = world.auctioneer.getAuctioneerAction(obs[e][2])
            out = env.step(off[e], acc[e], auc)
            obs[e] = out[:3]

    for _ in range(warm_steps):
        one_step(draw())
    done, el = 0, 0.0
    while True:
        acts = draw()
        t0 = time.perf_counter()
        one_step(acts)
        el += time.perf_counter() - t0
        done += 1
        if (seconds is None and done >= steps) or (seconds is not None and el >= seconds):
            break
    return dict(value=envs * N * done / el, env_steps=envs * done, seconds=el, envs=envs, steps=done,
                warm_steps=warm_steps, source=RH.REFERENCE_SRC)


def l2_rule(args):
    """How the GPU arm keeps its inputs out of L2 between timed iterations -- a function of the arguments alone, so
    that both arms print the same `config` (the reference arm runs on the GPU arm's config; it has no L2 to flush)."""
    if args.no_flush:
        return "warm (no flush)"
    if args.l2 == "flush":
        return "flushed before every timed step (256 MiB write), per-launch events"
    return ("inputs larger than L2: independent shards of the batch (>= 200 MB of records in flight) visited round-robin, "
            f"launches in blocks on {args.streams} stream(s)" + (", each block one CUDA-graph replay" if args.graph == 1 else "") +
            ", every timed block enqueued behind a 150 us spin kernel (host enqueue latency outside the events); "
            "`--l2 auto` falls back to the flush protocol for domains whose step is two launches")


def workload_config(args, cfg, B):
    """The workload description both arms print (same keys and values, so the two lines name the same work)."""
    return {"workload": args.config, "domain": cfg["desc"], "envs_per_gpu": B,
            "observations": "none" if cfg.get("obs") == "none" else args.obs,
            "auctioneer": "hard-coded auction rule, random arg-max tie-break",
            "spawn": "typed spawn distribution, one draw per agent and round",
            "actions": "uniform random indices, fresh draws every step (untimed)",
            "l2": l2_rule(args)}


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU implementation of the path -- the unmodified Python
    SchedulingEnv.step + getAuctioneerAction from oracle/_ref (one process, one thread: all the host
    threads it can use) -- on a bounded sample of the workload per step.  The C port of the same path on
    every host core at the full 65,536 envs per step is reported beside it as `cpu_baseline_port`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.envs or cfg["envs"]
    K, W = max(1, args.steps), max(0, args.warmup)
    N = cfg["dom"]["N"]
    threads = os.cpu_count() or 1
    # per step: a sample of `envs` reference worlds (the reference advances ~1.8 k env-steps/s per core)
    envs = 128 if N * cfg["dom"]["L"] <= 16 else 2
    warm = max(W, 200 if envs > 2 else 5)
    ref = python_reference_sample(cfg, envs=envs, steps=K, warm_steps=warm)
    port = cpu_baseline_sample(cfg, seconds=5.0, threads=threads, envs=max(64, B // threads)) \
        if cfg["dom"]["N"] * cfg["dom"]["L"] <= 64 else None
    if ref is not None:
        res, kind, cores = ref, "reference", 1
        sample = (f"{ref['envs']} independent reference worlds x {ref['steps']} timed steps after {ref['warm_steps']} "
                  f"untimed, SchedulingEnv.step + Auctioneer.getAuctioneerAction of the unmodified Python reference "
                  f"({os.path.relpath(ref['source'], ROOT) if ref['source'].startswith(ROOT) else ref['source']}), 1 process, 1 thread of {threads} host cores (the reference is single-threaded)")
    else:  # the reference files did not travel: the C port stands in
        res, kind, cores = port, "port", threads
        sample = f"{port['envs']} envs x {port['steps']} steps, {threads} threads, oracle/msched_oracle.c"
    line = {
        "impl": "reference", "metric": "agent-steps/sec, batched env step+auction @65,536 envs",
        "value": res["value"], "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / res["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": workload_config(args, cfg, B),
        "sample_envs_per_step": res["envs"],
        "cpu_baseline": {"value": res["value"], "unit": "agent-steps/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "cpu_baseline_port": None if port is None else {
            "value": port["value"], "unit": "agent-steps/s", "cores": threads, "kind": "port",
            "sample": f"{port['envs']} envs per step x {port['steps']} steps, {threads} threads, "
                      "oracle/msched_oracle.c (C restatement of the same path)"},
        "e2e": {"value": res["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def ppo_update_dp_leg(torch, dist, dev, world, rank, T=8, B=65536, epochs=3, reps=3):
    """BASELINE configs[3] ("cfg4"): PPO with globally shared parameters, env shards on every GPU, NCCL gradient
    all-reduce (GloballySharedPPO.update, src/PPOmodules.py:395-449; sampling rule src/SchedulingEnvironment.py:
    314-329).  The cfg2/cfg4 domain's shared acceptor net (27->16->16->13 + critic) learns from one sampled unit's
    buffer of T steps x B envs PER RANK: every epoch = msched_ppo_grad on the rank's shard -> ONE flat all_reduce
    (both heads) -> msched_adam_step, through agents.BatchedPPO.update (the product path).  Timed on the device with
    CUDA events, max over ranks.  The in-bench check (the driver's pytest box has one GPU): on a small buffer every
    rank computes the gradient of its shard, the all-reduced mean must equal the single-rank gradient of the WHOLE
    buffer, and after the K Adam steps all ranks must hold bit-identical weights."""
    from marl_scheduling_b200 import policy
    from marl_scheduling_b200.agents import BatchedPPO
    N, C, L = 4, 4, 3
    NL, U = N * L, N * C
    n_in, A = 3 + 2 * NL, NL + 1
    kw = dict(lr_actor=3e-3, lr_critic=1e-2, gamma=0.8733, eps_clip=0.2, k_epochs=epochs, device=dev, seed=5)

    def fill(ppo, T_, B_, gen, lo=0, hi=None):
        xs = torch.randint(-1, 11, (T_, B_, U, n_in), generator=gen, dtype=torch.int16, device=dev)
        acts = torch.randint(0, A, (T_, B_, U), generator=gen, dtype=torch.int32, device=dev)
        lps = torch.rand((T_, B_, U), generator=gen, device=dev) * 0.5 - 2.8
        rws = torch.randint(-5, 11, (T_, B_, U), generator=gen, device=dev).float()
        hi = B_ if hi is None else hi
        ppo.buf_x = [t[lo:hi].contiguous() for t in xs]
        ppo.buf_a = [t[lo:hi].contiguous() for t in acts]
        ppo.buf_lp = [t[lo:hi].contiguous() for t in lps]
        ppo.buf_r = [t[lo:hi].contiguous() for t in rws]

    # ---- check: sharded + all-reduced == whole buffer on one rank; weights bit-identical across ranks ----
    Bc = 512
    gen = torch.Generator(device=dev).manual_seed(1234)  # the same whole buffer on every rank
    ppo_s = BatchedPPO(n_in, A, 16, 1, U, 1, **kw)
    fill(ppo_s, 4, Bc * world, gen, rank * Bc, (rank + 1) * Bc)
    gen = torch.Generator(device=dev).manual_seed(1234)
    ppo_w = BatchedPPO(n_in, A, 16, 1, U, 1, **dict(kw, k_epochs=1))
    fill(ppo_w, 4, Bc * world, gen)
    unit = [5]
    # one epoch on each: gradient buffers (the whole-buffer instance must not all-reduce: call the kernel directly)
    Tn = 4
    Gs = policy.returns(torch.stack(ppo_s.buf_r).reshape(Tn, -1), kw["gamma"], True).view(Tn * Bc, U)
    Gw = policy.returns(torch.stack(ppo_w.buf_r).reshape(Tn, -1), kw["gamma"], True).view(Tn * Bc * world, U)
    ids = torch.zeros(1, dtype=torch.int32, device=dev)
    uid = torch.tensor([unit], dtype=torch.int32, device=dev)

    def grad_of(ppo, Gn, TB):
        flat = torch.zeros(ppo.actor.numel() + ppo.critic.numel(), device=dev)
        ga, gc = flat[: ppo.actor.numel()].view_as(ppo.actor), flat[ppo.actor.numel():].view_as(ppo.critic)
        policy.ppo_grad(ppo.actor.data, ppo.critic.data, n_in, A, torch.stack(ppo.buf_x).view(TB, U, n_in),
                        torch.stack(ppo.buf_a).view(TB, U), torch.stack(ppo.buf_lp).view(TB, U), Gn, ids, uid, ga, gc)
        return flat
    g_shard = grad_of(ppo_s, Gs, Tn * Bc)
    if world > 1:
        dist.all_reduce(g_shard)
        g_shard /= world
    g_whole = grad_of(ppo_w, Gw, Tn * Bc * world)
    rel = float((g_shard - g_whole).abs().max() / g_whole.abs().max())
    ppo_s.update(unit)  # K epochs with the all-reduce inside
    w = torch.cat([ppo_s.actor.data.view(-1), ppo_s.critic.data.view(-1)])
    wmax, wmin = w.clone(), w.clone()
    if world > 1:
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
    identical = bool(torch.equal(wmax.view(torch.int32), wmin.view(torch.int32)))
    moved = float((w - torch.cat([ppo_w.actor.data.view(-1), ppo_w.critic.data.view(-1)])).abs().max())
    del ppo_s, ppo_w

    # ---- timing at the full per-rank size ----
    ppo = BatchedPPO(n_in, A, 16, 1, U, 1, **kw)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    fill(ppo, T, B, gen)
    keep = (ppo.buf_x, ppo.buf_a, ppo.buf_lp, ppo.buf_r)
    ms = []
    for r in range(reps + 1):
        ppo.buf_x, ppo.buf_a, ppo.buf_lp, ppo.buf_r = keep
        ppo.profile = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ppo.update([3])
        e1.record()
        torch.cuda.synchronize()
        ep = [(a.elapsed_time(b), b.elapsed_time(c), c.elapsed_time(d)) for a, b, c, d in ppo.profile]
        ms.append((e0.elapsed_time(e1), sum(x[0] for x in ep) / len(ep), sum(x[1] for x in ep) / len(ep),
                   sum(x[2] for x in ep) / len(ep)))
    best = min(ms[1:], key=lambda t: t[0])
    t = torch.tensor(best, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    upd_ms, grad_ms, ar_ms, adam_ms = (float(v) for v in t)
    samples = T * B
    return {"what": "cfg4: globally shared acceptor net, one sampled unit per update (src/SchedulingEnvironment.py:314-329); "
                    "update = returns kernel + K x (msched_ppo_grad on the rank's env shard -> one flat all_reduce of both "
                    "heads' gradients -> msched_adam_step), through agents.BatchedPPO.update; device-timed, max over ranks",
            "net": f"{n_in}->16->16->{A} actor + {n_in}->16->16->1 critic, shared by {U} units",
            "ranks": world, "T": T, "envs_per_rank": B, "samples_per_rank_per_epoch": samples, "epochs": epochs,
            "update_ms": upd_ms, "epoch_grad_kernel_ms": grad_ms, "epoch_allreduce_us": 1e3 * ar_ms,
            "epoch_adam_us": 1e3 * adam_ms, "allreduce_bytes": 4 * (ppo.actor.numel() + ppo.critic.numel()),
            "samples_per_s_all_ranks": world * samples * epochs / (upd_ms * 1e-3),
            "check": {"buffer": f"T=4 x {Bc} envs per rank", "sharded_allreduced_vs_whole_buffer_grad_max_rel_err": rel,
                      "weights_bit_identical_across_ranks": identical, "weights_moved_by": moved,
                      "ok": bool(identical and rel < 1e-4 and moved > 0)}}


def other_configs(args):
    """The step leg of BASELINE.json's other configurations, each in its own process after this one's measurement is
    over (same timing rules, same K / W; parity of every one of them is the GPU test suite's job): not bench lines
    of their own, a record that the kernels those configurations select were run by this command."""
    import subprocess
    out = {}
    for key in ("cfg1", "cfg2", "cfg4", "cfg5"):
        cmd = [sys.executable, os.path.abspath(__file__), "--config", key, "--gpus", "1", "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--no-cpu-baseline", "--e2e-steps", "0",
               "--rollout-steps", "200" if key == "cfg1" else "0",  # cfg1 = the hard-coded agents' rollout
               "--update-T", "0", "--update-dp", "0", "--other-configs", "0"]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out[key] = {"workload": d["config"].get("domain"), "envs": d["config"].get("envs_per_gpu"),
                        "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                        "kernel": d["roofline"]["kernel"], "launches_per_step": d["gpu_launches"] / max(1, d["steps"]),
                        "roofline_frac": d["roofline"]["frac"],
                        "algorithmic_bytes_per_env_step": d["roofline"]["algorithmic_bytes_per_env_step"],
                        "l2": d.get("l2_detail"), "sticky_flags": d.get("sticky_flags_after_warm")}
            ms = d["roofline"].get("multi_step") or {}
            if "us_per_step" in ms:  # the dependent chain through msched_step_multi (T steps per launch)
                out[key]["multi_step"] = {"us_per_step": ms["us_per_step"], "frac": ms.get("frac"),
                                          "steps_per_launch": ms.get("steps_per_launch")}
            out[key]["serial_frac"] = d["roofline"].get("serial_frac")
            if d.get("rollout_with_policy"):
                out[key]["rollout"] = d["rollout_with_policy"]
        except Exception as e:  # a side leg must never take the bench line down
            out[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=list(CONFIGS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the config's)")
    ap.add_argument("--obs", default="dense", choices=["dense", "none", "compact"])
    ap.add_argument("--l2", default="auto", choices=["auto", "rotate", "flush"],
                    help="rotate: shards larger than L2 visited round-robin, back-to-back launches; "
                         "flush: 256 MiB write before every step, per-launch events")
    ap.add_argument("--sets", type=int, default=0, help="shards for --l2 rotate (default: >= 200 MB in flight)")
    ap.add_argument("--graph", type=int, default=1, help="--l2 rotate: replay each block from a CUDA graph")
    ap.add_argument("--device-round", action="store_true", help="device-side round counter without graphs")
    ap.add_argument("--streams", type=int, default=3,
                    help="streams the independent shards alternate between in --l2 rotate (1 = strictly serial launches)")
    ap.add_argument("--no-multi", action="store_true", help="skip the msched_step_multi (T steps per launch) measurement")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (diagnostic)")
    ap.add_argument("--state-warm", type=int, default=1000, help="untimed steps to reach steady state")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--update-T", type=int, default=16, help="buffer length of the PPO.update measurement (0 = skip)")
    ap.add_argument("--update-dp", type=int, default=1, help="the cfg4 data-parallel PPO update leg (0 = skip)")
    ap.add_argument("--update-dp-T", type=int, default=8, help="buffer length (steps) of that leg")
    ap.add_argument("--rollout-steps", type=int, default=200,
                    help="steps of the secondary metric (env step + actor forward); 0 = skip")
    ap.add_argument("--other-configs", type=int, default=1,
                    help="N=1, default workload only: after the measurement, run the step leg of BASELINE's other "
                         "configurations (each its own process, one after the other) and report them under "
                         "`other_configs`; 0 = skip")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from marl_scheduling_b200 import _lib as L
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv, world_params_from_dom

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from marl_scheduling_b200.distributed import pin_host_to_gpu
    host_pin = pin_host_to_gpu(local)  # before any pinned host buffer is allocated
    if world > 1:
        # NCCL prints its version banner to stdout on the first communicator: keep stdout for the ONE JSON line
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            try:  # NCCL writes through C stdio: flush ITS buffer while fd 1 still points at stderr
                import ctypes
                ctypes.CDLL(None).fflush(None)
            except Exception:
                pass
            os.dup2(keep, 1)
            os.close(keep)

    dom, mode = cfg["dom"], cfg["mode"]
    B = args.envs or cfg["envs"]
    N = dom["N"]
    if cfg.get("obs") in ("none", "compact"):
        args.obs = cfg["obs"]
    dense = args.obs == "dense"
    compact = args.obs == "compact"
    serial_us = None
    multi = None

    def make_env(k):
        return BatchedSchedulingEnv(B, world_params_from_dom(dom, mode.startswith("free")), reward=mode,
                                    auction="random", spawn="philox", seed=0,
                                    env_offset=(rank * 64 + k) * B,
                                    net_zero_offer_reward=dom.get("netZero", 0.5), device=local)

    env = make_env(0)
    lay = env.layout
    info = env.info()
    fused = dense and info["fuses_observations"]
    l2 = args.l2
    if l2 == "auto":  # one launch per step: time back-to-back launches over rotating shards
        l2 = "rotate" if (fused or not dense) else "flush"
    if args.no_flush:
        l2 = "warm"
    K = args.steps
    sampler = ClockSampler(local)

    def step_on(e, action, result):
        if dense:
            e.step_observe_records(action, result)
        elif compact:   # transition + compact observations (ONE launch on the warp-per-environment kernel)
            e.step_compact_records(action, result)
        else:
            e.step_records(action, result)

    if l2 == "rotate":
        # ---- inputs larger than L2: S independent shards of B envs each, visited round-robin, so a
        # shard's records were last touched (S-1) launches ago and (S-1) x its working set has gone
        # through the 126 MB L2 since.  Launches are back to back on one stream; each timed block
        # of S*G launches has its own freshly drawn action records (drawn untimed between blocks).
        per_set = B * (lay.state_words * 4 + lay.action_halfs * 2 + lay.result_words * 4 +
                       (lay.obs_halfs * 2 if dense else (lay.cobs_halfs * 2 if compact else 0)))
        big = per_set > 1_000_000_000  # one shard alone is many times the 126 MB L2 (config 5: 5.9 GB)
        S = args.sets or (2 if big else max(3, -(-200_000_000 // per_set) + 1))
        G = 2 if big else 8
        envs = [env] + [make_env(k) for k in range(1, S)]
        gen = torch.Generator(device=dev).manual_seed(1 + rank)
        recs = [torch.zeros((lay.padded_envs, lay.action_halfs), dtype=torch.int16, device=dev)
                for _ in range(S * G)]
        results = [torch.zeros_like(env.result) for _ in range(S)]

        # independent shards alternate between two streams (shard k always on stream k % 2, so a
        # shard's launches stay ordered): the first CTAs of one launch fill the SMs that the tail
        # of the previous launch has already left.  A block is timed on the default stream, which
        # the two streams fork from and join back into.
        streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]
        use_graph = args.graph == 1
        if use_graph or args.device_round:
            for e in envs:
                e.set_device_round(True)  # world.round in a device counter: a block can be replayed from a graph

        def launch_block(n, strs=None):
            strs = strs or streams
            cur = torch.cuda.current_stream(dev)
            fork = torch.cuda.Event()
            fork.record(cur)
            for st in strs:
                st.wait_event(fork)
            for i in range(n):
                with torch.cuda.stream(strs[(i % S) % len(strs)]):
                    step_on(envs[i % S], recs[i], results[i % S])
            for st in strs:
                join = torch.cuda.Event()
                join.record(st)
                cur.wait_event(join)

        # the Python launch loop costs ~15 us per launch, as much as the kernel: a block of launches is
        # captured ONCE in a CUDA graph (its action records are refreshed in place between replays)
        graphs = {}

        def block_graph(n, strs=None):
            key = (n, len(strs or streams))
            if key not in graphs:
                launch_block(n, strs)  # warm-up outside the capture
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream(device=dev)
                cap.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(cap):
                    with torch.cuda.graph(g, stream=cap):
                        launch_block(n, strs)
                torch.cuda.current_stream(dev).wait_stream(cap)
                graphs[key] = g
            return graphs[key]

        def run_block(n, timed, strs=None):
            g = block_graph(n, strs) if use_graph else None
            for r in recs[:n]:
                refresh_actions(env, r, gen)
            if timed is not None:
                # the block is enqueued BEHIND a ~150 us spin kernel: by the time the first event is stamped the graph
                # launch is already in the queue, so the host's enqueue latency (13-15 us, as much as a step) is not
                # inside the timed window -- the events bracket the device's execution of exactly n steps
                torch.cuda._sleep(300000)
                timed[0].record()
            if g is not None:
                g.replay()
            else:
                launch_block(n, strs)
            if timed is not None:
                timed[1].record()

        # the clocks are sampled from here on: the state warm-up runs the same launches as the timed region, which at
        # the default step count is too short (a fraction of a millisecond) for more than one NVML reading
        sampler.start()
        for _ in range(-(-args.state_warm * S // (S * G))):
            run_block(S * G, None)
        torch.cuda.synchronize()
        flags = max(int(r[:B, lay.r_flags].max().item()) for r in results) if args.state_warm else 0
        for _ in range(-(-max(args.warmup, 3) // (S * G))):
            run_block(S * G, None)
        blocks = [min(S * G, K - b0) for b0 in range(0, K, S * G)]
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in blocks]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall0 = time.perf_counter()
        for n, e in zip(blocks, ev):
            run_block(n, e)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.finish()
        tot_ms = sum(e[0].elapsed_time(e[1]) for e in ev)
        t = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms = stepk_ms = float(t[0])
        obs_us = None
        # the same launches strictly one after the other on ONE stream (what a single dependent rollout sees)
        if len(streams) > 1 and rank == 0:
            one = streams[:1]
            run_block(S * G, None, one)
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(6)]
            torch.cuda.synchronize()
            for e in evs:
                run_block(S * G, e, one)
            torch.cuda.synchronize()
            serial_us = 1e3 * sum(e[0].elapsed_time(e[1]) for e in evs) / (len(evs) * S * G)
        # msched_step_multi: T steps per launch (a CTA keeps its 32 environments; observations after EVERY step, so each
        # step does the work of the headline step; the state is re-read from L2 instead of HBM between a launch's steps)
        if rank == 0 and fused and not args.no_multi:
            try:
                Tm = 8
                m_acts = [torch.zeros((Tm, lay.padded_envs, lay.action_halfs), dtype=torch.int16, device=dev) for _ in range(S)]
                m_res = [torch.zeros((Tm, lay.padded_envs, lay.result_words), dtype=torch.int32, device=dev) for _ in range(S)]
                m_obs = [torch.zeros((Tm, lay.padded_envs, lay.obs_halfs), dtype=torch.int16, device=dev) for _ in range(S)]

                def multi_block():
                    for k in range(S):
                        envs[k].step_multi_records(m_acts[k], m_res[k], m_obs[k], obs_every=True)

                def multi_refresh():
                    for k in range(S):
                        for t_ in range(Tm):
                            refresh_actions(env, m_acts[k][t_], gen)
                multi_refresh()
                multi_block()
                torch.cuda.synchronize()
                gm = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream(device=dev)
                cap.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(cap):
                    with torch.cuda.graph(gm, stream=cap):
                        multi_block()
                torch.cuda.current_stream(dev).wait_stream(cap)
                evm = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(6)]
                for e in evm:
                    multi_refresh()
                    e[0].record()
                    gm.replay()
                    e[1].record()
                torch.cuda.synchronize()
                multi_us = 1e3 * sum(e[0].elapsed_time(e[1]) for e in evm) / (len(evm) * S * Tm)
                multi = {"steps_per_launch": Tm, "us_per_step": multi_us,
                         "what": "msched_step_multi: %d dependent steps per launch, observations after every step, launches of %d "
                                 "shards one after the other on ONE stream (graph replay)" % (Tm, S)}
                del m_acts, m_res, m_obs
            except L.MschedError as e:
                multi = {"unavailable": str(e)}
        l2_note = (f"inputs larger than L2: {S} shards x {per_set / 1e6:.0f} MB visited round-robin, launches "
                   f"in blocks of {S * G} on {len(streams)} stream(s)" + (", each block one CUDA-graph replay" if use_graph else "") +
                   ", every timed block enqueued behind a 150 us spin kernel (host enqueue latency outside the events)")
        n_launch = K
    else:
        ring, gen = make_actions(torch, env, cfg.get("ring", 8), seed=1 + rank)
        results = [torch.zeros_like(env.result) for _ in range(2)]
        flush = None if l2 == "warm" else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def one_step(i):
            refresh_actions(env, ring[i % len(ring)], gen)
            step_on(env, ring[i % len(ring)], results[i & 1])

        sampler.start()  # from the state warm-up on (the same launches as the timed region)
        for i in range(args.state_warm):
            one_step(i)
        torch.cuda.synchronize()
        flags = int(results[(args.state_warm - 1) & 1][:B, lay.r_flags].max().item()) if args.state_warm else 0
        for i in range(max(args.warmup, 3)):
            if flush is not None:
                flush.fill_(i & 0xFF)
            one_step(i)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall0 = time.perf_counter()
        for i in range(K):
            refresh_actions(env, ring[i % len(ring)], gen)
            if flush is not None:
                flush.fill_(i & 0xFF)
            ev[i][0].record()
            if fused or not dense:      # ONE launch: transition (+ observations of the new state)
                step_on(env, ring[i % len(ring)], results[i & 1])
                ev[i][1].record()
            else:
                env.step_records(ring[i % len(ring)], results[i & 1])
                ev[i][1].record()
                env.observe()
            ev[i][2].record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.finish()
        step_ms = [e[0].elapsed_time(e[1]) for e in ev]
        obs_ms = [e[1].elapsed_time(e[2]) for e in ev]
        tot_ms = sum(step_ms) + sum(obs_ms)
        t = torch.tensor([tot_ms, sum(step_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms, stepk_ms = float(t[0]), float(t[1])
        obs_us = None if (fused or not dense) else 1e3 * sum(obs_ms) / K
        l2_note = "warm (no flush)" if l2 == "warm" else "flushed before every timed step (256 MiB write), per-launch events"
        n_launch = K * (2 if (dense and not fused) else 1)
    value = world * B * N * K / (tot_ms * 1e-3)

    # ---- e2e: host records in, host records out, through msched_step_host ----
    e2e = None
    nE = min(args.e2e_steps, K) if K > 0 else 0
    if nE > 0:
        ah = []
        for _ in range(16):
            tmp = torch.zeros((lay.padded_envs, lay.action_halfs), dtype=torch.int16, device=dev)
            refresh_actions(env, tmp, gen)
            ah.append(tmp[:B].cpu().pin_memory())
        # the result comes back as the COMPACT record (int16 / half planes) when the domain has one; the full record
        # is timed beside it
        try:
            cwords = env.compact_result_layout().words if info["step_impl"] == "fused" else 0
        except L.MschedError:
            cwords = 0
        rh_full = torch.zeros((B, lay.result_words), dtype=torch.int32).pin_memory()
        rh = torch.zeros((B, cwords), dtype=torch.int32).pin_memory() if cwords else rh_full
        host_step = env.step_host_compact if cwords else env.step_host
        for i in range(5):
            env.step_host(ah[i % 16], rh_full, observe=dense)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(nE):
            env.step_host(ah[i % 16], rh_full, observe=dense)
        torch.cuda.synchronize()
        te_full = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te_full, op=dist.ReduceOp.MAX)
        for i in range(5):
            host_step(ah[i % 16], rh, observe=dense)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(nE):
            host_step(ah[i % 16], rh, observe=dense)  # observations stay on the device for the policy kernels
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # what bounds e2e: the PCIe copies of the step's records (measured here, pinned 64 MiB, both directions)
        hb = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        db = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        pcie = {}
        if world > 1:
            dist.barrier()  # every rank copies at the same time: the rates include the contention on the host side
        for name, (dst, src) in (("h2d_gbs", (db, hb)), ("d2h_gbs", (hb, db))):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            pcie[name] = 4 * (64 << 20) / (c0.elapsed_time(c1) * 1e-3) / 1e9
        rwords = cwords or lay.result_words
        t_bound = max(B * lay.action_halfs * 2 / (pcie["h2d_gbs"] * 1e9), B * rwords * 4 / (pcie["d2h_gbs"] * 1e9))
        pcie["bound_value"] = world * B * N / t_bound  # copies at the measured rates, both directions fully overlapped
        pcie["frac_of_bound"] = world * B * N * nE / float(te[0]) / pcie["bound_value"]
        if world > 1:  # the slowest rank's copy rates (all ranks copying concurrently) bound the job
            tr = torch.tensor([pcie["h2d_gbs"], pcie["d2h_gbs"]], dtype=torch.float64, device=dev)
            dist.all_reduce(tr, op=dist.ReduceOp.MIN)
            pcie["h2d_gbs_min_rank"], pcie["d2h_gbs_min_rank"] = float(tr[0]), float(tr[1])
        pcie["host_pinning"] = host_pin
        pcie["limiter"] = ("PCIe / host memory: the records of a step cross the bus once in each direction; %d rank(s) copying "
                           "concurrently reach %.0f / %.0f GB/s per GPU (H2D / D2H, copy engine), the kernel's SM-originated "
                           "zero-copy traffic runs at %.2f of the bound those rates set" %
                           (world, pcie["h2d_gbs"], pcie["d2h_gbs"], pcie["frac_of_bound"]))
        del hb, db
        e2e = {"value": world * B * N * nE / float(te[0]), "unit": "agent-steps/s", "pcie": pcie,
               "h2d_bytes_per_step": B * lay.action_halfs * 2, "d2h_bytes_per_step": B * rwords * 4,
               "full_record": {"value": world * B * N * nE / float(te_full[0]), "d2h_bytes_per_step": B * lay.result_words * 4,
                               "api": "msched_step_host"},
               "steps": nE, "api": ("msched_step_host_compact" if cwords else "msched_step_host") +
                                   " (pinned action records -> " + ("compact " if cwords else "") + "result records, read and written by "
                                   "the kernel's bulk copies over PCIe (zero-copy); observations stay on the device)"}

    # ---- secondary metric: the same step INCLUDING the batched actor forward of the
    # divided PPO agents (one net per unit, src/Agent.py:495-619) and, separately, the return scan ----
    rollout = None
    if args.rollout_steps > 0 and dense and world == 1 and cfg.get("policy") == "hardcoded":
        # config 1: the reference's heuristic agents (msched_hardcoded_actions) + the fused env step, from a CUDA graph
        env.set_device_round(True)
        res_r = torch.zeros_like(env.result)

        def hc_step():
            env.hardcoded_actions(random_ties=True)
            env.step_observe_records(env.action, res_r)
        for i in range(5):
            hc_step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            with torch.cuda.graph(graph, stream=cap):
                for k in range(8):
                    hc_step()
        torch.cuda.current_stream(dev).wait_stream(cap)
        graph.replay()
        torch.cuda.synchronize()
        n_rep = max(1, args.rollout_steps // 8)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_rep):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        r_ms = e0.elapsed_time(e1) / (n_rep * 8)
        # the same loop inside ONE launch (msched_rollout_hardcoded: the agents read the observation tile in shared
        # memory, their actions become the next step's action tile), observations written after every step
        one = None
        try:
            Tm = 8
            res_m = torch.zeros((Tm, lay.padded_envs, lay.result_words), dtype=torch.int32, device=dev)
            obs_m = torch.zeros((Tm, lay.padded_envs, lay.obs_halfs), dtype=torch.int16, device=dev)
            env.hardcoded_actions(random_ties=True)
            for i in range(3):
                env.rollout_hardcoded(res_m, obs_m, obs_every=True, random_ties=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(n_rep):
                env.rollout_hardcoded(res_m, obs_m, obs_every=True, random_ties=True)
            e1.record()
            torch.cuda.synchronize()
            m_ms = e0.elapsed_time(e1) / (n_rep * Tm)
            one = {"value": B * N / (m_ms * 1e-3), "ms_per_step": m_ms, "steps_per_launch": Tm,
                   "sticky_flags": int(res_m[:, :B, lay.r_flags].max().item()),
                   "what": "msched_rollout_hardcoded: %d x (step ; hard-coded agents on the new observations) per launch, "
                           "observations written after every step" % Tm}
        except L.MschedError as e:
            one = {"unavailable": str(e)}
        env.set_device_round(False)
        rollout = {"value": B * N / (r_ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": r_ms, "steps": n_rep * 8,
                   "launches_per_step": 2, "sticky_flags": int(res_r[:B, lay.r_flags].max().item()),
                   "what": "DividedHardcodedAgent.getActions of every agent (msched_hardcoded_actions, Philox ties) + the fused env "
                           "step + observations, 8 steps per CUDA-graph replay",
                   "one_launch": one}
    elif args.rollout_steps > 0 and dense and world == 1:
        from marl_scheduling_b200 import policy
        free = mode.startswith("free")
        Cc, Lc, NL = dom["C"], dom["L"], N * dom["L"]
        P = max(dom["prios"])
        acc_net = policy.MlpGroup.random(3 + 2 * NL, 16, NL + 1, N * Cc, dev, seed=1)
        off_net = policy.MlpGroup.random(2 * Cc + 2, 16, Cc + 1, NL, dev, seed=2)
        price_net = policy.MlpGroup.random(4, 16, P + 1, NL, dev, seed=3) if free else None
        ov = env.obs_views()
        a_act = torch.empty(B * N * Cc, dtype=torch.int32, device=dev)
        a_lp = torch.empty(B * N * Cc, dtype=torch.float32, device=dev)
        o_act = torch.empty(B * NL, dtype=torch.int32, device=dev)
        o_lp = torch.empty(B * NL, dtype=torch.float32, device=dev)
        p_act, p_lp = torch.empty_like(o_act), torch.empty_like(o_lp)
        res_r = torch.zeros_like(env.result)

        # ONE launch for every PPO unit of the step (msched_policy_step: acceptor units, core chooser and price
        # chooser of every offer unit; actions straight into the action record; state / action / log-prob into the
        # experience-buffer slot of the step), then the fused env step + observations.  EIGHT consecutive steps --
        # eight different buffer slots, 250 MB of experience written -- are captured in one CUDA graph and replayed:
        # world.round and the policy's Philox step live in device counters.
        SLOTS = 8
        step_t = torch.zeros(1, dtype=torch.int64, device=dev)
        env.set_device_round(True)
        one_launch = policy.policy_step_supported(acc_net, off_net, price_net)
        ring = dict(xa=torch.zeros((SLOTS, B, N * Cc, lay.o_acc_row), dtype=torch.int16, device=dev),
                    xo=torch.zeros((SLOTS, B, NL, lay.o_off_row), dtype=torch.int16, device=dev),
                    xp=torch.zeros((SLOTS, B, NL, 4), dtype=torch.int16, device=dev),
                    a=torch.zeros((SLOTS, 3, B, max(N * Cc, NL)), dtype=torch.int32, device=dev),
                    lp=torch.zeros((SLOTS, 3, B, max(N * Cc, NL)), dtype=torch.float32, device=dev))

        def rollout_step(k=0):
            if one_launch:
                ga = policy.policy_step_group(acc_net, N * Cc, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor, 1,
                                              ring["a"][k, 0, :, :N * Cc], ring["lp"][k, 0, :, :N * Cc], x_used=ring["xa"][k])
                go = policy.policy_step_group(off_net, NL, lay.o_offer, lay.o_off_row, lay.a_offer_core, 2,
                                              ring["a"][k, 1, :, :NL], ring["lp"][k, 1, :, :NL], x_used=ring["xo"][k])
                gp = policy.policy_step_group(price_net, NL, lay.o_offer, lay.o_off_row, lay.a_offer_price, 3,
                                              ring["a"][k, 2, :, :NL], ring["lp"][k, 2, :, :NL], x_used=ring["xp"][k]) if free else None
                policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, Cc, ga, go, gp, action_rec=env.action,
                                   action_rec_stride=lay.action_halfs, env_offset=0, step_dev=step_t, input_bound=max(max(dom['prios']), max(dom['lens']), 8))
            else:  # shapes without a one-launch kernel: one launch per net group
                policy.actor_forward(off_net, ov["offer"], lay.o_off_row, NL, B, env_stride=lay.obs_halfs, seed=2,
                                     step_dev=step_t, action=o_act, logprob=o_lp, action_rec=env.offer_core_actions,
                                     action_rec_stride=lay.action_halfs)
                if free:
                    policy.actor_forward(price_net, ov["offer"], lay.o_off_row, NL, B, env_stride=lay.obs_halfs,
                                         seed=3, step_dev=step_t, action=p_act, logprob=p_lp,
                                         action_rec=env.offer_price_actions, action_rec_stride=lay.action_halfs,
                                         gather_core=o_act, n_cores=Cc)
                policy.actor_forward(acc_net, ov["acceptor"], lay.o_acc_row, N * Cc, B, env_stride=lay.obs_halfs,
                                     seed=1, step_dev=step_t, action=a_act, logprob=a_lp,
                                     action_rec=env.acceptor_actions, action_rec_stride=lay.action_halfs)
            if side is not None:  # graph capture: the policy's step counter moves on a branch beside the env step
                cur = torch.cuda.current_stream(dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    step_t.add_(1)
                env.step_observe_records(env.action, res_r)
                cur.wait_stream(side)
            else:
                env.step_observe_records(env.action, res_r)
                step_t.add_(1)

        side = None
        for i in range(5):
            rollout_step(i % SLOTS)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(cap):
            with torch.cuda.graph(graph, stream=cap):
                for k in range(SLOTS):
                    rollout_step(k)
        side = None
        torch.cuda.current_stream(dev).wait_stream(cap)
        for i in range(3):
            graph.replay()
        torch.cuda.synchronize()
        n_rep = max(1, args.rollout_steps // SLOTS)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_rep):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        r_ms = e0.elapsed_time(e1) / (n_rep * SLOTS)
        # the policy launch alone (same graph-free launch, L2 warm), for the split of the step
        e0.record()
        for i in range(20):
            if one_launch:
                policy.policy_step(env._obs_buffer(), lay.obs_halfs, B, Cc, *[g for g in (
                    policy.policy_step_group(acc_net, N * Cc, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor, 1, ring["a"][0, 0, :, :N * Cc], ring["lp"][0, 0, :, :N * Cc], x_used=ring["xa"][0]),
                    policy.policy_step_group(off_net, NL, lay.o_offer, lay.o_off_row, lay.a_offer_core, 2, ring["a"][0, 1, :, :NL], ring["lp"][0, 1, :, :NL], x_used=ring["xo"][0]),
                    policy.policy_step_group(price_net, NL, lay.o_offer, lay.o_off_row, lay.a_offer_price, 3, ring["a"][0, 2, :, :NL], ring["lp"][0, 2, :, :NL], x_used=ring["xp"][0]) if free else None)],
                    action_rec=env.action, action_rec_stride=lay.action_halfs, env_offset=0, step_dev=step_t, input_bound=max(max(dom['prios']), max(dom['lens']), 8))
        e1.record()
        torch.cuda.synchronize()
        pol_us = e0.elapsed_time(e1) * 1e3 / 20 if one_launch else None
        rflags = int(res_r[:B, lay.r_flags].max().item())
        env.set_device_round(False)
        T = 200
        rew = torch.randn(T, B * N * Cc, device=dev)
        for _ in range(2):
            policy.returns(rew, 0.8733, True)
        e0.record()
        for _ in range(5):
            policy.returns(rew, 0.8733, True)
        e1.record()
        torch.cuda.synchronize()
        ret_us = e0.elapsed_time(e1) * 1e3 / 5
        macs = N * Cc * (16 * (3 + 2 * NL) + 256 + 16 * (NL + 1)) + NL * (16 * (2 * Cc + 2) + 256 + 16 * (Cc + 1)) + \
            (NL * (64 + 256 + 16 * (P + 1)) if free else 0)
        rollout = {"value": B * N / (r_ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": r_ms,
                   "steps": n_rep * SLOTS, "launches_per_step": 2 if one_launch else (4 if free else 3),
                   "policy_launch_us": pol_us, "policy_macs_per_env": macs,
                   "policy_tflops": None if not pol_us else 2.0 * macs * B / pol_us / 1e6,
                   "experience_bytes_per_step": int(B * (2 * (N * Cc * lay.o_acc_row + NL * lay.o_off_row + (4 * NL if free else 0))
                                                          + 8 * (N * Cc + NL * (2 if free else 1)))),
                   "what": "PPO.selectAction of every divided PPO unit (acceptor, offer/core chooser"
                           + (", price chooser" if free else "") + ") in ONE launch (msched_policy_step: sample + log-prob, actions "
                           "into the action record, state / action / log-prob into the experience-buffer slot of the step) + the fused "
                           "env step + observations; 8 consecutive steps (8 buffer slots) captured in a CUDA graph and replayed (the one-thread increment of the policy's step counter sits on a graph branch beside the env step)",
                   "sticky_flags": rflags,
                   # 8 B/element (read r, write G); the tiled kernels normalise in shared memory, so
                   # that is also their HBM traffic (rounds 1-2 counted 12 B for a streaming second pass)
                   "returns_kernel": {"T": T, "units": B * N * Cc, "us": ret_us,
                                      "gbs": 8.0 * rew.numel() / ret_us / 1e3,
                                      "hbm_frac": 8.0 * rew.numel() / ret_us / 1e3 / measured_peak()[0],
                                      "algorithmic_bytes_per_element": 8}}

    # ---- PPO.update on the device: one epoch = one forward+backward launch over the
    # acceptor units' buffer (T x B samples per net) + Adam; the PyTorch autograd version of the same
    # update is timed beside it on a smaller buffer ----
    update = None
    if args.update_T > 0 and dense and world == 1:
        from marl_scheduling_b200 import policy
        from marl_scheduling_b200.agents import BatchedPPO
        Cc, NL = dom["C"], N * dom["L"]
        U, n_in, A = N * Cc, 3 + 2 * NL, NL + 1
        kw = dict(lr_actor=3e-4, lr_critic=1e-3, gamma=0.9, eps_clip=0.2, k_epochs=1, device=dev, seed=5)
        gen_u = torch.Generator(device=dev).manual_seed(7)

        def timed_update(mode, T, Bu, reps):
            os.environ["MSCHED_PPO_UPDATE"] = mode
            ppo = BatchedPPO(n_in, A, 16, U, U, 1, **kw)
            xs = torch.randint(-1, 9, (T, Bu, U, n_in), generator=gen_u, dtype=torch.int16, device=dev)
            acts = torch.randint(0, A, (T, Bu, U), generator=gen_u, dtype=torch.int32, device=dev)
            lps = torch.full((T, Bu, U), -1.9459, device=dev)
            rws = torch.randint(-5, 9, (T, Bu, U), generator=gen_u, device=dev).float()
            ms = []
            for _ in range(reps + 1):
                ppo.buf_x, ppo.buf_a, ppo.buf_lp, ppo.buf_r = list(xs), list(acts), list(lps), list(rws)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ppo.update()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            grad_ms = None
            if mode == "kernel":  # the gradient launch alone
                X = xs.view(T * Bu, U, n_in)
                Gn = policy.returns(rws.view(T, Bu * U), 0.9, True).view(T * Bu, U)
                ids = torch.arange(U, dtype=torch.int32, device=dev)
                ga, gc = torch.zeros_like(ppo.actor.data), torch.zeros_like(ppo.critic.data)
                _, ws = policy.ppo_grad(ppo.actor.data, ppo.critic.data, n_in, A, X, acts.view(T * Bu, U), lps.view(T * Bu, U),
                                        Gn, ids, ids.view(U, 1), ga, gc)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    policy.ppo_grad(ppo.actor.data, ppo.critic.data, n_in, A, X, acts.view(T * Bu, U), lps.view(T * Bu, U),
                                    Gn, ids, ids.view(U, 1), ga, gc, workspace=ws)
                e1.record()
                torch.cuda.synchronize()
                grad_ms = e0.elapsed_time(e1) / reps
            os.environ.pop("MSCHED_PPO_UPDATE", None)
            del xs, acts, lps, rws
            return min(ms[1:]), grad_ms

        Tk = args.update_T
        k_ms, g_ms = timed_update("kernel", Tk, B, 3)
        Bs = max(1, B // 16)
        a_ms, _ = timed_update("autograd", Tk, Bs, 2)
        k_small_ms, _ = timed_update("kernel", Tk, Bs, 3)
        rows = Tk * B * U
        update = {"what": "PPO.update, one epoch, acceptor nets of the workload (one net per unit): returns + forward/backward "
                          "gradient kernel (msched_ppo_grad) + Adam (msched_adam_step)",
                  "net": f"{n_in}->16->16->{A} actor + {n_in}->16->16->1 critic, {U} nets", "T": Tk, "envs": B, "samples": rows,
                  "update_ms": k_ms, "grad_kernel_ms": g_ms, "samples_per_s": rows / (g_ms * 1e-3),
                  "autograd_baseline": {"envs": Bs, "samples": Tk * Bs * U, "update_ms": a_ms, "kernel_path_update_ms_same_size": k_small_ms,
                                        "speedup": a_ms / k_small_ms, "what": "the same update through PyTorch autograd + torch.optim.Adam"}}

    # ---- config 4: data-parallel PPO update (kernel gradient + NCCL all-reduce + Adam), every rank ----
    update_dp = None
    if args.update_dp and dense:
        update_dp = ppo_update_dp_leg(torch, dist, dev, world, rank, T=args.update_dp_T)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ab = algorithmic_bytes(dom, mode)
    peak, peak_src = measured_peak()
    step_launch_s = stepk_ms * 1e-3 / K
    kname = {"lane": "step_kernel", "coop": "coop_step_kernel", "fused": "fused_step_kernel",
             "warp": "warp_step_kernel"}[info["step_impl"]]
    # the fused launch also writes the dense observation record: SURVEY 8(d) bytes = 2S + a + r + o; the warp
    # kernel's launch the compact one: 2S + a + r + o_c
    one_launch_compact = compact and info["step_impl"] == "warp"
    alg_bytes = ab["step"] + (ab["obs"] if fused else (ab["obs_compact"] if one_launch_compact else 0))
    achieved = alg_bytes * B / step_launch_s / 1e9
    cpu = cpu_port = None
    if not args.no_cpu_baseline and world == 1:
        small = N * dom["L"] <= 64
        if small:
            c = cpu_baseline_sample(cfg, seconds=args.cpu_seconds / 2, threads=1, envs=2048)
            cpu_port = {"value": c["value"], "unit": "agent-steps/s", "cores": 1, "kind": "port",
                        "sample": f"{c['envs']} envs x {c['steps']} steps of the same workload, 1 thread of "
                                  f"{os.cpu_count()} host cores, oracle/msched_oracle.c"}
        r = python_reference_sample(cfg, envs=64 if small else 1, warm_steps=200 if small else 3, seconds=args.cpu_seconds)
        if r is not None:
            cpu = {"value": r["value"], "unit": "agent-steps/s", "cores": 1, "kind": "reference",
                   "sample": f"{r['envs']} independent worlds x {r['steps']} steps ({r['seconds']:.1f} s) after {r['warm_steps']} untimed, "
                             f"unmodified Python reference SchedulingEnv.step + getAuctioneerAction, 1 thread of {os.cpu_count()} host cores"}
        else:
            cpu = cpu_port
    line = {
        "metric": "agent-steps/sec, batched env step+auction @65,536 envs",
        "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": tot_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, cfg, B),
        "l2_detail": l2_note,
        "kernel_config": {"auctioneer": "in-kernel, random arg-max (Philox)", "spawn": "device Philox",
                          "state_warm_steps": args.state_warm, "step_impl": info["step_impl"],
                          "observations_fused_into_step_launch": bool(fused or (compact and info["step_impl"] == "warp")),
                          "envs_per_cta": info["envs_per_cta"], "smem_bytes_per_cta": info["smem_bytes_per_cta"]},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": n_launch,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_from_profile(args.config, kname),
                     "peak_source": peak_src, "algorithmic_bytes_per_env_step": alg_bytes,
                     "algorithmic_bytes_formula": "2S+a+r" + ("+o (dense observations)" if fused else
                                                              ("+o_c (compact observations)" if one_launch_compact else "")),
                     "units_per_launch": B, "launch_us": step_launch_s * 1e6,
                     "serial_launch_us": serial_us,
                     "serial_frac": (alg_bytes * B / (serial_us * 1e-6) / 1e9 / peak) if serial_us else None,
                     "multi_step": (dict(multi, frac=alg_bytes * B / (multi["us_per_step"] * 1e-6) / 1e9 / peak)
                                    if multi and "us_per_step" in multi else multi)},
        "kernels": {"step_us": 1e3 * stepk_ms / K, "observe_us": obs_us,
                    "observe_algorithmic_bytes_per_env": ab["obs"] + ab["state"]},
        "cpu_baseline": cpu,
        "cpu_baseline_port": cpu_port,
        "rollout_with_policy": rollout,
        "ppo_update": update,
        "ppo_update_dp": update_dp,
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
        "sticky_flags_after_warm": flags,
    }
    if world == 1 and args.config == "cfg3" and args.other_configs and not args.envs:
        line["other_configs"] = other_configs(args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
