#!/usr/bin/env python
"""Long-run pin of the C oracle against the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  SURVEY.md section 8(c) asks for >= 1e5 reference env-steps across all
reward variants before the restatement is trusted; the committed fixtures hold ~5.7 k.  This script
drives the stub-imported reference with fresh seeds (random policy biased towards valid acceptances,
the reference's own auctioneer with its random tie-break, recorded spawn draws), replays every step
through oracle/msched_oracle.c and compares, bit-exactly and step by step: the full integer state
(cores, slots, offers, liability chains), accepted offers and terminations, all rewards, done,
acception quality (1e-12), the dense observations and the offer-ID tables.

  python oracle/longrun_check.py [--steps-scale 1.0] [--out oracle/LONGRUN_RESULT.json]

The last result is committed as oracle/LONGRUN_RESULT.json.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import oracle as O  # noqa: E402  (oracle/oracle.py)
import ref_harness as H  # noqa: E402
from gen_golden import DOMS  # noqa: E402
from helpers import assert_state_equal  # noqa: E402

# (domain, reward mode, agent kind, steps per run, seeds)
PLAN = [
    ("C", "free_comm", "divided", 6000, (101, 102, 103, 104)),
    ("C", "free_ncomm", "divided", 6000, (111, 112, 113)),
    ("B", "fix", "divided", 5000, (121, 122, 123)),
    ("B", "agg", "divided", 5000, (131, 132)),
    ("A", "fix", "divided", 6000, (141, 142)),
    ("E", "agg", "divided", 5000, (151,)),
    ("G", "free_comm", "divided", 4000, (161,)),
    ("G", "fix", "divided", 4000, (171,)),
    ("H", "fix", "divided", 3000, (181,)),
    ("F", "free_ncomm", "divided", 1500, (191,)),
    ("F", "fix", "divided", 1500, (201,)),
    ("S", "fix", "divided", 2000, (211,)),
    ("A", "fix", "hardcoded", 6000, (221, 222)),
]


def replay(tr, dom, mode, agent_kind, K):
    """One recorded reference trace through the oracle; raises AssertionError on the first difference."""
    T = tr["done"].shape[0]
    free = mode.startswith("free")
    orc = O.Oracle(1, dom, mode, chain_cap=K)
    for t in range(T):
        orc.step(tr["in_offc"][t][None], tr["in_acc"][t][None], tr["in_auc"][t][None],
                 offp=tr["in_offp"][t][None] if free else None, spawn_u=tr["in_spawn_u"][t][None])
        e = orc.export(0)
        assert e["flags"] == 0, (t, e["flags"])
        assert_state_equal(e, tr, t)
        na, nt = int(tr["n_accepted"][t]), int(tr["n_term"][t])
        assert orc.n_accepted[0] == na and orc.n_term[0] == nt, t
        assert np.array_equal(e["accepted"][:na], tr["accepted"][t][:na]), t
        assert np.array_equal(e["term"][:nt, :5], tr["term"][t][:nt]), t
        for k, got in (("r_offer", orc.r_offer), ("r_price", orc.r_price), ("r_acceptor", orc.r_acceptor),
                       ("r_auctioneer", orc.r_auctioneer), ("r_agent", orc.r_agent)):
            assert np.array_equal(got[0], tr[k][t]), (t, k)
        assert orc.done[0] == tr["done"][t] and orc.quality_cnt[0] == tr["quality_cnt"][t], t
        if tr["quality_cnt"][t] > 0:
            mean = orc.quality_sum[0] / orc.quality_cnt[0]
            assert abs(mean - float(tr["quality"][t])) <= 1e-12 * max(1.0, abs(mean)), t
        obs = orc.observe(0)
        for k in ("obs_acc", "obs_off", "obs_auc", "ids", "auc_ids"):
            assert np.array_equal(obs[k], np.asarray(tr[k][t]).astype(np.int32)), (t, k)
    if mode == "fix":
        assert e["term_revenue"] == int(tr["term_revenue"][-1])
    return T, int(tr["n_accepted"].sum()), int(tr["n_term"].sum()), int(tr["chain_len"].max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps-scale", type=float, default=1.0)
    ap.add_argument("--out", default=os.path.join(HERE, "LONGRUN_RESULT.json"))
    args = ap.parse_args()
    if not H.reference_available():
        raise SystemExit("reference not present")
    runs, total, t0 = [], 0, time.time()
    for dk, mode, kind, steps, seeds in PLAN:
        dom = DOMS[dk]
        steps = max(10, int(steps * args.steps_scale))
        for seed in seeds:
            rng = np.random.default_rng(seed)
            pol = H.hardcoded_policy() if kind == "hardcoded" else H.random_policy(dom, mode.startswith("free"), rng)
            K = 48
            tr = H.record_trace(dom, mode, steps, pol, agent_kind=kind, K=K, seed=seed)
            T, na, nt, mc = replay(tr, dom, mode, kind, K)
            total += T
            runs.append(dict(domain=dk, N=dom["N"], C=dom["C"], L=dom["L"], mode=mode, agents=kind, seed=seed,
                             steps=T, accepted=na, terminated=nt, max_chain=mc, result="bit-exact"))
            print(f"{dk} {mode:10s} {kind:9s} seed {seed}: {T} steps, {na} accepted, {nt} terminated, "
                  f"max chain {mc}: bit-exact   (total {total}, {time.time() - t0:.0f} s)", flush=True)
    out = dict(total_reference_env_steps=total, result="bit-exact", seconds=round(time.time() - t0, 1),
               compared="state (cores, slots, offers, chains), accepted offers, terminations, all rewards, done, "
                        "acception quality (1e-12), dense observations, offer-ID tables; every step",
               reference=H.REFERENCE_SRC, runs=runs)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("total", total, "reference env-steps: bit-exact ->", args.out)


if __name__ == "__main__":
    main()
