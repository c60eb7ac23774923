"""Shared helpers for the parity tests: golden-trace loading and state comparison."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STATE_KEYS = ["core_owner", "core_prio", "core_rem", "core_jobid", "core_kind", "core_birth",
              "core_init", "slot_prio", "slot_rem", "slot_jobid", "slot_kind", "slot_wait",
              "slot_birth", "slot_init", "off_core", "off_recip", "off_price", "off_time", "off_id",
              "chain_len"]


def golden_names(kind="divided"):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        n = os.path.basename(f)[:-4]
        if n in ("torch_vectors", "ppo_update", "dqn"):
            continue
        if kind == "divided" and n.startswith("aggobs"):
            continue
        out.append(n)
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    tr = {k: z[k] for k in z.files}
    meta = json.loads(bytes(tr.pop("meta")).decode())
    return tr, meta


def assert_state_equal(exp, tr, t, K=None, prefix=""):
    """exp: dict from Oracle.export / device export; tr: golden arrays; t: step index (or None
    for the init_ snapshot)."""
    def g(k):
        return tr["init_" + k] if t is None else tr[k][t]
    for k in STATE_KEYS:
        a, b = np.asarray(exp[k]).astype(np.int64), np.asarray(g(k)).astype(np.int64)
        assert a.shape == b.shape, (prefix, k, a.shape, b.shape)
        assert np.array_equal(a, b), f"{prefix} step {t}: {k} differs\n got {a}\n ref {b}"
    ch_ref = np.asarray(g("chain")).astype(np.int64)
    ch = np.asarray(exp["chain"]).astype(np.int64)
    kk = min(ch.shape[1], ch_ref.shape[1])
    assert ch_ref[:, kk:].size == 0 or (ch_ref[:, kk:] == -1).all()
    assert np.array_equal(ch[:, :kk], ch_ref[:, :kk]), f"{prefix} step {t}: chain differs"
