"""ctypes wrapper around oracle/libmsched_oracle.so (CPU checker; TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmsched_oracle.so")

MAX_KINDS = 16
MODES = {"fix": 0, "free_comm": 1, "free_ncomm": 2, "agg": 3}
TIE_FIRST, TIE_PHILOX = 0, 1


class MsorConfig(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("L", C.c_int32), ("J", C.c_int32),
        ("newJobs", C.c_int32), ("rewardMultiplier", C.c_int32), ("episodeLength", C.c_int32),
        ("freePrices", C.c_int32), ("rewardMode", C.c_int32), ("chainCap", C.c_int32),
        ("tieMode", C.c_int32),
        ("prio", C.c_int32 * MAX_KINDS), ("len", C.c_int32 * MAX_KINDS),
        ("fix", C.c_int32 * MAX_KINDS),
        ("cumProb", C.c_double * MAX_KINDS),
        ("netZeroOfferReward", C.c_double),
        ("seed", C.c_uint64),
        ("envOffset", C.c_int64),
    ]


def build(force=False):
    src = os.path.join(HERE, "msched_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.msor_create.restype = C.c_void_p
        _lib.msor_create.argtypes = [C.POINTER(MsorConfig)]
        _lib.msor_destroy.argtypes = [C.c_void_p]
        _lib.msor_reset.argtypes = [C.c_void_p]
        _lib.msor_step_range.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 18
        _lib.msor_observe.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        _lib.msor_export.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        _lib.msor_returns.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]
        _lib.msor_mlp_forward.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
                                          + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 4)
        _lib.msor_philox.argtypes = [C.c_void_p] * 3
        _lib.msor_stats.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


def cum_prob(probabilities):
    """World.accProbabilities (src/world.py:220-222): Python float prefix sums."""
    return [sum(probabilities[: i + 1]) for i in range(len(probabilities))]


def make_config(B, dom, mode, chain_cap=64, tie_mode=TIE_FIRST, seed=0, env_offset=0):
    cfg = MsorConfig()
    J = len(dom["prios"])
    cfg.B, cfg.N, cfg.C, cfg.L, cfg.J = B, dom["N"], dom["C"], dom["L"], J
    cfg.newJobs = dom.get("newJobs", 1)
    cfg.rewardMultiplier = dom.get("mult", 1)
    cfg.episodeLength = dom.get("episodeLength", 100)
    cfg.freePrices = int(mode.startswith("free"))
    cfg.rewardMode = MODES[mode]
    cfg.chainCap = chain_cap
    cfg.tieMode = tie_mode
    fix = list(dom.get("fix", [0] * J))
    fix = (fix + [0] * J)[:J]
    cp = cum_prob(list(dom["probs"]))
    for k in range(J):
        cfg.prio[k] = dom["prios"][k]
        cfg.len[k] = dom["lens"][k]
        cfg.fix[k] = fix[k]
        cfg.cumProb[k] = cp[k]
    cfg.netZeroOfferReward = dom.get("netZero", 0.5)
    cfg.seed = seed
    cfg.envOffset = env_offset
    return cfg


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """B independent reference worlds advanced in lock-step on the CPU."""

    def __init__(self, B, dom, mode, chain_cap=64, tie_mode=TIE_FIRST, seed=0, env_offset=0):
        self.cfg = make_config(B, dom, mode, chain_cap, tie_mode, seed, env_offset)
        self.B, self.N, self.C, self.L = B, dom["N"], dom["C"], dom["L"]
        self.NL = self.N * self.L
        self.nj = self.cfg.newJobs
        self.K = chain_cap
        self.agg = mode == "agg"
        self.RL = 1 if self.agg else self.L
        self.RC = 1 if self.agg else self.C
        self.h = lib().msor_create(C.byref(self.cfg))
        if not self.h:
            raise ValueError("bad oracle config")
        lib().msor_reset(self.h)
        B, N, Cc = self.B, self.N, self.C
        self.r_offer = np.zeros((B, N, self.RL), np.float64)
        self.r_price = np.zeros((B, N, self.RL), np.float64)
        self.r_acceptor = np.zeros((B, N, self.RC), np.int64)
        self.r_auctioneer = np.zeros((B, Cc), np.int64)
        self.r_agent = np.zeros((B, N), np.int64)
        self.quality_sum = np.zeros(B, np.float64)
        self.quality_cnt = np.zeros(B, np.int32)
        self.done = np.zeros(B, np.uint8)
        self.auc_out = np.zeros((B, Cc), np.int32)
        self.n_accepted = np.zeros(B, np.int32)
        self.n_term = np.zeros(B, np.int32)
        self.flags = np.zeros(B, np.uint32)

    def __del__(self):
        try:
            if self.h:
                lib().msor_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self):
        lib().msor_reset(self.h)

    def step(self, offc, acc, auc=None, offp=None, spawn_u=None, spawn_kind=None, b0=0, b1=None):
        """Argument order mirrors SchedulingEnv.step(offerActions, acceptorActions, auctioneer)."""
        B = self.B
        b1 = B if b1 is None else b1
        offc = np.ascontiguousarray(offc, np.int32).reshape(B, self.N, self.L)
        acc = np.ascontiguousarray(acc, np.int32).reshape(B, self.N, self.C)
        auc = None if auc is None else np.ascontiguousarray(auc, np.int32).reshape(B, self.C)
        offp = None if offp is None else np.ascontiguousarray(offp, np.int32).reshape(B, self.N, self.L)
        if self.cfg.freePrices and offp is None:
            raise ValueError("free prices need offp")
        spawn_u = None if spawn_u is None else np.ascontiguousarray(spawn_u, np.float64).reshape(
            B, self.N, self.nj)
        spawn_kind = None if spawn_kind is None else np.ascontiguousarray(
            spawn_kind, np.uint8).reshape(B, self.N, self.nj)
        self._keep = (offc, acc, auc, offp, spawn_u, spawn_kind)
        lib().msor_step_range(self.h, b0, b1, _p(offc), _p(offp), _p(acc), _p(auc), _p(spawn_u),
                              _p(spawn_kind), _p(self.r_offer), _p(self.r_price),
                              _p(self.r_acceptor), _p(self.r_auctioneer), _p(self.r_agent),
                              _p(self.quality_sum), _p(self.quality_cnt), _p(self.done),
                              _p(self.auc_out), _p(self.n_accepted), _p(self.n_term),
                              _p(self.flags))

    def observe(self, b):
        N, Cc, L, NL = self.N, self.C, self.L, self.NL
        W = 3 + 2 * NL
        o = dict(obs_acc=np.zeros((N, Cc, W), np.int32), obs_off=np.zeros((N, L, 2 * Cc + 2), np.int32),
                 obs_auc=np.zeros((Cc, W), np.int32), ids=np.zeros((N, Cc, NL), np.int32),
                 auc_ids=np.zeros((Cc, NL), np.int32))
        lib().msor_observe(self.h, b, _p(o["obs_acc"]), _p(o["obs_off"]), _p(o["obs_auc"]),
                           _p(o["ids"]), _p(o["auc_ids"]))
        return o

    def stats(self):
        """Episode statistics [B][J][4]: per job kind (sum of accepted prices, #accepted, sum of
        (dwell - 1), #terminated) since the last reset."""
        out = np.zeros((self.cfg.B, self.cfg.J, 4), np.int32)
        lib().msor_stats(self.h, _p(out))
        return out

    def export(self, b):
        N, Cc, L, NL, K = self.N, self.C, self.L, self.NL, self.K
        core = np.zeros((Cc, 7), np.int32)
        slot = np.zeros((NL, 7), np.int32)
        off = np.zeros((NL, 5), np.int32)
        chain = np.zeros((Cc, K, 5), np.int32)
        clen = np.zeros(Cc, np.int32)
        accepted = np.zeros((Cc, 5), np.int32)
        term = np.zeros((Cc, 8), np.int32)
        tnorm = np.zeros(Cc, np.float64)
        misc = np.zeros(6, np.int32)
        lib().msor_export(self.h, b, _p(core), _p(slot), _p(off), _p(chain), _p(clen),
                          _p(accepted), _p(term), _p(tnorm), _p(misc))
        s = slot.reshape(N, L, 7)
        f = off.reshape(N, L, 5)
        return dict(
            core_owner=core[:, 0], core_prio=core[:, 1], core_rem=core[:, 2], core_jobid=core[:, 3],
            core_kind=core[:, 4], core_birth=core[:, 5], core_init=core[:, 6],
            slot_prio=s[..., 0], slot_rem=s[..., 1], slot_jobid=s[..., 2], slot_kind=s[..., 3],
            slot_wait=s[..., 4], slot_birth=s[..., 5], slot_init=s[..., 6],
            off_core=f[..., 0], off_recip=f[..., 1], off_price=f[..., 2], off_time=f[..., 3],
            off_id=f[..., 4], chain=chain, chain_len=clen, accepted=accepted, term=term,
            term_norm=tnorm, round=int(misc[0]), job_counter=int(misc[1]),
            n_accepted=int(misc[2]), n_term=int(misc[3]), flags=int(misc[4]),
            term_revenue=int(misc[5]))


def returns(rewards, gamma, normalise=True):
    r = np.ascontiguousarray(rewards, np.float64)
    T, M = r.shape
    out = np.zeros((T, M), np.float32)
    lib().msor_returns(_p(r), T, M, float(gamma), int(normalise), _p(out))
    return out


def mlp_forward(x, W1, b1, W2, b2, W3, b3, softmax=True, u=None):
    x = np.ascontiguousarray(x, np.float32)
    M, nin = x.shape
    h = W1.shape[0]
    A = W3.shape[0]
    ws = [np.ascontiguousarray(w, np.float32) for w in (W1, b1, W2, b2, W3, b3)]
    out = np.zeros((M, A), np.float32)
    act = lp = None
    if u is not None:
        u = np.ascontiguousarray(u, np.float32)
        act = np.zeros(M, np.int32)
        lp = np.zeros(M, np.float32)
    lib().msor_mlp_forward(_p(x), M, nin, h, A, *[_p(w) for w in ws], int(softmax), _p(out),
                           _p(u), _p(act), _p(lp))
    return out, act, lp


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, np.uint32)
    k = np.ascontiguousarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().msor_philox(_p(c), _p(k), _p(o))
    return o


# ---- PPO.update, numpy float64 restatement with a hand-written backward sweep --------
def _unpack_net(flat, n_in, H, A):
    o, out = 0, []
    for shape in ((H, n_in), (H,), (H, H), (H,), (A, H), (A,)):
        n = int(np.prod(shape))
        out.append(np.asarray(flat[o:o + n], np.float64).reshape(shape))
        o += n
    return out


def _mlp_fwd_bwd(flat, x, dout_fn, n_in, H, A):
    """Linear-Tanh-Linear-Tanh-Linear (src/PPOmodules.py:32-50); dout_fn(z) -> dL/dz; returns (z, flat grad)."""
    W1, b1, W2, b2, W3, b3 = _unpack_net(flat, n_in, H, A)
    h1 = np.tanh(x @ W1.T + b1)
    h2 = np.tanh(h1 @ W2.T + b2)
    z = h2 @ W3.T + b3
    dz = dout_fn(z)
    da2 = (dz @ W3) * (1 - h2 * h2)
    da1 = (da2 @ W2) * (1 - h1 * h1)
    g = [da1.T @ x, da1.sum(0), da2.T @ h1, da2.sum(0), dz.T @ h2, dz.sum(0)]
    return z, np.concatenate([a.reshape(-1) for a in g])


def ppo_loss_grads(actor, critic, x, action, logp_old, G, eps_clip, n_hidden=16):
    """Gradient of loss.mean() of src/PPOmodules.py:139-174 for one ActorCritic over its samples:
    evaluate (:65-72), ratios, surr1/surr2 with torch.min / torch.clamp subgradients, 0.5*MseLoss (a batch
    mean), -0.01*entropy.  Returns (grad actor, grad critic, [mean -surr, mean mse, mean entropy])."""
    x = np.asarray(x, np.float64)
    M, n_in = x.shape
    H = n_hidden
    A = (len(actor) - (H * n_in + H + H * H + H)) // (H + 1)
    G = np.asarray(G, np.float64)
    vbox = {}

    def dcritic(z):
        vbox["v"] = z[:, 0]
        return ((z[:, 0] - G) / M)[:, None]          # d/dv of 0.5*mean((v-G)^2)
    _, gc = _mlp_fwd_bwd(critic, x, dcritic, n_in, H, 1)
    v = vbox["v"]
    sbox = {}

    def dactor(z):
        z = z - z.max(1, keepdims=True)
        logp = z - np.log(np.exp(z).sum(1, keepdims=True))
        p = np.exp(logp)
        ent = -(p * logp).sum(1)
        lpa = logp[np.arange(M), action]
        ratio = np.exp(lpa - logp_old)
        adv = G - v
        s1, s2 = ratio * adv, np.clip(ratio, 1 - eps_clip, 1 + eps_clip) * adv
        inr = (ratio >= 1 - eps_clip) & (ratio <= 1 + eps_clip)
        d = np.where((s1 < s2) | inr, adv, 0.0)
        onehot = np.zeros_like(p)
        onehot[np.arange(M), action] = 1.0
        sbox["s"] = [float((-np.minimum(s1, s2)).mean()), float(((v - G) ** 2).mean()), float(ent.mean())]
        return ((-d * ratio)[:, None] * (onehot - p) + 0.01 * p * (logp + ent[:, None])) / M
    _, ga = _mlp_fwd_bwd(actor, x, dactor, n_in, H, A)
    return ga, gc, sbox["s"]


def adam_step(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (src/PPOmodules.py:100-105), float64."""
    m[:] = m + (g - m) * (1 - b1)
    v[:] = v * b2 + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p -= (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)


def ppo_update(actor0, critic0, states, actions, logp_old, rewards, gamma, eps_clip, K, lr_actor, lr_critic):
    """PPO.update (src/PPOmodules.py:127-174) for one world: returns, K epochs, two-learning-rate Adam."""
    T = len(rewards)
    G = returns(np.asarray(rewards, np.float64).reshape(T, 1), gamma, True)[:, 0].astype(np.float64)
    a, c = np.asarray(actor0, np.float64).copy(), np.asarray(critic0, np.float64).copy()
    ma, va, mc, vc = (np.zeros_like(t) for t in (a, a, c, c))
    for k in range(1, K + 1):
        ga, gc, _ = ppo_loss_grads(a, c, states, actions, logp_old, G, eps_clip)
        adam_step(a, ga, ma, va, lr_actor, k)
        adam_step(c, gc, mc, vc, lr_critic, k)
    return a, c


# ---- hard-coded agents (BASELINE config 1), restated from the reference ---------------------------------------
def _hc_ratio(a, b):
    """calculateRewardRatio, src/HardcodedModules.py:5-13, as an exact Fraction."""
    from fractions import Fraction
    if a in (-1, -2) or b in (-1, -2):
        return Fraction(-1)
    return Fraction(int(a), int(b))


def hardcoded_actions(obs_acc, obs_off, u_acc=None, u_off=None, cands_out=None):
    """DividedHardcodedAgent.getActions for all agents of ONE world (src/Agent.py:622-641):
    HardcodedAcceptor.selectAction per (agent, core) (src/HardcodedModules.py:16-45) and
    HardcodedOfferer.selectAction per (agent, slot) (:81-109) on the dense observations
    obs_acc [N,C,3+2NL], obs_off [N,L,2C+2].  The reference breaks ties with random.sample; here candidate
    floor(u * #candidates) in index order is taken (u = 0 -> the first), the contract of the CUDA kernel.
    Returns (acc [N,C], off [N,L], ncand_acc [N,C], ncand_off [N,L]); cands_out (a dict) receives the candidate
    index lists under ("acc", i, j) / ("off", i, q)."""
    obs_acc, obs_off = np.asarray(obs_acc), np.asarray(obs_off)
    N, Cc, W = obs_acc.shape
    L = obs_off.shape[1]
    NL = (W - 3) // 2
    acc = np.full((N, Cc), NL, np.int32)
    off = np.zeros((N, L), np.int32)
    nca = np.zeros((N, Cc), np.int32)
    nco = np.zeros((N, L), np.int32)

    def pick(cands, u):
        k = int(np.float32(u) * np.float32(len(cands)))
        return cands[min(k, len(cands) - 1)]
    for i in range(N):
        for j in range(Cc):
            row = obs_acc[i, j].tolist()
            if row[0] == 0:
                continue
            own = _hc_ratio(row[1], row[2])
            ratios = [_hc_ratio(p, t) for p, t in zip(row[3::2], row[4::2])]
            mx = max(ratios)
            if mx > own:
                cands = [k for k, r in enumerate(ratios) if r == mx]
                nca[i, j] = len(cands)
                if cands_out is not None:
                    cands_out[("acc", i, j)] = cands
                acc[i, j] = pick(cands, 0.0 if u_acc is None else u_acc[i][j])
        for q in range(L):
            row = obs_off[i, q].tolist()
            ratios = [_hc_ratio(p, t) for p, t in zip(row[:-2:2], row[1:-2:2])]
            mn = min(ratios)
            cands = [k for k, r in enumerate(ratios) if r == mn]
            nco[i, q] = len(cands)
            if cands_out is not None:
                cands_out[("off", i, q)] = cands
            off[i, q] = pick(cands, 0.0 if u_off is None else u_off[i][q])
    return acc, off, nca, nco


def hardcoded_draws(seed, env, round_, n_units):
    """The u of every unit of one env and round as the CUDA kernel draws them: Philox counter (env lo, env hi,
    round, 3 << 28 | call << 12), key = seed; unit n uses word n % 4 of call n // 4, u = (x >> 8) * 2^-24."""
    out = np.zeros(n_units, np.float32)
    for call in range((n_units + 3) // 4):
        x = philox([env & 0xFFFFFFFF, (env >> 32) & 0xFFFFFFFF, round_ & 0xFFFFFFFF, (3 << 28) | (call << 12)],
                   [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
        for w in range(4):
            if 4 * call + w < n_units:
                out[4 * call + w] = np.float32(int(x[w]) >> 8) * np.float32(1.0 / 16777216.0)
    return out


def compact_observation(e):
    """The compact observation record (include/msched.h: MschedLayout.c_core / c_slot / c_offer) of one world from
    its exported state `e` (Oracle.export): core [C,4] = ownerID, priority, remainingLength, jobKind; slot [N,L,2] =
    priority, remainingLength; offer [N,L,2] = coreID (0 = none), offeredReward.  The same fields the dense rows of
    src/Agent.py:167-300 and src/Auctioneer.py:34-77 are built from."""
    core = np.stack([e["core_owner"], e["core_prio"], e["core_rem"], e["core_kind"]], -1).astype(np.int32)
    slot = np.stack([e["slot_prio"], e["slot_rem"]], -1).astype(np.int32)
    has = np.asarray(e["off_core"]) > 0
    offer = np.stack([np.where(has, e["off_core"], 0), np.where(has, e["off_price"], 0)], -1).astype(np.int32)
    return dict(core=core, slot=slot, offer=offer)


# ---- DQN: numpy float64 restatement of src/DQNmodules.py ---------------------------------------
def _dqn_unpack(flat, n_in, A, H=16):
    o, out = 0, []
    for shape in ((H, n_in), (H,), (A, H), (A,)):
        n = int(np.prod(shape))
        out.append(np.asarray(flat[o:o + n], np.float64).reshape(shape))
        o += n
    return out


def dqn_forward(flat, x, n_actions):
    """DQNEntity.forward (src/DQNmodules.py:41-54): Linear(in,16)-Tanh-Linear(16,A); returns (Q [M,A], hidden [M,16])."""
    x = np.asarray(x, np.float64)
    W1, b1, W2, b2 = _dqn_unpack(flat, x.shape[1], n_actions)
    h = np.tanh(x @ W1.T + b1)
    return h @ W2.T + b2, h


def dqn_select(flat, x, n_actions, sample, rand_action, eps):
    """DQNEntity.selectAction (src/DQNmodules.py:56-76): `sample > eps_treshold` exploits (first arg-max of Q), else
    the given random action."""
    q, _ = dqn_forward(flat, x, n_actions)
    return np.where(np.asarray(sample) > eps, q.argmax(1), np.asarray(rand_action)).astype(np.int32)


def dqn_grad(policy, target, S, A_, S2, R, gamma, n_actions):
    """Gradient of optimize_model's loss (src/DQNmodules.py:97-154) w.r.t. the policy net, AFTER the clamp to
    [-1, 1]: SmoothL1Loss (beta 1, mean over the batch) between Q_policy(s)[a] and gamma * max_a' Q_target(s') + r."""
    S, S2 = np.asarray(S, np.float64), np.asarray(S2, np.float64)
    M, n_in = S.shape
    W1, b1, W2, b2 = _dqn_unpack(policy, n_in, n_actions)
    q, h = dqn_forward(policy, S, n_actions)
    qn, _ = dqn_forward(target, S2, n_actions)
    expected = qn.max(1) * gamma + np.asarray(R, np.float64)
    A_ = np.asarray(A_).astype(np.int64)
    d = q[np.arange(M), A_] - expected
    g = np.clip(d, -1.0, 1.0) / M                       # d/dq of mean Huber
    dq = np.zeros_like(q)
    dq[np.arange(M), A_] = g
    dh = dq @ W2
    dz = dh * (1 - h * h)
    grads = [dz.T @ S, dz.sum(0), dq.T @ h, dq.sum(0)]
    flat = np.concatenate([a.reshape(-1) for a in grads])
    loss = float(np.where(np.abs(d) < 1, 0.5 * d * d, np.abs(d) - 0.5).mean())
    return np.clip(flat, -1.0, 1.0), loss


def dqn_optimize(policy0, target, S, A_, S2, R, idx_steps, gamma, n_actions, lr=1e-3):
    """optimize_model for the recorded batches `idx_steps` [steps, batch]; returns the parameters after every step."""
    w = np.asarray(policy0, np.float64).copy()
    m, v = np.zeros_like(w), np.zeros_like(w)
    out = []
    for k, idx in enumerate(idx_steps, 1):
        g, _ = dqn_grad(w, target, S[idx], A_[idx], S2[idx], R[idx], gamma, n_actions)
        adam_step(w, g, m, v, lr, k)
        out.append(w.copy())
    return np.stack(out)
