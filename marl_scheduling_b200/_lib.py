"""ctypes binding of include/msched.h (libmsched.so, built in-tree by marl_scheduling_b200/csrc).

There is no CPU fallback: if the shared library is missing this module raises at import of the
symbol table, and every launch fails with MSCHED_E_NODEVICE when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSCHED_LIB") or os.path.join(HERE, "libmsched.so")  # MSCHED_LIB: A/B builds of the same ABI
CSRC = os.path.join(HERE, "csrc")

ABI_VERSION = 2
MAX_KINDS = 16
TILE_ENVS = 128

OK, E_ARG, E_CUDA, E_NODEVICE, E_STATE = 0, -1, -2, -3, -4
FLAG_CHAIN_OVERFLOW, FLAG_COLLECTION_FULL, FLAG_ACTION_RANGE, FLAG_SPAWN_RANGE, FLAG_COMPACT_RANGE = 1, 2, 4, 8, 16

REWARD = {"fix": 0, "divided_fixed": 0, "free_comm": 1, "divided_free_commercial": 1,
          "free_ncomm": 2, "divided_free_noncommercial": 2, "agg": 3, "aggregated_fixed": 3}
AUCTION = {"external": 0, "first": 1, "first_max": 1, "random": 2, "random_max": 2}
SPAWN = {"philox": 0, "kinds": 1, "u64": 2}


class MschedConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("B", C.c_int32),
        ("N", C.c_int32), ("C", C.c_int32), ("L", C.c_int32), ("J", C.c_int32),
        ("newJobsPerRound", C.c_int32), ("rewardMultiplier", C.c_int32),
        ("episodeLength", C.c_int32), ("freePrices", C.c_int32), ("rewardVariant", C.c_int32),
        ("chainCapacity", C.c_int32), ("auctionMode", C.c_int32), ("spawnMode", C.c_int32),
        ("prio", C.c_int32 * MAX_KINDS), ("len", C.c_int32 * MAX_KINDS),
        ("fixPrice", C.c_int32 * MAX_KINDS), ("cumProb", C.c_double * MAX_KINDS),
        ("netZeroOfferReward", C.c_double), ("seed", C.c_uint64), ("envOffset", C.c_int64),
    ]


class MschedLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "padded_envs", "state_words", "action_halfs", "result_words", "obs_halfs", "ids_halfs",
        "chain_words",
        "a_acceptor", "a_offer_core", "a_offer_price", "a_auctioneer", "a_spawn_kind",
        "r_offer", "r_price", "r_acceptor", "r_auctioneer", "r_agent", "r_quality", "r_counts",
        "r_flags", "r_auctioneer_idx", "RL", "RC",
        "o_acceptor", "o_offer", "o_auctioneer", "o_acc_row", "o_off_row",
        "cobs_halfs", "c_core", "c_slot", "c_offer")]


class MschedCompactResultLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("words", "c_offer", "c_price", "c_acceptor", "c_auctioneer", "c_agent",
                                         "c_quality", "c_counts", "c_flags")]


class MschedInfo(C.Structure):
    _fields_ = [("step_impl", C.c_int32), ("fuses_observations", C.c_int32), ("envs_per_cta", C.c_int32),
                ("threads_per_cta", C.c_int32), ("smem_bytes_per_cta", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class MschedMlpGroup(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_hidden", C.c_int32), ("n_actions", C.c_int32),
                ("n_nets", C.c_int32), ("unit_div", C.c_int32), ("reserved", C.c_int32),
                ("weights", C.c_void_p)]


class MschedActorIO(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_stride", C.c_int32), ("units", C.c_int32),
                ("n_envs", C.c_int32), ("n_cores", C.c_int32), ("env_stride", C.c_int64),
                ("row_offset", C.c_int64), ("seed", C.c_uint64), ("step", C.c_uint64),
                ("u_override", C.c_void_p), ("action", C.c_void_p), ("logprob", C.c_void_p),
                ("probs", C.c_void_p), ("action_rec", C.c_void_p), ("action_rec_stride", C.c_int64),
                ("gather_core", C.c_void_p), ("x_used", C.c_void_p), ("timeline", C.c_void_p), ("step_dev", C.c_void_p)]


class MschedPolicyGroup(C.Structure):
    _fields_ = [("nets", MschedMlpGroup), ("units", C.c_int32), ("x_offset", C.c_int32), ("x_stride", C.c_int32),
                ("rec_offset", C.c_int32), ("seed", C.c_uint64), ("action", C.c_void_p), ("logprob", C.c_void_p),
                ("x_used", C.c_void_p), ("x_used_stride", C.c_int32), ("reserved", C.c_int32),
                ("u_override", C.c_void_p), ("probs", C.c_void_p)]


class MschedPolicyStep(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("obs_stride", C.c_int64), ("n_envs", C.c_int32), ("n_cores", C.c_int32),
                ("action_rec", C.c_void_p), ("action_rec_stride", C.c_int64), ("env_offset", C.c_int64),
                ("step", C.c_uint64), ("step_dev", C.c_void_p),
                ("acceptor", MschedPolicyGroup), ("core", MschedPolicyGroup), ("price", MschedPolicyGroup),
                ("input_bound", C.c_int32), ("reserved", C.c_int32)]


class MschedDqnBatch(C.Structure):
    _fields_ = [("policy", C.c_void_p), ("target", C.c_void_p), ("n_in", C.c_int32), ("n_hidden", C.c_int32),
                ("n_actions", C.c_int32), ("n_nets", C.c_int32), ("batch", C.c_int32), ("reserved", C.c_int32),
                ("state", C.c_void_p), ("next_state", C.c_void_p), ("action", C.c_void_p), ("reward", C.c_void_p),
                ("gamma", C.c_float), ("reserved2", C.c_float), ("grad", C.c_void_p), ("loss", C.c_void_p)]


class MschedPpoBatch(C.Structure):
    _fields_ = [("actor_weights", C.c_void_p), ("critic_weights", C.c_void_p),
                ("n_in", C.c_int32), ("n_hidden", C.c_int32), ("n_actions", C.c_int32), ("n_nets", C.c_int32),
                ("x", C.c_void_p), ("x_tb_stride", C.c_int64), ("x_unit_stride", C.c_int64),
                ("action", C.c_void_p), ("logprob_old", C.c_void_p), ("returns", C.c_void_p),
                ("n_tb", C.c_int64),
                ("units", C.c_int32), ("n_sel", C.c_int32), ("units_per_net", C.c_int32), ("reserved", C.c_int32),
                ("net_ids", C.c_void_p), ("unit_ids", C.c_void_p),
                ("eps_clip", C.c_float), ("entropy_coef", C.c_float), ("value_coef", C.c_float), ("reserved2", C.c_float),
                ("grad_actor", C.c_void_p), ("grad_critic", C.c_void_p), ("stats", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64)]


# every symbol include/msched.h declares: name -> (restype, argtypes)
P = C.c_void_p
SYMBOLS = {
    "msched_abi_version": (C.c_int, []),
    "msched_last_error": (C.c_char_p, []),
    "msched_padded_envs": (C.c_int, [C.c_int]),
    "msched_get_layout": (C.c_int, [C.POINTER(MschedConfig), C.POINTER(MschedLayout)]),
    "msched_create": (C.c_int, [C.POINTER(MschedConfig), C.c_int, C.POINTER(P)]),
    "msched_destroy": (C.c_int, [P]),
    "msched_get_info": (C.c_int, [P, C.POINTER(MschedInfo)]),
    "msched_debug_timeline": (C.c_int, [P, P]),
    "msched_bind_state": (C.c_int, [P, P, P]),
    "msched_bind_stats": (C.c_int, [P, P]),
    "msched_stats_sums": (C.c_int, [P, P, P]),
    "msched_result_sums": (C.c_int, [P, P, P, P]),
    "msched_reset": (C.c_int, [P, P]),
    "msched_get_round": (C.c_int, [P, C.POINTER(C.c_int64)]),
    "msched_set_round": (C.c_int, [P, C.c_int64]),
    "msched_set_round_mode": (C.c_int, [P, C.c_int, P]),
    "msched_step": (C.c_int, [P, P, P, P, P]),
    "msched_step_observe": (C.c_int, [P, P, P, P, P, P]),
    "msched_step_multi": (C.c_int, [P, P, C.c_int, P, P, C.c_int, P]),
    "msched_rollout_hardcoded": (C.c_int, [P, P, C.c_int, P, P, C.c_int, C.c_int, P]),
    "msched_step_host": (C.c_int, [P, P, P, P, P]),
    "msched_get_compact_result_layout": (C.c_int, [C.POINTER(MschedConfig), C.POINTER(MschedCompactResultLayout)]),
    "msched_step_host_compact": (C.c_int, [P, P, P, P, P]),
    "msched_observe_dense": (C.c_int, [P, P, P, P]),
    "msched_observe_compact": (C.c_int, [P, P, P]),
    "msched_step_compact": (C.c_int, [P, P, P, P, P, P]),
    "msched_auctioneer_action": (C.c_int, [P, C.c_int, P, P]),
    "msched_hardcoded_actions": (C.c_int, [P, P, C.c_int, P, P, P, P]),
    "msched_export_state": (C.c_int, [P, C.c_int, C.c_int, P, P, P, P, P, P, P]),
    "msched_mlp_param_count": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "msched_actor_forward": (C.c_int, [C.POINTER(MschedMlpGroup), C.POINTER(MschedActorIO), P]),
    "msched_policy_step": (C.c_int, [C.POINTER(MschedPolicyStep), P]),
    "msched_dqn_param_count": (C.c_int, [C.c_int, C.c_int]),
    "msched_dqn_select": (C.c_int, [C.POINTER(MschedMlpGroup), C.POINTER(MschedActorIO), C.c_float, P, P]),
    "msched_dqn_grad": (C.c_int, [C.POINTER(MschedDqnBatch), P]),
    "msched_ppo_workspace_bytes": (C.c_int, [C.POINTER(MschedPpoBatch), C.POINTER(C.c_uint64)]),
    "msched_ppo_grad": (C.c_int, [C.POINTER(MschedPpoBatch), P]),
    "msched_adam_step": (C.c_int, [P, P, P, P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, P]),
    "msched_returns": (C.c_int, [P, C.c_int, C.c_int, C.c_double, C.c_int, P, P]),
}


class MschedError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libmsched.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]))
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library; raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MschedError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the header disagree
            fn.restype = res
            fn.argtypes = args
        if L.msched_abi_version() != ABI_VERSION:
            raise MschedError("libmsched.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        msg = lib().msched_last_error().decode(errors="replace")
        raise MschedError(f"msched error {rc}: {msg}")


def cum_prob(probabilities):
    """World.accProbabilities (src/world.py:220-222): Python float prefix sums."""
    probabilities = list(probabilities)
    return [sum(probabilities[: i + 1]) for i in range(len(probabilities))]


def make_config(B, world_params, reward="fix", auction="external", spawn="philox",
                chain_capacity=64, seed=0, env_offset=0, net_zero_offer_reward=0.5):
    """world_params uses the reference's World(params) keys (src/world.py:211-246)."""
    wp = world_params
    prios = list(wp["possibleJobPriorities"])
    lens = list(wp["possibleJobLengths"])
    J = len(prios)
    if len(lens) != J or len(wp["probabilities"]) != J or J > MAX_KINDS:
        raise ValueError("possibleJobPriorities/Lengths/probabilities must have equal length <= 16")
    cfg = MschedConfig()
    cfg.abi_version = ABI_VERSION
    cfg.B = int(B)
    cfg.N, cfg.C, cfg.L, cfg.J = (int(wp["numberOfAgents"]), int(wp["numberOfCores"]),
                                  int(wp["collectionLength"]), J)
    cfg.newJobsPerRound = int(wp["newJobsPerRoundPerAgent"])
    cfg.rewardMultiplier = int(wp["rewardMultiplier"])
    cfg.episodeLength = int(wp["episodeLength"])
    cfg.freePrices = int(bool(wp["freePrices"]))
    cfg.rewardVariant = REWARD[reward]
    cfg.chainCapacity = int(chain_capacity)
    cfg.auctionMode = AUCTION[auction]
    cfg.spawnMode = SPAWN[spawn]
    fix = list(wp.get("fixPricesList") or [])
    if not cfg.freePrices and len(fix) < J:
        raise ValueError("fixPricesList needs one price per job kind")
    cp = cum_prob(wp["probabilities"])
    for k in range(J):
        cfg.prio[k], cfg.len[k] = int(prios[k]), int(lens[k])
        cfg.fixPrice[k] = int(fix[k]) if k < len(fix) else 0
        cfg.cumProb[k] = float(cp[k])
    cfg.netZeroOfferReward = float(net_zero_offer_reward)
    cfg.seed = int(seed)
    cfg.envOffset = int(env_offset)
    return cfg


def get_layout(cfg):
    lay = MschedLayout()
    check(lib().msched_get_layout(C.byref(cfg), C.byref(lay)))
    return lay
