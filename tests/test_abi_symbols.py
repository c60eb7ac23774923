"""CPU: the C-ABI library loads, exports every symbol include/msched.h declares, validates
configurations, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from marl_scheduling_b200 import _lib as L
from marl_scheduling_b200.batched_env import world_params_from_dom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOM = dict(N=2, C=3, L=3, prios=[2, 4, 8], lens=[5, 5, 5], probs=[1 / 3] * 3, fix=[1])


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "msched.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msched_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert names, "no declarations parsed"
    assert sorted(L.SYMBOLS) == names
    lib = C.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libmsched.so does not export {n}"
    assert L.lib().msched_abi_version() == L.ABI_VERSION


def test_layout_sizes_and_bank_rules():
    for B, dom, mode in [(65536, DOM, "free_comm"),
                         (4096, dict(N=4, C=4, L=3, prios=[3, 10], lens=[6, 3], probs=[.8, .2], fix=[2, 7]), "fix"),
                         (1, dict(N=1, C=1, L=1, prios=[3], lens=[2], probs=[1], fix=[2]), "agg")]:
        cfg = L.make_config(B, world_params_from_dom(dom, mode.startswith("free")), reward=mode,
                            auction="external", spawn="kinds")
        lay = L.get_layout(cfg)
        assert lay.padded_envs % L.TILE_ENVS == 0 and lay.padded_envs >= B
        assert lay.state_words % 2 == 1 and lay.result_words % 2 == 1
        assert lay.action_halfs % 2 == 0 and (lay.action_halfs // 2) % 2 == 1
        assert lay.obs_halfs % 2 == 0 and (lay.obs_halfs // 2) % 2 == 1
        N, Cc, Lc = dom["N"], dom["C"], dom["L"]
        assert lay.state_words >= 2 + 3 * Cc + 4 * N * Lc
        assert lay.a_auctioneer >= 0 and lay.a_spawn_kind >= 0
        assert (lay.a_offer_price >= 0) == mode.startswith("free")


@pytest.mark.parametrize("bad", [dict(N=0), dict(C=65), dict(lens=[300, 5, 5]), dict(newJobs=9)])
def test_bad_configs_are_rejected(bad):
    dom = dict(DOM, **bad)
    cfg = L.make_config(8, world_params_from_dom(dom, True), reward="free_comm")
    with pytest.raises(L.MschedError):
        L.get_layout(cfg)


def test_reward_variant_must_match_free_prices():
    cfg = L.make_config(8, world_params_from_dom(DOM, True), reward="fix")
    with pytest.raises(L.MschedError):
        L.get_layout(cfg)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = L.make_config(8, world_params_from_dom(DOM, True), reward="free_comm")
    h = C.c_void_p()
    rc = L.lib().msched_create(C.byref(cfg), 0, C.byref(h))
    assert rc == L.E_NODEVICE and not h.value
    from marl_scheduling_b200.batched_env import BatchedSchedulingEnv
    with pytest.raises(L.MschedError):
        BatchedSchedulingEnv(8, world_params_from_dom(DOM, True), reward="free_comm")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "marl_scheduling_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            src = open(os.path.join(dp, f)).read()
            assert "oracle" not in src.replace("oracle/", ""), os.path.join(dp, f)
