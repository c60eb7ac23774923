"""CPU: the result format keeps the reference's argsDict keys (src/trainPPO.py:229-243)."""
import pickle

from marl_scheduling_b200.results import KEYS, ResultLog


def test_result_log_has_reference_keys_and_pickles(tmp_path):
    log = ResultLog({"numberOfAgents": 2, "is_PPO": True}, plot_path="p", mean_job_fraction=0.4)
    for e in range(3):
        log.append({k: ([1.0, None] if k in ("prices", "dwellTimes") else float(e)) for k in KEYS})
    ref_keys = {"plotPath", "acceptorRew", "coreChooserRew", "priceChooserRew", "prices", "auctioneerRew",
                "dwellTimes", "meanJob", "agentRew", "acceptionQuality", "acceptionAmount",
                "terminationRevenues", "tradeRevenues", "params"}
    assert set(log.argsDict()) == ref_keys
    p0 = log.dump(str(tmp_path / "data{}.pkl"))
    p1 = log.dump(str(tmp_path / "data{}.pkl"))
    assert p0.endswith("data0.pkl") and p1.endswith("data1.pkl")
    d = pickle.load(open(p0, "rb"))
    assert d["acceptorRew"] == [0.0, 1.0, 2.0] and d["params"]["is_PPO"] is True and d["meanJob"] == 0.4
