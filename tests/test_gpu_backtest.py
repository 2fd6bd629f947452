"""Evaluation phase of src/main.cpp:216-241 on the GPU: after training, Agent::GoGreedy, a NEW Intraday object
per env and experiment::serial::Backtester (serial.cpp:18-34,121-137) -- against the reference's own records
(tests/golden/steps_bt_*), bitwise, plus the profit_log / test_stats writers."""
import ctypes as C
import os

import pytest

import golden_util as G
from rl_markets_b200 import abi, backtest, config

pytestmark = pytest.mark.gpu


def _train_then_backtest(rlm, case, n_envs=3):
    cfg = G.case_config(case, n_envs=n_envs, env_index0=case["env"])
    cfg.flow.t0_ms = case["t0_ms"]
    cfg.record_envs, cfg.record_cap = 1, 600
    m = rlm.BatchedMarket(cfg)
    m.run_ticks(case["ticks"])
    m.sync()
    assert all(s.terminal for s in m.stats())
    train_recs, _k = m.records(0)
    m.handle_terminal(0)              # Learner::RunEpisode, serial.cpp:79
    m.go_greedy()                     # main.cpp:217
    m.set_mode(abi.MODE_BACKTEST)     # experiment::serial::Backtester
    t = case["test"]
    flow = config.from_dict(case["yaml"], flow_seed=t["flow_seed"]).flow
    flow.t0_ms = t["t0_ms"]
    m.new_env(flow)                   # main.cpp:219-222: new Intraday + LoadData of the test day
    theta_before = bytes(m.theta(0))
    m.run_ticks(t["ticks"])
    m.sync()
    return m, train_recs, theta_before


def test_backtest_matches_the_reference(rlm):
    for case in G.backtest_manifest():
        m, train_recs, theta_before = _train_then_backtest(rlm, case)
        gold, _k = G.records(case["name"])
        assert len(train_recs) == len(gold)
        for i, g in enumerate(gold):
            assert not abi.record_fields_equal(g, train_recs[i]), (case["name"], "train", i)
        recs, _k2 = m.records(0)
        gold2, _k3 = G.records(case["name"] + "_test")
        assert len(recs) == len(gold2) > 100, (case["name"], len(recs), len(gold2))
        for i, g in enumerate(gold2):
            bad = abi.record_fields_equal(g, recs[i])
            assert not bad, "%s evaluation step %d (reference, cuda): %r" % (case["name"], i, G.describe_diff(g, recs[i], bad))
        assert bytes(m.theta(0)) == theta_before, "backtest must not touch theta"
        st = m.stats()[0]
        s = case["summary"]  # after Runner::RunEpisode's ClearInventory
        assert st.terminal == 1
        assert (st.position, st.episode_pnl, st.episode_reward, st.ask_transactions, st.bid_transactions, st.market_buys,
                st.market_sells) == (s["test_position"], s["test_ep_pnl"], s["test_ep_reward"], s["test_ask_tx"],
                                     s["test_bid_tx"], s["test_market_buys"], s["test_market_sells"])
        m.close()


def test_backtest_logs(rlm, tmp_path):
    case = G.backtest_manifest()[0]
    m, _tr, _th = _train_then_backtest(rlm, case)
    out = backtest.write_logs(m, str(tmp_path), env=0, date=20100104)
    rows = open(out["profit_log"]).read().splitlines()
    assert rows[0] == "episode,step,action,position,midprice,spread,quoted_ask,quoted_bid,ask_level,bid_level,pnl_step,bandh_step"
    recs, _k = m.records(0)
    assert len(rows) == 1 + len(recs)
    first = rows[1].split(",")
    assert first[0] == "20100104" and int(first[2]) == recs[0].action and float(first[4]) == recs[0].midprice
    stats = dict(l.split(",") for l in open(out["test_stats"]).read().splitlines())
    st = m.stats()[0]
    # Base::writeStats (base.cpp:451-456) reopens the file for each writer: only TradeStatistics survives
    assert list(stats) == ["asks_placed", "bids_placed", "asks_cancelled", "bids_cancelled", "ask_transactions",
                           "bid_transactions", "market_sells", "market_buys"]
    assert int(stats["ask_transactions"]) == recs[-2].ask_transactions and int(stats["market_buys"]) == st.market_buys
    # and the files are, byte for byte, what the reference itself wrote for this case (tools/make_golden.py)
    assert open(out["profit_log"]).read() == open(os.path.join(G.GOLD, case["name"] + "_profit_log.csv")).read()
    assert open(out["test_stats"]).read() == open(os.path.join(G.GOLD, case["name"] + "_test_stats.csv")).read()
    assert os.path.getsize(out["theta"]) == 8 * m.cfg.memory_size
    m.close()


def test_backtest_mode_needs_the_synchronous_engine(rlm, monkeypatch):
    monkeypatch.setenv("RLM_ENGINE", "f")
    cfg = config.from_dict(config.example_dict(**{"learning.memory_size": 4096, "learning.algorithm": "q_learn"}), n_envs=2)
    m = rlm.BatchedMarket(cfg)
    with pytest.raises(rlm.RlmError) as ei:
        m.set_mode(abi.MODE_BACKTEST)
    assert ei.value.code == abi.RLM_ERR_UNSUPPORTED
    m.close()
