"""rl_markets_b200/backtest.py renders the evaluation logs from step records.  Pinned here against the files the
UNMODIFIED reference wrote for the same run (spdlog "%v" profit_log of Backtester / Intraday::LogProfit, and
Base::writeStats' test_stats.csv; tests/golden/bt_*_{profit_log,test_stats}.csv from tools/make_golden.py), using the
reference's own step records -- no GPU involved."""
import os

import golden_util as G
from rl_markets_b200 import backtest


def test_profit_log_and_test_stats_match_the_reference_files():
    for case in G.backtest_manifest():
        recs, _keep = G.records(case["name"] + "_test")
        want = open(os.path.join(G.GOLD, case["name"] + "_profit_log.csv")).read().splitlines()
        got = [backtest.HEADER] + list(backtest.profit_rows(recs, 20100104))
        assert got == want, case["name"]
        s = case["summary"]
        rows = backtest.test_stats_rows(recs, s["test_market_sells"], s["test_market_buys"])
        assert "".join("%s,%d\n" % kv for kv in rows) == open(os.path.join(G.GOLD, case["name"] + "_test_stats.csv")).read()


def test_number_formatting_is_fmt_shortest_round_trip():
    assert [backtest._num(x) for x in (2750.0, 2750.5, -0.25, 1e-05, 5, -10, 0.1 + 0.2)] == \
        ["2750", "2750.5", "-0.25", "1e-05", "5", "-10", "0.30000000000000004"]
