#!/usr/bin/env python3
"""Generate tests/golden/* from the UNMODIFIED reference compiled into oracle/_ref (run in the
build container, where /root/reference exists; the fixtures are committed so that the GPU box,
which has no reference tree, can check against them).

  tests/golden/units.json          reference unit-level vectors (oracle/_ref/ref_units)
  tests/golden/steps_<name>.bin    first N rlm_step_record of a reference run (oracle/_ref/ref_driver)
  tests/golden/bt_*_{profit_log,test_stats}.csv   the reference's own evaluation logs for the backtest cases
  tests/golden/manifest.json       the configs that produced them
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from rl_markets_b200 import abi, config  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_RECORDS = 400
CASES = [
    dict(name="q_learn_m65536", algo="q_learn", M=65536, flow_seed=7, env=0, ticks=2500, over={}),
    dict(name="sarsa_m16384", algo="sarsa", M=16384, flow_seed=7, env=1, ticks=2500, over={}),
    dict(name="double_q_m65536", algo="double_q_learn", M=65536, flow_seed=9, env=2, ticks=2500, over={}),
    dict(name="q_learn_m5003_greedy", algo="q_learn", M=5003, flow_seed=11, env=3, ticks=2500,
         over={"policy.eps_init": 0.05}),
    dict(name="q_learn_m4096_random_init", algo="q_learn", M=4096, flow_seed=13, env=5, ticks=2500,
         over={"learning.random_init": True, "debug.random_seed": 77}),
    # R-learning agents (agent.cpp:357-467) and the Boltzmann policy (policy.cpp:85-122)
    dict(name="r_learn_m16384", algo="r_learn", M=16384, flow_seed=15, env=6, ticks=2500, over={"policy.eps_init": 0.3}),
    dict(name="online_r_learn_m16384", algo="online_r_learn", M=16384, flow_seed=15, env=7, ticks=2500,
         over={"policy.eps_init": 0.3}),
    dict(name="double_r_learn_m8209", algo="double_r_learn", M=8209, flow_seed=17, env=8, ticks=2500,
         over={"policy.eps_init": 0.3, "learning.alpha_start": 0.01}),
    dict(name="sarsa_boltzmann_m8192", algo="sarsa", M=8192, flow_seed=19, env=9, ticks=2500,
         over={"policy.type": "boltzmann", "policy.tau_init": 0.05, "policy.tau_floor": 0.01, "policy.tau_T": 10}),
]


# train on one (short) synthetic day until the close, then main.cpp's evaluation phase (GoGreedy, a NEW Intraday,
# Backtester::RunEpisode) on another one: steps_<name>.bin = training records, steps_<name>_test.bin = evaluation
BACKTEST_CASES = [
    dict(name="bt_q_learn_m8192", algo="q_learn", M=8192, flow_seed=21, env=4, ticks=1200, train_open_ticks=900,
         test=dict(flow_seed=22, env=4, ticks=1000, open_ticks=700), over={"policy.eps_T": 3}),
    dict(name="bt_double_q_m8192", algo="double_q_learn", M=8192, flow_seed=23, env=5, ticks=1200, train_open_ticks=900,
         test=dict(flow_seed=24, env=5, ticks=1000, open_ticks=700), over={"learning.alpha_start": 0.01}),
]


def day_t0(cfg, open_ticks, dt_ms=250):
    """t0 such that the market closes (venue close - 30 min, market.cpp:67-70) `open_ticks` rows after the first one."""
    return int(cfg.close_ms) - 30 * 60000 - open_ticks * dt_ms


def main():
    os.makedirs(GOLD, exist_ok=True)
    units = subprocess.check_output([ol.REF_UNITS]).decode()
    json.loads(units)
    with open(os.path.join(GOLD, "units.json"), "w") as f:
        f.write(units)
    manifest = []
    for c in CASES:
        y = config.example_dict(**{"learning.memory_size": c["M"], "learning.algorithm": c["algo"], **c["over"]})
        # the reference seeds every generator with debug.random_seed (main.cpp:84-88); env b of a batch uses
        # random_seed + b, so the single-env reference run for env b gets that seed
        seed = y["debug"]["random_seed"] + c["env"]
        y_run = json.loads(json.dumps(y))
        y_run["debug"]["random_seed"] = seed
        ref = ol.run_ref(y_run, c["flow_seed"], c["env"], c["ticks"])
        recs = ref["records"][:N_RECORDS]
        with open(os.path.join(GOLD, "steps_%s.bin" % c["name"]), "wb") as f:
            for r in recs:
                f.write(bytes(r))
        manifest.append(dict(c, yaml=y, n_records=len(recs), summary=ref["summary"]))
        print(c["name"], len(recs), "records;", ref["summary"]["steps"], "reference steps")
    for c in BACKTEST_CASES:
        y = config.example_dict(**{"learning.memory_size": c["M"], "learning.algorithm": c["algo"], **c["over"]})
        seed = y["debug"]["random_seed"] + c["env"]
        y_run = json.loads(json.dumps(y))
        y_run["debug"]["random_seed"] = seed
        cfg = config.from_dict(y)
        t0 = day_t0(cfg, c["train_open_ticks"])
        test = dict(c["test"], t0_ms=day_t0(cfg, c["test"]["open_ticks"]))
        ref = ol.run_ref(y_run, c["flow_seed"], c["env"], c["ticks"], t0_ms=t0, test=dict(test, logs=True))
        # the evaluation logs as the reference itself writes them (Backtester's profit_log, Base::writeStats)
        for fname in ("profit_log.csv", "test_stats.csv"):
            with open(os.path.join(GOLD, "%s_%s" % (c["name"], fname)), "w") as f:
                f.write(ref["logs"][fname])
        assert ref["summary"]["terminal"] == 1
        for suffix, recs in (("", ref["records"]), ("_test", ref["test_records"])):
            with open(os.path.join(GOLD, "steps_%s%s.bin" % (c["name"], suffix)), "wb") as f:
                for r in recs:
                    f.write(bytes(r))
        manifest.append(dict(c, yaml=y, backtest=True, t0_ms=t0, test=test, n_records=len(ref["records"]),
                             n_test_records=len(ref["test_records"]), summary=ref["summary"]))
        print(c["name"], len(ref["records"]), "training records;", len(ref["test_records"]), "evaluation records")
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
