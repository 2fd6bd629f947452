"""The reference's class surface over the C ABI (include/rlm_facade.hpp) and the loop of src/experiment/serial.cpp written
against it (examples/serial_driver.cpp, a compiled C++ host): two training episodes must leave exactly the weights the
fused rlm_run_ticks path leaves."""
import ctypes as C
import json
import os
import subprocess
import tempfile

import pytest

from rl_markets_b200 import abi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "examples", "serial_driver")


@pytest.mark.parametrize("algo", ["q_learn", "sarsa"])
def test_serial_driver_through_the_facade_matches_the_fused_path(rlm, algo):
    assert os.path.exists(DRIVER), "examples/serial_driver is built by __graft_entry__.build()"
    M, open_ticks, episodes = 8192, 400, 2
    with tempfile.TemporaryDirectory() as d:
        thp = os.path.join(d, "theta.bin")
        out = subprocess.check_output([DRIVER, "--episodes", str(episodes), "--algo", algo, "--memory-size", str(M), "--open-ticks",
                                       str(open_ticks), "--theta", thp]).decode()
        raw = open(thp, "rb").read()
    eps = [json.loads(l) for l in out.strip().splitlines()]
    assert len(eps) == episodes and all(e["steps"] > 40 for e in eps)
    L = rlm.load()
    cfg = abi.Config()
    rlm.check(L.rlm_config_default(C.byref(cfg)))
    cfg.algorithm = abi.ALGO[algo]
    cfg.memory_size = M
    cfg.flow.seed = 41
    cfg.flow.t0_ms = int(cfg.close_ms) - 30 * 60000 - open_ticks * cfg.flow.dt_ms
    m = rlm.BatchedMarket(cfg)
    for ep in range(episodes):
        m.run_ticks(open_ticks + 200)
        m.sync()
        st = m.stats(0, 1)[0]
        assert st.terminal == 1
        assert st.steps == eps[ep]["steps"] and st.episode_pnl == eps[ep]["pnl"] and st.episode_reward == eps[ep]["reward"]
        m.handle_terminal(ep)
        if ep + 1 < episodes:
            m.reset()
    assert bytes(m.theta(0, 0)) == raw
    m.close()
