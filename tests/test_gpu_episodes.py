"""Episode end and restart on the GPU vs the oracle: the env reaches the close (Intraday::isTerminal,
intraday.cpp:152-157), Runner::RunEpisode clears the inventory with a market order (serial.cpp:31,
WalkTheBook book.cpp:429-539), Agent/Policy::HandleTerminal reschedule alpha and epsilon
(agent.cpp:103-109, policy.cpp:79-82), and the next episode starts from Intraday::Initialise."""
import ctypes as C

import pytest

from rl_markets_b200 import abi, config

pytestmark = pytest.mark.gpu


def test_episode_end_handle_terminal_and_second_episode(rlm, oracle):
    n_envs, cap = 6, 900
    y = config.example_dict(**{"learning.memory_size": 8192, "learning.omega": 0.9, "learning.alpha_start": 0.01,
                               "policy.eps_T": 3})
    cfg = config.from_dict(y, n_envs=n_envs, flow_seed=41, dt_ms=250)
    # market closes (16:30 - 30 min) 500 ticks after the first row
    cfg.flow.t0_ms = int(cfg.close_ms) - 30 * 60000 - 500 * 250
    cfg.record_envs, cfg.record_cap = n_envs, cap
    m = rlm.BatchedMarket(cfg)
    m.run_ticks(700)  # more than the episode holds: envs stop at the close
    m.sync()
    st = m.stats()
    L = oracle.lib()
    hs, ticks = [], []
    for b in range(n_envs):
        t = rlm.flow_generate(cfg.flow, b, 0, 1400)
        h = L.lobo_create(C.byref(cfg), b)
        recs = (abi.StepRecord * cap)()
        used = C.c_int64()
        n1 = L.lobo_run(h, t, 700, -1, recs, cap, C.byref(used))
        assert L.lobo_is_terminal(h) == 1 and used.value < 700
        so = abi.EnvStats()
        L.lobo_stats(h, C.byref(so))
        got, _k = m.records(b)
        assert len(got) == n1 > 50
        for i in range(n1):
            assert not abi.record_fields_equal(got[i], recs[i]), (b, i)
        assert st[b].terminal == 1 and bytes(st[b]) == bytes(so), (b, [(f, getattr(st[b], f), getattr(so, f)) for f, _ in abi.EnvStats._fields_])
        hs.append((h, recs, n1))
        ticks.append(t)
    # Learner::RunEpisode: HandleTerminal(episode), then the next episode
    m.handle_terminal(1)
    m.reset()
    m.run_ticks(300)
    m.sync()
    for b in range(n_envs):
        h, recs, n1 = hs[b]
        L.lobo_handle_terminal(h, 1)
        L.lobo_reset(h)
        recs2 = (abi.StepRecord * cap)()
        used = C.c_int64()
        n2 = L.lobo_run(h, ticks[b], 300, -1, recs2, cap, C.byref(used))
        got, _k = m.records(b)
        assert len(got) == n1 + n2 and n2 > 20, (b, len(got), n1, n2)
        for i in range(n2):
            bad = abi.record_fields_equal(got[n1 + i], recs2[i])
            assert not bad, (b, i, bad)
        assert bytes(m.theta(b)) == bytes((C.c_double * cfg.memory_size).from_address(C.addressof(L.lobo_theta(h, 0).contents)))
        L.lobo_destroy(h)
    m.close()
