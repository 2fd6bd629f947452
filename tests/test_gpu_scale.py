"""Parity IN THE REGIME THE BENCH TIMES: full batches, tens of thousands of ticks, weight tables that have filled up,
trace lists in steady state -- CUDA path vs the CPU oracle on sampled envs, bitwise (records, theta, statistics).

The oracle (pinned against the compiled reference, tests/test_oracle_golden.py) runs ~1e5 ticks/s per env, so a dozen
sampled envs over 20 000 ticks cost a few seconds."""
import ctypes as C

import pytest

from rl_markets_b200 import abi, config

pytestmark = pytest.mark.gpu


def _cfg(B, M, algo, seed, n_rec, rec_cap):
    y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": algo})
    cfg = config.from_dict(y, n_envs=B, flow_seed=seed, dt_ms=1)  # dt 1 ms: no env reaches the close
    cfg.record_envs, cfg.record_cap = n_rec, rec_cap
    return cfg


def _check(rlm, oracle, m, cfg, T, rec_envs, theta_envs, label):
    st = m.stats()
    for b in sorted(set(rec_envs) | set(theta_envs)):
        port = oracle.run_port(cfg, b, oracle.generate_ticks(cfg, b, T))
        if b in rec_envs:
            recs, _keep = m.records(b)
            assert len(recs) == port["steps"] > T // 8, (label, b, len(recs), port["steps"])
            for i in range(len(recs)):
                bad = abi.record_fields_equal(recs[i], port["records"][i])
                assert not bad, "%s env %d step %d of %d differs: %r" % (label, b, i, len(recs), bad)
        assert bytes(m.theta(b, 0)) == bytes((C.c_double * cfg.memory_size)(*port["theta"])), "%s env %d: theta differs" % (label, b)
        assert bytes(st[b]) == bytes(port["stats"]), (label, b)
        nz = sum(1 for x in port["theta"] if x != 0.0)
        assert nz > cfg.memory_size // 4, (label, b, nz)  # the tables have filled up


def test_c1_long_run_matches_oracle(rlm, oracle):
    """BASELINE.json configs[1]: 4096 LOBs, Q-learning, M = 2^16 per env; 20 000 ticks (~5 800 learner steps per env)."""
    B, T = 4096, 20000
    cfg = _cfg(B, 65536, "q_learn", 2024, 6, 7000)
    m = rlm.BatchedMarket(cfg)
    for _ in range(T // 250):
        m.run_ticks(250)
    m.sync()
    c = m.counters()
    assert c.ticks == B * (T - 1) and c.steps > B * T // 5
    _check(rlm, oracle, m, cfg, T, rec_envs=range(6), theta_envs=(777, 2048, 3333, 4095), label="C1")
    m.close()


def test_c2_sarsa_lambda_matches_oracle(rlm, oracle):
    """configs[2]: 65 536 LOBs, SARSA(lambda), M = 2^14 per env (thread-per-env tick kernel: B > 16384)."""
    B, T = 65536, 3000
    cfg = _cfg(B, 16384, "sarsa", 77, 4, 1200)
    m = rlm.BatchedMarket(cfg)
    for _ in range(T // 250):
        m.run_ticks(250)
    m.sync()
    _check(rlm, oracle, m, cfg, T, rec_envs=range(4), theta_envs=(12345, 40000, 65535), label="C2")
    m.close()


def test_thread_per_env_tick_kernel_at_32768(rlm, oracle):
    """The tick kernel variant large batches use (one thread per env) at B = 32768, Double-Q."""
    B, T = 32768, 2000
    cfg = _cfg(B, 8192, "double_q_learn", 5, 3, 900)
    m = rlm.BatchedMarket(cfg)
    m.run_ticks(T)
    m.sync()
    _check(rlm, oracle, m, cfg, T, rec_envs=range(3), theta_envs=(1000, 32767), label="B32768")
    m.close()
