"""CUDA path vs the CPU oracle, through the C ABI (rlm_create / rlm_run_ticks / rlm_read_records).

Bar (BASELINE.json north_star): bit-exact integer book state, <= 1e-5 relative on TD deltas.
This implementation keeps the reference's fp64 operation order, so every field of the step
record -- fp64 included -- is compared BITWISE, and so is theta.
"""
import ctypes as C

import pytest

from rl_markets_b200 import abi, config

pytestmark = pytest.mark.gpu


def _mk(algo, M, n_envs, flow_seed, rec_cap, source=abi.SOURCE_GENERATOR, **over):
    y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": algo, **over})
    cfg = config.from_dict(y, n_envs=n_envs, flow_seed=flow_seed, source=source)
    cfg.record_envs = n_envs
    cfg.record_cap = rec_cap
    return y, cfg


def _compare_env(rlm_records, port, label):
    n = min(len(rlm_records), port["steps"])
    assert n > 0, label + ": no steps to compare"
    for i in range(n):
        bad = abi.record_fields_equal(rlm_records[i], port["records"][i])
        if bad:
            det = []
            for f in bad:
                a, b = getattr(rlm_records[i], f), getattr(port["records"][i], f)
                if hasattr(a, "_fields_"):
                    det.append((f, [(k, getattr(a, k), getattr(b, k)) for k, _ in a._fields_]))
                elif hasattr(a, "__len__"):
                    det.append((f, list(a), list(b)))
                else:
                    det.append((f, a, b))
            raise AssertionError("%s: step %d differs (cuda, oracle): %r" % (label, i, det))
    return n


@pytest.mark.parametrize("algo,M", [("q_learn", 65536), ("sarsa", 16384), ("double_q_learn", 65536), ("q_learn", 4096)])
def test_generator_mode_matches_oracle(rlm, oracle, algo, M):
    n_envs, n_ticks = 8, 3000
    y, cfg = _mk(algo, M, n_envs, flow_seed=11, rec_cap=1500)
    m = rlm.BatchedMarket(cfg)
    m.run_ticks(1000)
    m.run_ticks(2000)  # chunked launches must not change anything
    m.sync()
    cnt = m.counters()
    total = 0
    for b in range(n_envs):
        ticks = oracle.generate_ticks(cfg, b, n_ticks)
        port = oracle.run_port(cfg, b, ticks)
        recs, keep = m.records(b)
        assert len(recs) == port["steps"], "env %d: step count cuda %d oracle %d" % (b, len(recs), port["steps"])
        total += _compare_env(recs, port, "%s env %d" % (algo, b))
        th = m.theta(b, 0)
        assert bytes(th) == bytes((C.c_double * M)(*port["theta"])), "%s env %d: theta differs" % (algo, b)
    assert cnt.steps == total
    assert cnt.ticks == sum([n_ticks - 1] * n_envs)  # the first row only opens the market (intraday.cpp:111-116)
    m.close()


def test_stream_mode_equals_generator_mode(rlm, oracle):
    n_envs, n_ticks = 4, 1500
    y, cfg = _mk("q_learn", 8192, n_envs, flow_seed=5, rec_cap=800, source=abi.SOURCE_STREAM)
    msgs = (abi.TickMsg * (n_ticks * n_envs))()
    for b in range(n_envs):
        one = rlm.flow_generate(cfg.flow, b, 0, n_ticks)
        for t in range(n_ticks):
            msgs[t * n_envs + b] = one[t]
    ms = rlm.BatchedMarket(cfg)
    ms.load_ticks(msgs, n_ticks)
    ms.run_ticks(700)
    ms.run_ticks(800)
    ms.sync()
    y2, cfg2 = _mk("q_learn", 8192, n_envs, flow_seed=5, rec_cap=800, source=abi.SOURCE_GENERATOR)
    mg = rlm.BatchedMarket(cfg2)
    mg.run_ticks(n_ticks)
    mg.sync()
    for b in range(n_envs):
        rs, _k1 = ms.records(b)
        rg, _k2 = mg.records(b)
        assert len(rs) == len(rg) and len(rs) > 50
        for i in range(len(rs)):
            assert not abi.record_fields_equal(rs[i], rg[i]), "env %d step %d" % (b, i)
        assert bytes(ms.theta(b)) == bytes(mg.theta(b))
    with pytest.raises(rlm.RlmError) as ei:
        ms.run_ticks(1)  # stream exhausted: performAction would return false (base.cpp:289)
    assert ei.value.code == abi.RLM_ERR_END_OF_DATA
    ms.close()
    mg.close()
