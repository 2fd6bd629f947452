"""Full-size (BASELINE.json configs[1]: 4096 LOBs) size-independent properties of the CUDA path."""
import ctypes as C

import pytest

from rl_markets_b200 import abi, config

pytestmark = pytest.mark.gpu


def _cfg(n_envs, M=65536, algo="q_learn", seed=2024, dt=1):
    y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": algo})
    return config.from_dict(y, n_envs=n_envs, flow_seed=seed, dt_ms=dt)


def test_full_size_run_is_deterministic_and_chunking_invariant(rlm):
    B, T = 4096, 192
    a = rlm.BatchedMarket(_cfg(B))
    a.run_ticks(T)
    a.sync()
    b = rlm.BatchedMarket(_cfg(B))
    for _ in range(3):
        b.run_ticks(T // 3)
    b.sync()
    ca, cb = a.counters(), b.counters()
    assert (ca.steps, ca.ticks, ca.sum_traces) == (cb.steps, cb.ticks, cb.sum_traces)
    assert ca.ticks == B * (T - 1) and ca.steps > B * 20
    sa, sb = a.stats(), b.stats()
    assert bytes(sa) == bytes(sb)
    for env in (0, 1, 777, 4095):
        assert bytes(a.theta(env)) == bytes(b.theta(env))
    assert list(a.actions()) == list(b.actions()) and bytes(a.rewards()) == bytes(b.rewards())
    # book-keeping identities: position bounded by the risk limits (+- one order), every env advanced
    for s in sa:
        assert -60 <= s.position <= 60 and s.phase == 2 and s.steps > 10
    a.close()
    b.close()


def test_env_results_do_not_depend_on_batch_or_shard(rlm):
    """Env g behaves identically as env g of a 64-env handle and as env 0 of a shard starting at g
    (this is what lets bench.py shard envs across GPUs with no data-path collective)."""
    T = 400
    big = rlm.BatchedMarket(_cfg(64, M=8192))
    big.run_ticks(T)
    big.sync()
    for g in (0, 17, 63):
        c = _cfg(2, M=8192)
        c.env_index0 = g
        one = rlm.BatchedMarket(c)
        one.run_ticks(T)
        one.sync()
        assert bytes(one.theta(0)) == bytes(big.theta(g))
        assert bytes(one.stats(0, 1)) == bytes(big.stats(g, 1))
        one.close()
    big.close()


def test_unsupported_configs_fail_loudly(rlm):
    c = _cfg(4)
    c.n_tilings = 16
    with pytest.raises(rlm.RlmError) as ei:
        rlm.BatchedMarket(c)
    assert ei.value.code == abi.RLM_ERR_UNSUPPORTED
    c = _cfg(4)
    c.policy_type = 7  # main.cpp:164-165 "Please specify a valid policy!"
    with pytest.raises(rlm.RlmError) as ei:
        rlm.BatchedMarket(c)
    assert ei.value.code == abi.RLM_ERR_INVALID_ARGUMENT
    c = _cfg(4)
    c.algorithm = abi.ALGO["r_learn"]  # rho is per agent: no shared-policy formulation
    c.shared_policy = 1
    with pytest.raises(rlm.RlmError) as ei:
        rlm.BatchedMarket(c)
    assert ei.value.code == abi.RLM_ERR_UNSUPPORTED
