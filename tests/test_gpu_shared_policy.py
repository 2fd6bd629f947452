"""Shared policy on the GPU vs the synchronous-batch CPU oracle.
fp64 atomics make the order of the dtheta sum nondeterministic: integer state must match exactly,
TD deltas and weights within 1e-5 relative (BASELINE.json north_star); in practice ~1e-12."""
import ctypes as C

import numpy as np
import pytest

from rl_markets_b200 import abi, config

pytestmark = pytest.mark.gpu
INT_FIELDS = ["step", "action", "time_ms", "terminal", "position", "ask_level", "bid_level", "ask_transactions",
              "bid_transactions", "market_buys", "market_sells", "lo_vol_step", "n_traces"]


@pytest.mark.parametrize("algo", ["q_learn", "double_q_learn"])
def test_shared_policy_matches_batch_oracle(rlm, oracle, algo):
    from test_shared_policy_cpu import _oracle_batch
    n_envs, n_ticks, M, cap = 16, 400, 4096, 300
    y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": algo})
    cfg = config.from_dict(y, n_envs=n_envs, shared_policy=True, flow_seed=33)
    cfg.record_envs, cfg.record_cap = n_envs, cap
    m = rlm.BatchedMarket(cfg)
    m.run_ticks(n_ticks)
    m.sync()
    L, b = _oracle_batch(cfg)
    streams = [rlm.flow_generate(cfg.flow, i, 0, n_ticks) for i in range(n_envs)]
    recs = (abi.StepRecord * (n_envs * cap))()
    cnt = (C.c_int32 * n_envs)()
    msgs = (abi.TickMsg * n_envs)()
    for t in range(n_ticks):
        for i in range(n_envs):
            msgs[i] = streams[i][t]
        L.lobo_batch_accumulate(b, msgs, C.addressof(recs), C.addressof(cnt), cap)
        L.lobo_batch_apply(b)
    total = 0
    for e in range(n_envs):
        got, _keep = m.records(e)
        assert len(got) == cnt[e] > 20, (e, len(got), cnt[e])
        for i, r in enumerate(got):
            o = recs[e * cap + i]
            for f in INT_FIELDS:
                assert getattr(r, f) == getattr(o, f), (algo, e, i, f)
            assert bytes(r.ask) == bytes(o.ask) and bytes(r.bid) == bytes(o.bid) and bytes(r.state) == bytes(o.state)
            assert r.ep_pnl == o.ep_pnl and r.ask_quote == o.ask_quote and r.bid_quote == o.bid_quote
            for f in ("delta", "reward", "pnl_step", "ep_reward"):
                a, bb = getattr(r, f), getattr(o, f)
                assert abs(a - bb) <= 1e-5 * max(abs(a), abs(bb), 1e-12), (algo, e, i, f, a, bb)
        total += len(got)
    assert m.counters().steps == total == L.lobo_batch_steps(b)
    th = np.frombuffer(m.theta(0, 0), dtype=np.float64)
    ref = np.ctypeslib.as_array(L.lobo_batch_theta(b, 0), shape=(M,))
    np.testing.assert_allclose(th, ref, rtol=1e-5, atol=1e-12)
    assert np.count_nonzero(ref) > 100
    L.lobo_batch_destroy(b)
    m.close()


def test_two_replicas_of_a_shared_policy_match_the_batch_oracle(rlm, oracle):
    """What two ranks do, in one process: two handles own the two halves of the batch and their own replica of theta;
    each tick both accumulate, the two dtheta buffers are summed into both (the all-reduce), both apply.  Weights written
    only by the OTHER replica's envs must be visible to this replica's evaluations (round 1 filtered its gathers through
    a rank-local occupancy bitmap)."""
    import torch
    from test_shared_policy_cpu import _oracle_batch
    n_envs, n_ticks, M, cap = 16, 300, 4096, 250
    y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": "q_learn"})
    halves = []
    for r in range(2):
        c = config.from_dict(y, n_envs=n_envs // 2, env_index0=r * (n_envs // 2), shared_policy=True, flow_seed=33)
        c.record_envs, c.record_cap = n_envs // 2, cap
        halves.append(rlm.BatchedMarket(c))
    stream = torch.cuda.Stream()  # (a real stream: 0 would mean "the handle's own stream" to rlm_set_stream)
    dths = []
    for m in halves:
        m.set_stream(stream.cuda_stream)
        dths.append(m.dtheta_tensor())
    with torch.cuda.stream(stream):
        for t in range(n_ticks):
            for m in halves:
                m.shared_tick_accumulate()
            total = dths[0] + dths[1]
            dths[0].copy_(total)
            dths[1].copy_(total)
            for m in halves:
                m.apply_dtheta()
    for m in halves:
        m.sync()
    cfg = config.from_dict(y, n_envs=n_envs, shared_policy=True, flow_seed=33)
    L, b = _oracle_batch(cfg)
    streams = [rlm.flow_generate(cfg.flow, i, 0, n_ticks) for i in range(n_envs)]
    recs = (abi.StepRecord * (n_envs * cap))()
    cnt = (C.c_int32 * n_envs)()
    msgs = (abi.TickMsg * n_envs)()
    for t in range(n_ticks):
        for i in range(n_envs):
            msgs[i] = streams[i][t]
        L.lobo_batch_accumulate(b, msgs, C.addressof(recs), C.addressof(cnt), cap)
        L.lobo_batch_apply(b)
    for e in range(n_envs):
        got, _keep = halves[e // (n_envs // 2)].records(e % (n_envs // 2))
        assert len(got) == cnt[e] > 20, (e, len(got), cnt[e])
        for i, r in enumerate(got):
            o = recs[e * cap + i]
            for f in INT_FIELDS:
                assert getattr(r, f) == getattr(o, f), (e, i, f)
            assert abs(r.delta - o.delta) <= 1e-5 * max(abs(r.delta), abs(o.delta), 1e-12), (e, i, r.delta, o.delta)
    ref = np.ctypeslib.as_array(L.lobo_batch_theta(b, 0), shape=(M,))
    for m in halves:
        np.testing.assert_allclose(np.frombuffer(m.theta(0, 0), dtype=np.float64), ref, rtol=1e-5, atol=1e-12)
    assert bytes(halves[0].theta(0, 0)) == bytes(halves[1].theta(0, 0))  # replicas stay bitwise identical
    L.lobo_batch_destroy(b)
    for m in halves:
        m.close()
