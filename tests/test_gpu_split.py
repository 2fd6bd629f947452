"""The split surface -- rlm_act / rlm_env_step / rlm_agent_update, the reference's Environment::step / Agent::update
seam (include/environment/base.h:132, include/rl/agent.h:60-67) -- driven in Learner::_step's order must reproduce the
fused rlm_run_ticks, and therefore the oracle, bit for bit."""
import ctypes as C

import pytest

from rl_markets_b200 import abi, config

pytestmark = pytest.mark.gpu


def _cfg(algo, n_envs, M=8192, seed=31, **over):
    y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": algo, **over})
    cfg = config.from_dict(y, n_envs=n_envs, flow_seed=seed)
    cfg.record_envs, cfg.record_cap = n_envs, 400
    return cfg


@pytest.mark.parametrize("algo,envv", [("q_learn", {}), ("sarsa", {}), ("double_q_learn", {}), ("q_learn", {"RLM_ENV_VARIANT": "1"})])
def test_split_surface_matches_oracle(rlm, oracle, monkeypatch, algo, envv):
    for k, v in envv.items():
        monkeypatch.setenv(k, v)
    n_envs, n_steps = 6, 150
    cfg = _cfg(algo, n_envs)
    m = rlm.BatchedMarket(cfg)
    rew, term = m.env_step(None)       # Runner::RunEpisode: environment.Initialise() (serial.cpp:18-25)
    assert not any(term)
    m.agent_update()                   # Q(first from-state, .)
    last_r, last_d = None, None
    for k in range(n_steps):
        a = m.act()                    # int action = m->action(*last_state)
        assert all(0 <= x < cfg.n_actions for x in a)
        rew, term = m.env_step(a)      # environment.performAction(action); getReward()
        d = m.agent_update()           # state->newState(env); m->HandleTransition(...)
        last_r, last_d = list(rew), list(d)
    m.sync()
    c = m.counters()
    assert c.steps == n_envs * n_steps
    st = m.stats()
    for b in range(n_envs):
        recs, _k = m.records(b)
        assert len(recs) == n_steps
        # the oracle stops after the same number of learner steps (each env consumed its own number of ticks)
        port = oracle.run_port(cfg, b, oracle.generate_ticks(cfg, b, 6000), max_steps=n_steps)
        assert port["steps"] == n_steps
        for i in range(n_steps):
            bad = abi.record_fields_equal(recs[i], port["records"][i])
            assert not bad, (algo, b, i, bad)
        assert recs[-1].reward == last_r[b] and recs[-1].delta == last_d[b]
        assert bytes(m.theta(b, 0)) == bytes((C.c_double * cfg.memory_size)(*port["theta"]))
    m.close()


def test_external_policy_and_fused_continuation(rlm, oracle):
    """Actions supplied from outside (no Agent::action, no generator draw); then the handle continues under rlm_run_ticks."""
    n_envs = 4
    cfg = _cfg("q_learn", n_envs, seed=7)
    m = rlm.BatchedMarket(cfg)
    m.env_step(None)
    m.agent_update()
    script = [(3 * k + b) % cfg.n_actions for k in range(40) for b in range(n_envs)]
    for k in range(40):
        acts = (C.c_int32 * n_envs)(*script[k * n_envs:(k + 1) * n_envs])
        m.env_step(acts)
        m.agent_update()
    for b in range(n_envs):
        recs, _k = m.records(b)
        assert [r.action for r in recs] == [script[k * n_envs + b] for k in range(40)]
    m.run_ticks(300)   # the fused path picks every env up where the split surface left it
    m.sync()
    assert m.counters().steps > n_envs * 40
    for s in m.stats():
        assert s.phase == 2
    m.close()


def test_split_surface_rejects_what_it_cannot_do(rlm):
    y = config.example_dict(**{"learning.memory_size": 4096})
    cfg = config.from_dict(y, n_envs=2, source=abi.SOURCE_STREAM)
    m = rlm.BatchedMarket(cfg)
    with pytest.raises(rlm.RlmError) as ei:
        m.act()
    assert ei.value.code == abi.RLM_ERR_UNSUPPORTED
    m.close()
