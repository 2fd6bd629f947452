"""Shared-policy plumbing on CPU: world_size-2 gloo all-reduce of dtheta between two shards must
reproduce the single-process batch (rl_markets_b200.parallel.shared_policy_tick).  The compute
backend here is the CPU oracle's batch object (tests only); on the GPU the same function drives
rlm_shared_tick_accumulate / rlm_apply_dtheta (tests/test_gpu_shared_policy.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl_markets_b200 import abi, config, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_ENVS, N_TICKS, M = 8, 260, 2048


def _cfg(n_envs, env_index0):
    y = config.example_dict(**{"learning.memory_size": M})
    c = config.from_dict(y, n_envs=n_envs, env_index0=env_index0, shared_policy=True, flow_seed=21)
    return c


def _oracle_batch(cfg):
    import oracle_lib as ol
    L = ol.lib()
    L.lobo_batch_create.restype = C.c_void_p
    L.lobo_batch_create.argtypes = [C.POINTER(abi.Config)]
    L.lobo_batch_destroy.argtypes = [C.c_void_p]
    L.lobo_batch_accumulate.argtypes = [C.c_void_p, C.POINTER(abi.TickMsg), C.c_void_p, C.c_void_p, C.c_int32]
    L.lobo_batch_dtheta.restype = C.POINTER(C.c_double)
    L.lobo_batch_dtheta.argtypes = [C.c_void_p, C.c_int]
    L.lobo_batch_theta.restype = C.POINTER(C.c_double)
    L.lobo_batch_theta.argtypes = [C.c_void_p, C.c_int]
    L.lobo_batch_apply.argtypes = [C.c_void_p]
    L.lobo_batch_steps.restype = C.c_int64
    L.lobo_batch_steps.argtypes = [C.c_void_p]
    return L, L.lobo_batch_create(C.byref(cfg))


def _run(rank, world, out_q, port):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = parallel.shard(N_ENVS, rank, world)
    cfg = _cfg(n, first)
    L, b = _oracle_batch(cfg)
    streams = [ol.lib_generate(cfg, first + i, N_TICKS) for i in range(n)]
    dth = torch.from_numpy(np.ctypeslib.as_array(L.lobo_batch_dtheta(b, 0), shape=(M,)))
    msgs = (abi.TickMsg * n)()
    for t in range(N_TICKS):
        for i in range(n):
            msgs[i] = streams[i][t]
        parallel.shared_policy_tick(lambda: L.lobo_batch_accumulate(b, msgs, None, None, 0), dth,
                                    lambda: L.lobo_batch_apply(b), dist if world > 1 else None)
    theta = np.ctypeslib.as_array(L.lobo_batch_theta(b, 0), shape=(M,)).copy()
    out_q.put((rank, theta, L.lobo_batch_steps(b)))
    L.lobo_batch_destroy(b)
    if world > 1:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    _run(0, 1, q, 0)
    _r, theta_1, steps_1 = q.get()
    assert steps_1 > N_ENVS * 10 and np.count_nonzero(theta_1) > 50
    procs = [ctx.Process(target=_run, args=(r, 2, q, 29541)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas agree bitwise after every all-reduce; the sharded run equals the single batch up to the
    # order of the fp64 sum (north_star: 1e-5 relative on weight deltas)
    assert np.array_equal(res[0][1], res[1][1])
    assert res[0][2] + res[1][2] == steps_1
    np.testing.assert_allclose(res[0][1], theta_1, rtol=1e-9, atol=1e-13)


def test_shard_partition():
    assert [parallel.shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert parallel.shard(262144, 7, 8) == (7 * 32768, 32768)
