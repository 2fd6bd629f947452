#!/usr/bin/env python3
"""Per-warp phase timeline of rlm_learn_kernel (the RLM_TIMING build: see rl_markets_b200/csrc/Makefile).

    RLM_LIB_PATH=rl_markets_b200/librlm_timing.so python tools/phase_probe_learn.py [pretrain_ticks] [envs] [M] [algo]
"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.')
from rl_markets_b200 import abi, config, lib
pre = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
M = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
algo = sys.argv[4] if len(sys.argv) > 4 else "q_learn"
y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": algo})
cfg = config.from_dict(y, n_envs=B, flow_seed=1, dt_ms=1)
m = lib.BatchedMarket(cfg)
left = pre
while left > 0:
    m.run_ticks(min(left, 250)); left -= 250
m.sync()
L = m.L
L.rlm_debug_read_phases.argtypes = [C.c_void_p, C.c_void_p]
names = ["stage AgentD", "hash", "gather issue", "tile table (under the gathers)", "gather wait + store", "sums", "TD",
         "trace pass", "threadfence", "patch", "sums 2", "write-back"]
for rep in range(3):
    m.run_ticks(1); m.sync()
    clk = (C.c_longlong * (4096 * 16))(); sm = (C.c_uint * 4096)()
    assert L.rlm_debug_read_phases(clk, sm) == 0
    a = np.frombuffer(clk, dtype=np.int64).reshape(4096, 16).copy()
    a = a[a[:, 12] > a[:, 0]]
    tl = a[:, 12].max()
    a = a[a[:, 0] > tl - 2000000]
    d = np.diff(a[:, :13], axis=1)
    print("rep %d: warps %d  launch span %d cycles  mean warp duration %.0f  max %d   start skew p50 %d p99 %d" % (
        rep, len(a), a[:, 12].max() - a[:, 0].min(), (a[:, 12] - a[:, 0]).mean(), (a[:, 12] - a[:, 0]).max(),
        np.percentile(a[:, 0] - a[:, 0].min(), 50), np.percentile(a[:, 0] - a[:, 0].min(), 99)))
    for i, n in enumerate(names):
        print("  %-32s mean %8.0f  p50 %8.0f  p90 %8.0f  max %8.0f" % (n, d[:, i].mean(), np.percentile(d[:, i], 50), np.percentile(d[:, i], 90), d[:, i].max()))
    print('  patch split: advance state %.0f | local patch %.0f | syncwarp %.0f' % ((a[:, 13] - a[:, 9]).mean(), (a[:, 14] - a[:, 13]).mean(), (a[:, 10] - a[:, 14]).mean()))
m.close()
