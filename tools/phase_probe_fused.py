#!/usr/bin/env python3
"""Where a warp of the fused persistent engine spends its cycles (RLM_TIMING build, RLM_ENGINE=F).
    RLM_ENGINE=F RLM_LIB_PATH=rl_markets_b200/librlm_timing.so python tools/phase_probe_fused.py [pretrain_ticks] [envs] [M] [algo]"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.')
from rl_markets_b200 import config, lib
pre = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
M = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
algo = sys.argv[4] if len(sys.argv) > 4 else "q_learn"
y = config.example_dict(**{"learning.memory_size": M, "learning.algorithm": algo})
cfg = config.from_dict(y, n_envs=B, flow_seed=1, dt_ms=1)
m = lib.BatchedMarket(cfg)
left = pre
while left > 0:
    m.run_ticks(min(left, 500)); left -= 500
m.sync()
L = m.L
L.rlm_debug_read_phases.argtypes = [C.c_void_p, C.c_void_p]
import time
t0 = time.perf_counter(); m.run_ticks(512); m.sync(); dt = time.perf_counter() - t0
clk = (C.c_longlong * (4096 * 16))(); sm = (C.c_uint * 4096)()
assert L.rlm_debug_read_phases(clk, sm) == 0
a = np.frombuffer(clk, dtype=np.int64).reshape(4096, 16)[:min(B, 4096), :6].astype(np.float64)
print("512 ticks: %.2f ms wall = %.1f us per tick" % (dt * 1e3, dt * 1e6 / 512))
print("per env: ticks %.0f  steps %.1f" % (a[:, 1].mean(), a[:, 4].mean()))
print("cycles per tick (envw_tick)        mean %8.0f  p10 %8.0f p90 %8.0f" % ((a[:, 0] / a[:, 1]).mean(), np.percentile(a[:, 0] / a[:, 1], 10), np.percentile(a[:, 0] / a[:, 1], 90)))
print("cycles per step waiting for a slot mean %8.0f  p90 %8.0f" % ((a[:, 2] / a[:, 4]).mean(), np.percentile(a[:, 2] / a[:, 4], 90)))
print("cycles per learner step            mean %8.0f  p10 %8.0f p90 %8.0f" % ((a[:, 3] / a[:, 4]).mean(), np.percentile(a[:, 3] / a[:, 4], 10), np.percentile(a[:, 3] / a[:, 4], 90)))
print("cycles per begin_step              mean %8.0f" % (a[:, 5] / a[:, 4]).mean())
tot = a[:, 0] + a[:, 2] + a[:, 3] + a[:, 5]
print("share of a warp's time: ticks %.2f  slot wait %.2f  learner %.2f  begin_step %.2f ; total cycles per env %.0f (= %.2f ms at 1.965 GHz)" % (
    a[:, 0].sum() / tot.sum(), a[:, 2].sum() / tot.sum(), a[:, 3].sum() / tot.sum(), a[:, 5].sum() / tot.sum(), tot.mean(), tot.mean() / 1.965e6))
m.close()
