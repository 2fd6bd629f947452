#!/usr/bin/env python3
"""Launch timeline of the tick-synchronous engine with sub-batch streams (RLM_TIMING build): when did each env tick
kernel and learner kernel of each sub-batch start and end (globaltimer, ns)?
    RLM_SUBBATCHES=2 RLM_LIB_PATH=rl_markets_b200/librlm_timing.so python tools/timeline_probe.py [pretrain] [envs] [M] [algo]"""
import ctypes as C, sys, os, numpy as np
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
    m.run_ticks(min(left, 250)); left -= 250
m.sync()
L = m.L
L.rlm_debug_klog.argtypes = [C.c_void_p, C.c_int]
assert L.rlm_debug_klog(None, 1) == 0
m.run_ticks(64); m.sync()
buf = (C.c_ulonglong * (256 * 8 * 2 * 2))()
assert L.rlm_debug_klog(buf, 0) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(256, 8, 2, 2).astype(np.int64)
S = int(os.environ.get("RLM_SUBBATCHES", "1"))
t0 = min(a[t, s, 0, 0] for t in range(64) for s in range(S) if a[t, s, 0, 1] > 0)
print("sub-batches %d; times in us since the first env kernel of the call" % S)
for t in list(range(20, 24)):
    row = []
    for s in range(S):
        e0, e1, l0, l1 = a[t, s, 0, 0], a[t, s, 0, 1], a[t, s, 1, 0], a[t, s, 1, 1]
        row.append("s%d env %7.1f-%7.1f (%5.1f) learn %7.1f-%7.1f (%5.1f)" % (s, (e0 - t0) / 1e3, (e1 - t0) / 1e3, (e1 - e0) / 1e3, (l0 - t0) / 1e3, (l1 - t0) / 1e3, (l1 - l0) / 1e3))
    print("tick %2d | " % t + " | ".join(row))
tend = max(a[63, s, 1, 1] for s in range(S))
print("64 ticks in %.1f us = %.1f us per tick" % ((tend - t0) / 1e3, (tend - t0) / 64e3))
for s in range(S):
    ed = [(a[t, s, 0, 1] - a[t, s, 0, 0]) / 1e3 for t in range(8, 64)]
    ld = [(a[t, s, 1, 1] - a[t, s, 1, 0]) / 1e3 for t in range(8, 64) if a[t, s, 1, 1] > 0]
    gap1 = [(a[t, s, 1, 0] - a[t, s, 0, 1]) / 1e3 for t in range(8, 64) if a[t, s, 1, 1] > 0]
    gap2 = [(a[t + 1, s, 0, 0] - a[t, s, 1, 1]) / 1e3 for t in range(8, 63) if a[t, s, 1, 1] > 0]
    print("sub %d: env kernel %.1f us, learner kernel %.1f us, gap env->learner %.1f us, gap learner->next env %.1f us" % (s, np.mean(ed), np.mean(ld), np.mean(gap1), np.mean(gap2)))
m.close()
