import ctypes as C, sys, numpy as np
sys.path.insert(0, '.')
from rl_markets_b200 import abi, config, lib
y = config.example_dict(**{"learning.memory_size": 65536, "learning.algorithm": "q_learn"})
cfg = config.from_dict(y, n_envs=4096, flow_seed=1)
m = lib.BatchedMarket(cfg)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10): m.run_ticks(64)
m.sync()
m.run_ticks(1); m.sync()   # last launch's phases are what is read
L = m.L
clk = (C.c_longlong * (4096 * 16))(); sm = (C.c_uint * 4096)()
L.rlm_debug_read_phases.argtypes = [C.c_void_p, C.c_void_p]
assert L.rlm_debug_read_phases(clk, sm) == 0
a = np.frombuffer(clk, dtype=np.int64).reshape(4096, 16)
# rows of this launch: those whose phase 12 > phase 0 and belong to the final launch (largest clocks)
ok = a[:, 12] > a[:, 0]
a = a[ok]
tl = a[:, 12].max()
a = a[(a[:, 0] > tl - 400000) & (a[:, 2] > a[:, 0])]   # the last launch, learner-step CTAs
t0 = a[:, 0].min()
d = np.diff(a[:, :13], axis=1)
names = ["stage AgentD", "hash(to)", "occ+gather issue/wait", "barrier", "sums", "td_decision", "trace_pass", "threadfence", "record/bookkeep+barrier", "re-gather", "sums2+barrier", "writeback"]
print("CTAs", len(a), "launch span (cycles)", a[:, 12].max() - a[:, 0].min(), "mean CTA duration", (a[:, 12] - a[:, 0]).mean(), "max", (a[:, 12] - a[:, 0]).max())
print("start skew: p50 %d p99 %d" % (np.percentile(a[:, 0] - t0, 50), np.percentile(a[:, 0] - t0, 99)))
for i, n in enumerate(names):
    print("%-28s mean %8.0f  p90 %8.0f  max %8.0f" % (n, d[:, i].mean(), np.percentile(d[:, i], 90), d[:, i].max()))

order = np.argsort(-d[:, 6])[:12]
print("slowest trace passes: cycles, n_traces before, decayed(rate != 0)")
for i in order:
    print("  %7d  n=%4d  decay=%d   (td_decision %d, gather %d)" % (d[i, 6], a[i, 13], a[i, 14], d[i, 5], d[i, 2]))
import collections
for dec in (0, 1):
    sel = a[:, 14] == dec
    if sel.any():
        print("decay=%d: %d CTAs, trace_pass mean %.0f, n mean %.1f" % (dec, sel.sum(), d[sel, 6].mean(), a[sel, 13].mean()))
