"""How the per-env weight tables fill up, and what a bench step (64 ticks, C1) costs along the way."""
import sys, time
sys.path.insert(0, ".")
import torch
from rl_markets_b200 import config, lib
y = config.example_dict(**{"learning.memory_size": 65536, "learning.algorithm": "q_learn"})
cfg = config.from_dict(y, n_envs=4096, flow_seed=2024, dt_ms=1)
m = lib.BatchedMarket(cfg)
m.run_ticks(64); m.sync()
t_total = 0
for block in range(14):
    n = 50 if block < 6 else 200
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        m.run_ticks(64)
    m.sync(); dt = time.perf_counter() - t0
    t_total += n
    occ = m.occupancy()
    mean = sum(occ) / len(occ)
    print("after %5d bench steps: %.2f ms/step, mean occupancy %.1f%% (max %.1f%%), steps/env %.0f" % (
        t_total, dt / n * 1e3, 100 * mean / 65536, 100 * max(occ) / 65536, m.counters().steps / 4096))
