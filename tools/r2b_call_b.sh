#!/bin/bash
# GPU call B: scan of the round-paced engine's cap / streams and of the learner build knobs at C1.
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 5 --warmup 3 --pretrain-ticks 20000"
run() {
  local name=$1; shift
  ( env "$@" $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err ) ; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ab_%s.json'%n).read().strip().splitlines()[-1])
    rf=d['roofline']
    print("AB %-28s value %.4e ms/step %.2f dom %.1fus other %.1fus" % (n, d['value'], d['ms_per_step'], 1e3*(rf.get('avg_launch_ms') or 0), 1e3*((rf.get('other_kernel') or {}).get('avg_launch_ms') or 0)))
except Exception as e:
    print("AB %s FAILED %s" % (n, e)); print(open('gpurun_out/ab_%s.err'%n).read()[-800:])
PY
}
L=$PWD/rl_markets_b200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_episodes.py -x -q -m gpu 2>&1 | tail -3
RLM_LIB_PATH=$L/librlm_g.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run base
for v in g h i; do run ${v}_base RLM_LIB_PATH=$L/librlm_$v.so; done
for cap in 2 3 4 6 0; do for s in 1 2; do run r_cap${cap}_s$s RLM_ROUNDS=1 RLM_ROUND_CAP=$cap RLM_ROUND_STREAMS=$s; done; done
run r_cap4_s4 RLM_ROUNDS=1 RLM_ROUND_CAP=4 RLM_ROUND_STREAMS=4
for v in g h i; do for cap in 2 4; do for s in 1 2; do run ${v}_cap${cap}_s$s RLM_LIB_PATH=$L/librlm_$v.so RLM_ROUNDS=1 RLM_ROUND_CAP=$cap RLM_ROUND_STREAMS=$s; done; done; done
run base2
