#!/bin/bash
# Round 2, session b, GPU call A: parity of the new switches, A/B timings at C1, by-line ncu captures of both C1 kernels.
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 5 --warmup 3 --pretrain-ticks 20000"
run() { # name, env assignments...
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
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
RLM_LIB_PATH=$PWD/rl_markets_b200/librlm_b.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run base
run envw4 RLM_ENVW_WARPS=4
run envw2 RLM_ENVW_WARPS=2
run rounds_cap1 RLM_ROUNDS=1 RLM_ROUND_CAP=1
run rounds_cap2 RLM_ROUNDS=1 RLM_ROUND_CAP=2
run rounds_cap2_s2 RLM_ROUNDS=1 RLM_ROUND_CAP=2 RLM_ROUND_STREAMS=2
LB=$PWD/rl_markets_b200/librlm_b.so
run b_base RLM_LIB_PATH=$LB
run b_rounds_cap2 RLM_LIB_PATH=$LB RLM_ROUNDS=1 RLM_ROUND_CAP=2
run b_rounds_cap2_s2 RLM_LIB_PATH=$LB RLM_ROUNDS=1 RLM_ROUND_CAP=2 RLM_ROUND_STREAMS=2
run b_rounds_cap4_s2 RLM_LIB_PATH=$LB RLM_ROUNDS=1 RLM_ROUND_CAP=4 RLM_ROUND_STREAMS=2
run base2
# by-line captures (direct launches so that every kernel is its own ncu launch)
RLM_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rlm_learn_kernel|rlm_env_kernel_w" --launch-skip 41000 -c 2 -f -o gpurun_out/r2b_c1_full \
  python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 4 --warmup 3 --ticks 64 --pretrain-ticks 20000 > gpurun_out/ncu_c1_full.log 2>&1
ls -la gpurun_out/ | head -40
