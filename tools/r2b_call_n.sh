#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --steps 5 --warmup 3 --pretrain-ticks 20000 --e2e-steps 10"
run() {
  local name=$1; shift
  ( env "$@" $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err ) ; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ab_%s.json'%n).read().strip().splitlines()[-1])
    rf=d['roofline']
    print("AB %-22s value %.4e e2e %.4e ms/step %.2f dom %.1fus (%.0f steps) other %.1fus" % (n, d['value'], d.get('e2e',{}).get('value',0), d['ms_per_step'], 1e3*(rf.get('avg_launch_ms') or 0), rf.get('env_steps_per_launch',0), 1e3*((rf.get('other_kernel') or {}).get('avg_launch_ms') or 0)))
except Exception as e:
    print("AB %s FAILED %s" % (n, e)); print(open('gpurun_out/ab_%s.err'%n).read()[-800:])
PY
}
run default
run carve50 RLM_ENVW_CARVEOUT=50
run carve35 RLM_ENVW_CARVEOUT=35
run ticksync RLM_ROUNDS=0
run ticksync_carve50 RLM_ROUNDS=0 RLM_ENVW_CARVEOUT=50
