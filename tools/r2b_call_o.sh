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
B="python bench.py --workload C4 --no-e2e --no-cpu-baseline --no-extras --steps 3 --warmup 3 --pretrain-ticks 3000"
run c4_88
run c4_0 RLM_ENVT_CARVEOUT=0
run c4_30 RLM_ENVT_CARVEOUT=30
B="python bench.py --workload C2 --no-e2e --no-cpu-baseline --no-extras --steps 3 --warmup 3 --pretrain-ticks 4000"
run c2_88
run c2_0 RLM_ENVT_CARVEOUT=0
run c2_30 RLM_ENVT_CARVEOUT=30
