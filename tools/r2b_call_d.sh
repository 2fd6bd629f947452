#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
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
run ticksync RLM_ROUNDS=0
run cap2 RLM_ROUND_CAP=2
run cap4 RLM_ROUND_CAP=4
run cap6 RLM_ROUND_CAP=6
run cap3_s2 RLM_ROUNDS=1 RLM_ROUND_STREAMS=2
run cap4_s2 RLM_ROUNDS=1 RLM_ROUND_CAP=4 RLM_ROUND_STREAMS=2
run ticksync_nograph RLM_ROUNDS=0 RLM_GRAPHS=0
( time python bench.py --workload C4 --no-e2e --no-cpu-baseline --no-extras --steps 3 --warmup 3 ) > gpurun_out/c4.json 2> gpurun_out/c4.err; python -c "
import json; d=json.loads(open('gpurun_out/c4.json').read().strip().splitlines()[-1]); print('C4 value %.4e' % d['value'], d['roofline'].get('kernel'), d['roofline'].get('avg_launch_ms'), d['roofline'].get('other_kernel',{}).get('avg_launch_ms'))" || tail -5 gpurun_out/c4.err
