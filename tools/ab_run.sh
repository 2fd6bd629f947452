#!/bin/bash
# A/B runs on ONE gpurun box (profiles/r2b_ab_runs.txt was made with this): every argument is "name[:VAR=value[,VAR=value...]]";
# each runs bench.py at C1 (or $AB_WORKLOAD) with those environment variables -- RLM_LIB_PATH=<other build> compares library
# variants (make EXTRA=-D... BUILD=build_x OUT=.../librlm_x.so) -- and prints one line per run:
#   tools/ab_run.sh base ticksync:RLM_ROUNDS=0 cap4:RLM_ROUND_CAP=4 w2:RLM_LIB_PATH=$PWD/rl_markets_b200/librlm_w2.so
mkdir -p gpurun_out
B="python bench.py --workload ${AB_WORKLOAD:-C1} --no-cpu-baseline --no-extras --steps 5 --warmup 3 --pretrain-ticks ${AB_PRETRAIN:-20000} --e2e-steps 10 ${AB_FLAGS:-}"
for spec in "$@"; do
  name=${spec%%:*}
  vars=""
  [ "$spec" != "$name" ] && vars=$(echo "${spec#*:}" | tr ',' ' ')
  ( env $vars $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err )
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/ab_%s.json' % n).read().strip().splitlines()[-1])
    rf = d['roofline']
    print("AB %-22s value %.4e e2e %.4e ms/step %.2f dom %.1fus (%.0f steps) other %.1fus" % (
        n, d['value'], d.get('e2e', {}).get('value', 0), d['ms_per_step'], 1e3 * (rf.get('avg_launch_ms') or 0),
        rf.get('env_steps_per_launch', 0), 1e3 * ((rf.get('other_kernel') or {}).get('avg_launch_ms') or 0)))
except Exception as e:
    print("AB %s FAILED %s" % (n, e)); print(open('gpurun_out/ab_%s.err' % n).read()[-800:])
PY
done
