#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r2b_c1.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python bench.py --workload C4 --no-e2e --no-cpu-baseline --no-extras --steps 5 --warmup 3 > gpurun_out/bench_r2b_c4_1gpu.json 2> gpurun_out/bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b_c1.json').read().strip().splitlines()[-1])
print("C1 value %.4e e2e %.4e C2 %.4e" % (d['value'], d['e2e']['value'], d['extra']['C2']['value']))
c=json.loads(open('gpurun_out/bench_r2b_c4_1gpu.json').read().strip().splitlines()[-1])
print("C4 value %.4e" % c['value'])
PY
