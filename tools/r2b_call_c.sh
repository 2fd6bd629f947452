#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print("value %.4e e2e %.4e C2 %.4e" % (d['value'], d['e2e']['value'], d['extra']['C2']['value']))
print(json.dumps(d['roofline'])[:1500])
PY
