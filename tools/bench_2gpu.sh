#!/bin/bash
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print("N=2 value %.4e e2e %.4e" % (d['value'], d.get('e2e',{}).get('value',0)))
for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('error'))
PY
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; tail -2 gpurun_out/bench_2gpu_ref.json | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_shared_policy.py -x -q -m gpu 2>&1 | tail -2
