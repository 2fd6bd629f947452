#!/bin/bash
# Round-2 (second half) evidence run under gpurun, one GPU: full GPU test suite, the default bench line, the reference arm,
# C4 on one GPU, ncu launch lists and full captures at C1 for both engines.  Outputs land in gpurun_out/; summaries are
# made from them by tools/ncu_summary.py / tools/ncu_by_line.py in the build container and committed under profiles/.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2b_c1_reference_arm.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_r2b_c1.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python bench.py --workload C4 --no-e2e --no-cpu-baseline --no-extras --steps 5 --warmup 3 > gpurun_out/bench_r2b_c4_1gpu.json 2> gpurun_out/bench_c4.err
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 4 --warmup 3 --pretrain-ticks 20000"
# round-paced engine (256-tick calls), direct launches so that every kernel is its own ncu launch
RLM_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 21000 -c 240 --csv --log-file gpurun_out/launches_r2b_c1_rounds.csv \
  $B --ticks 256 > gpurun_out/ncu_c1_launches_rounds.log 2>&1
RLM_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:"rlm_learn_kernel|rlm_env_round_kernel" --launch-skip 21000 -c 2 -f -o gpurun_out/r2b_c1_rounds_full \
  $B --ticks 256 > gpurun_out/ncu_c1_rounds_full.log 2>&1
# tick-synchronous engine (64-tick calls: what the end-to-end leg runs)
RLM_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 41000 -c 200 --csv --log-file gpurun_out/launches_r2b_c1_ticksync.csv \
  $B --ticks 64 > gpurun_out/ncu_c1_launches_ticksync.log 2>&1
RLM_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:"rlm_learn_kernel|rlm_env_kernel_w" --launch-skip 41000 -c 2 -f -o gpurun_out/r2b_c1_ticksync_full \
  $B --ticks 64 > gpurun_out/ncu_c1_ticksync_full.log 2>&1
ls -la gpurun_out/ | tail -20
