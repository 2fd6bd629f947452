#!/bin/bash
# Round-2 ncu captures (run under gpurun, one GPU).  Outputs land in gpurun_out/; summaries are made from them by
# tools/ncu_summary.py in the build container and committed under profiles/.
set -x
B="python bench.py --no-e2e --no-cpu-baseline --no-extras"
# C1: launch list at steady state (tick-synchronous engine, direct launches so that every kernel is its own ncu launch)
RLM_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 41000 -c 200 --csv --log-file gpurun_out/launches_r2_c1.csv \
  $B --steps 4 --warmup 3 --ticks 64 --pretrain-ticks 20000 > gpurun_out/ncu_c1_launches.log 2>&1
# C1: full sections of one learner launch and one tick launch
RLM_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:"rlm_learn_kernel|rlm_env_kernel_w" --launch-skip 41000 -c 2 -f -o gpurun_out/r2_c1_full \
  $B --steps 4 --warmup 3 --ticks 64 --pretrain-ticks 20000 > gpurun_out/ncu_c1_full.log 2>&1
# C4: the TMA-staged learner and the thread-per-env tick kernel
RLM_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:"rlm_learn_staged_kernel|rlm_env_kernel" --launch-skip 3000 -c 2 -f -o gpurun_out/r2_c4_full \
  $B --workload C4 --steps 3 --warmup 3 --ticks 16 --pretrain-ticks 2400 > gpurun_out/ncu_c4_full.log 2>&1
RLM_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 60 --csv --log-file gpurun_out/launches_r2_c4.csv \
  $B --workload C4 --steps 3 --warmup 3 --ticks 16 --pretrain-ticks 2400 > gpurun_out/ncu_c4_launches.log 2>&1
ls -la gpurun_out/
