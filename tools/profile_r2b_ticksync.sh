#!/bin/bash
# tick-synchronous engine at C1 (what run calls shorter than 128 ticks -- the end-to-end leg -- use): launch list + full captures
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 4 --warmup 3 --pretrain-ticks 20000 --ticks 64"
RLM_ROUNDS=0 RLM_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 41000 -c 200 --csv --log-file gpurun_out/launches_r2b_c1_ticksync.csv \
  $B > gpurun_out/ncu_c1_launches_ticksync.log 2>&1
RLM_ROUNDS=0 RLM_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:"rlm_learn_kernel|rlm_env_kernel_w" --launch-skip 41000 -c 2 -f -o gpurun_out/r2b_c1_ticksync_full \
  $B > gpurun_out/ncu_c1_ticksync_full.log 2>&1
ls -la gpurun_out/ | grep -i "ticksync"
