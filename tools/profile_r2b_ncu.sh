#!/bin/bash
# ncu launch list + full captures of the round-paced engine at C1 (final code of the round); see tools/profile_r2b.sh
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 4 --warmup 3 --pretrain-ticks 20000"
RLM_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 21000 -c 240 --csv --log-file gpurun_out/launches_r2b_c1_rounds.csv \
  $B --ticks 256 > gpurun_out/ncu_c1_launches_rounds.log 2>&1
RLM_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:"rlm_learn_kernel|rlm_env_round_kernel" --launch-skip 21000 -c 2 -f -o gpurun_out/r2b_c1_rounds_full \
  $B --ticks 256 > gpurun_out/ncu_c1_rounds_full.log 2>&1
ls -la gpurun_out/ | grep -i rounds
