#!/bin/bash
# usage: bash tools/condense.sh TAG   -- condense gpurun_out/*_TAG.* (tools/gpu_round.sh) into profiles/TAG_* (run where ncu is installed)
set -u
TAG=$1
OUT=gpurun_out
P=profiles
declare -A KREG=( [leduc-holdem]='k_rollout_leduc_fsmIhLi512E' [limit-holdem]='k_rollout_limit_pipeIhLi3ELi1' [uno]='k_rollout.*UnoT.*Philox.*EhLi64ELb1ELi16'
                  [doudizhu]='k_wrollout.*Doudizhu.*Philox.*EhLi128ELb0ELi3' [scout]='k_wrollout_multi.*ScoutTILi8EEENS_12ChancePhiloxEfLi448ELi8' [blackjack]='k_rollout.*Blackjack.*Philox.*EhLi64ELb1ELi32'
                  [no-limit-holdem]='k_rollout.*NoLimit.*Philox.*EhLi64ELb1ELi32' )
declare -A OBJ=( [leduc-holdem]=tu_leduc [limit-holdem]=tu_limit [uno]=tu_uno [doudizhu]=tu_doudizhu [scout]=tu_scout [blackjack]=tu_blackjack [no-limit-holdem]=tu_nolimit )
declare -A KEY=( [leduc-holdem]=leduc-holdem:uint8:65536:128 [limit-holdem]=limit-holdem:uint8:16384:128 [uno]=uno:uint8:16384:128 [doudizhu]=doudizhu:uint8:8192:128
                 [scout]=scout:float32:8192:128 [blackjack]=blackjack:uint8:65536:128 [no-limit-holdem]=no-limit-holdem:uint8:16384:128 )
: > $P/${TAG}_bench_lines.jsonl
tail -n 1 $OUT/bench_default_$TAG.json >> $P/${TAG}_bench_lines.jsonl
for g in blackjack no-limit-holdem; do tail -n 1 $OUT/bench_${g}_$TAG.json >> $P/${TAG}_bench_lines.jsonl; done   # configs 3-5 ride in the default line's `configs`
tail -n 1 $OUT/bench_reference_$TAG.json >> $P/${TAG}_bench_lines.jsonl
cp $OUT/pytest_gpu_$TAG.log $P/${TAG}_pytest_gpu.log
cp $OUT/smoke_$TAG.log $P/${TAG}_smoke.log
for g in "${!KREG[@]}"; do
  [ -f $OUT/prof_${g}_$TAG.ncu-rep ] || continue
  python tools/ncu_summary.py full $OUT/prof_${g}_$TAG.ncu-rep > $P/${TAG}_${g}_full.txt
  python tools/ncu_summary.py launches $OUT/launches_${g}_$TAG.csv > $P/${TAG}_${g}_launches.txt
  python tools/ncu_lines.py $OUT/prof_${g}_$TAG.ncu-rep rlcard_b200/csrc/${OBJ[$g]}.o "${KREG[$g]}" 40 > $P/${TAG}_${g}_lines.txt 2>&1
  python tools/ncu_traffic.py $OUT/prof_${g}_$TAG.ncu-rep ${KEY[$g]}
  head -3 $P/${TAG}_${g}_lines.txt | tail -2
done
