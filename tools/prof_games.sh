#!/bin/bash
# usage: bash tools/prof_games.sh TAG game [game ...]  -- ncu launch list + one `--set full` capture of the rollout kernel per game
# (at most ~5 games per gpurun call: the captures are 8-14 MB each and gpurun merges back at most 64 MiB)
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for g in "$@"; do
  CMD="python bench.py --game $g --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0"
  $CMD > $OUT/plain_${g}_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${g}_$TAG.csv $CMD > $OUT/ncu_launches_${g}_$TAG.log 2>&1
  echo "launch list $g rc=$?"
  $CMD > $OUT/plain2_${g}_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rollout -s 3 -c 1 -f -o $OUT/prof_${g}_$TAG $CMD > $OUT/ncu_full_${g}_$TAG.log 2>&1
  echo "full capture $g rc=$?"
done
