#!/bin/bash
# usage: bash tools/prof_env.sh TAG game VAR=VAL ...  -- one ncu --set full capture of the rollout kernel with env overrides
TAG=$1; g=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --game $g --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0"
env "$@" $CMD > $OUT/plain_${g}_$TAG.log 2>&1 && env "$@" ncu --set full --clock-control none --import-source on -k regex:rollout -s 3 -c 1 -f -o $OUT/prof_${g}_$TAG $CMD > $OUT/ncu_full_${g}_$TAG.log 2>&1
echo "full capture $g rc=$?"
