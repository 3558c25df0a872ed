#!/bin/bash
# usage: bash tools/scout_ab.sh TAG  -- Scout throughput rollout: a warp per env (RLC_WROLLOUT_LPE=32) vs two envs per warp (default)
set -u
TAG=$1
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "scout" > $OUT/pytest_scout_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_scout_$TAG.log
tail -3 $OUT/pytest_scout_$TAG.log
for r in 1 2; do
  for v in 32 16 8; do
    RLC_WROLLOUT_LPE=$v python bench.py --game scout --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/scout_ab_err.log | tail -n 1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lpe=$v', d['ms_per_step'], d['roofline']['frac'], d['value'])"
  done
done
