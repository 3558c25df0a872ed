#!/bin/bash
# usage: bash tools/limit_fsm_ab.sh TAG  -- Limit Hold'em throughput rollout: generic register engine (RLC_LIMIT_FSM=0) vs tabulated engine, one warp (1) / two warps (2) per group
set -u
TAG=$1
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "limit or holdem or judge or kat" > $OUT/pytest_limit_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_limit_$TAG.log
tail -3 $OUT/pytest_limit_$TAG.log
for r in 1 2; do
  for v in 0 1 2; do
    RLC_LIMIT_FSM=$v python bench.py --game limit-holdem --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/limit_ab_err.log | tail -n 1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fsm=$v', d['ms_per_step'], d['roofline']['frac'], d['value'])"
  done
done
