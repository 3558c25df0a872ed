#!/bin/bash
# usage: bash tools/limit_pipe_ab.sh TAG [variants...] -- Limit Hold'em: default two-warp tabulated rollout vs the role-warp pipeline (RLC_LIMIT_PIPE=<EMIT><DEAL>)
set -u
TAG=$1; shift
VARS=${*:-"0 32 22 42 31"}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "limit_pipelined" > $OUT/pytest_limit_pipe_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_limit_pipe_$TAG.log
tail -3 $OUT/pytest_limit_pipe_$TAG.log
for r in 1 2; do
  for v in $VARS; do
    RLC_LIMIT_PIPE=$v timeout 120 python bench.py --game limit-holdem --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/limit_pipe_err.log | tail -n 1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe=$v', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['value'])"
  done
done
