#!/bin/bash
# usage: bash tools/limit_pipe_sweep.sh TAG -- compile-time variants of the pipelined Limit rollout (librlcard_b200_<v>.so built by hand with
# -DRLC_PIPE_RING / -DRLC_PIPE_CHUNK / -DRLC_PIPE_POLICY), each with parity test + timing at RLC_LIMIT_PIPE=31 and 42
set -u
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
for v in "" pol r20 r24 c8 c32; do
  export RLC_SO_VARIANT=$v
  timeout 200 python -m pytest tests -m gpu -x -q -k "limit_pipelined" 2>&1 | tail -1
  for pipe in 31 42; do
    RLC_LIMIT_PIPE=$pipe timeout 120 python bench.py --game limit-holdem --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/limit_pipe_err.log | tail -n 1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant=$v pipe=$pipe', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"
  done
done
