#!/bin/bash
# usage: bash tools/limit_pipe_sweep.sh TAG "<so variants>" "<pipe shapes>" -- compile-time variants of the pipelined Limit rollout (librlcard_b200_<v>.so
# built by hand from tu_limit.cu with -DRLC_PIPE_* flags; "-" = the default library): parity test first (a variant that fails or hangs is skipped),
# then two timing rounds per shape; every command under its own short timeout
set -u
TAG=$1; VARS=${2:-"-"}; PIPES=${3:-"31"}
OUT=gpurun_out; mkdir -p $OUT
for v in $VARS; do
  if [ "$v" = "-" ]; then unset RLC_SO_VARIANT; else export RLC_SO_VARIANT=$v; fi
  timeout 60 python -m pytest tests -m gpu -x -q -k "limit_pipelined" 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "variant=$v: parity failed or timed out, skipped"; continue; fi
  for r in 1 2; do for pipe in $PIPES; do
    RLC_LIMIT_PIPE=$pipe timeout 40 python bench.py --game limit-holdem --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/limit_pipe_err.log | tail -n 1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant=$v pipe=$pipe', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))" 2>/dev/null || echo "variant=$v pipe=$pipe: no line"
  done; done
done
