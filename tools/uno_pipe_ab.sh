#!/bin/bash
# UNO: chunked ENV / EMIT rollout (RLC_UNO_PIPE=16|32) vs the generic thread-per-env kernel; every command under its own timeout
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python -m pytest tests -m gpu -x -q -k "uno_chunked" 2>&1 | tail -1
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "parity failed or timed out"; exit 1; fi
for round in 1 2; do
for v in 0 32 16; do
  RLC_UNO_PIPE=$v timeout 40 python bench.py --game uno --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/uno_pipe_err.log | tail -n 1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipe=$v', round(d['roofline']['kernel_ms'],5), round(d['roofline']['frac'],4))" 2>/dev/null || echo "pipe=$v: no line"
done
done
