#!/bin/bash
# Limit Hold'em: parity of the warp-specialised rollout, then A/B against the generic kernel (RLC_LIMIT_GENERIC=1)
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "limit" > $OUT/pytest_limit_$TAG.log 2>&1; echo "limit parity rc=$?"; tail -3 $OUT/pytest_limit_$TAG.log
for v in ws generic ws generic; do
  if [ $v = generic ]; then export RLC_LIMIT_GENERIC=1; else unset RLC_LIMIT_GENERIC; fi
  timeout 300 python bench.py --game limit-holdem --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/limit_ab_err.log | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])"
done
unset RLC_LIMIT_GENERIC
