#!/bin/bash
# UNO: two-warp ENV / EMIT rollout (RLC_UNO_EE=16|32) vs the generic thread-per-env kernel
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
for v in 16 32; do
  RLC_UNO_EE=$v timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "uno" > $OUT/pytest_uno_ee${v}_$TAG.log 2>&1; echo "ee=$v parity rc=$?"; tail -2 $OUT/pytest_uno_ee${v}_$TAG.log
done
for round in 1 2; do
for v in 0 16 32; do
  RLC_UNO_EE=$v timeout 300 python bench.py --game uno --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/uno_ab_err.log | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ee=$v', d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])"
done
done
