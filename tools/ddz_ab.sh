#!/bin/bash
# DouDizhu: bulk-copy (UBLKCP) mask / obs rows in the warp-per-env rollout vs LDS+STG loops.  RLC_WROLLOUT_BULK = 0 | 1 | 3
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
for v in 1 3; do
  RLC_WROLLOUT_BULK=$v timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_dmc.py tests/test_reference_kats.py -m gpu -x -q -k "doudizhu" > $OUT/pytest_ddz_bulk${v}_$TAG.log 2>&1; echo "bulk=$v parity rc=$?"; tail -2 $OUT/pytest_ddz_bulk${v}_$TAG.log
done
for round in 1 2; do
for v in 0 1 3; do
  RLC_WROLLOUT_BULK=$v timeout 300 python bench.py --game doudizhu --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/ddz_ab_err.log | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bulk=$v', d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])"
done
done
