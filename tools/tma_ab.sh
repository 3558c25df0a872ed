#!/bin/bash
# A/B of the tile-flush implementation on one B200: LDS+STG loops (default build) vs cp.async.bulk (TMA=1 build).
# usage (under gpurun, from the repo root): bash tools/tma_ab.sh TAG
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
RLC_SO_VARIANT=tma python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "oracle or full_size or stale or partial" > $OUT/pytest_tma_$TAG.log 2>&1
echo "tma parity rc=$?"; tail -2 $OUT/pytest_tma_$TAG.log
: > $OUT/tma_ab_$TAG.jsonl
for round in 1 2; do
for g in leduc-holdem limit-holdem uno doudizhu scout no-limit-holdem blackjack; do
  for v in base tma; do
    if [ $v = tma ]; then export RLC_SO_VARIANT=tma; else unset RLC_SO_VARIANT; fi
    python bench.py --game $g --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2> $OUT/tma_ab_err.log | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'game':'$g','variant':'$v','round':$round,'kernel_ms':d['roofline']['kernel_ms'],'frac':d['roofline']['frac'],'value':d['value'],'sm_mhz':d['clocks']['sm_mhz']}))" >> $OUT/tma_ab_$TAG.jsonl
  done
done
done
unset RLC_SO_VARIANT
cat $OUT/tma_ab_$TAG.jsonl
