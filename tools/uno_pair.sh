#!/bin/bash
set -u
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "uno" > $OUT/pytest_uno_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_uno_$TAG.log
tail -3 $OUT/pytest_uno_$TAG.log
for r in 1 2; do
  python bench.py --game uno --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/uno_err.log | tail -n 1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('uno', d['ms_per_step'], d['roofline']['frac'], d['value'])"
done
