#!/bin/bash
# Limit FSM rollout: block size x split sweep
OUT=gpurun_out; mkdir -p $OUT
run() { env "$@" python bench.py --game limit-holdem --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/limit_blk_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', d['ms_per_step'], d['roofline']['frac'])"; }
python -m pytest tests -m gpu -x -q -k "limit" 2>&1 | tail -2
RLC_LIMIT_BLOCK=128 python -m pytest tests -m gpu -x -q -k "limit and (oracle or full_size)" 2>&1 | tail -2
for r in 1 2; do
run RLC_LIMIT_FSM=1 RLC_LIMIT_BLOCK=32
run RLC_LIMIT_FSM=1 RLC_LIMIT_BLOCK=64
run RLC_LIMIT_FSM=1 RLC_LIMIT_BLOCK=128
run RLC_LIMIT_FSM=2 RLC_LIMIT_BLOCK=64
run RLC_LIMIT_FSM=2 RLC_LIMIT_BLOCK=128
run RLC_LIMIT_FSM=2 RLC_LIMIT_BLOCK=256
done
