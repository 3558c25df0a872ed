#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { g=$1; shift; env "$@" python bench.py --game $g --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/blk3_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$g $*', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"; }
for sy in 0 1 2 4 8 16 32; do run leduc-holdem RLC_LEDUC_SYNC=$sy; done
for sy in 0 4 16; do run leduc-holdem RLC_LEDUC_SYNC=$sy RLC_LEDUC_BLOCK=64; run leduc-holdem RLC_LEDUC_SYNC=$sy RLC_LEDUC_BLOCK=256; done
