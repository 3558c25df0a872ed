#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { env "$@" python bench.py --game leduc-holdem --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/leduc_blk_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"; }
RLC_LEDUC_BLOCK=512 python -m pytest tests -m gpu -x -q -k "leduc" 2>&1 | tail -2
for b in 256 384 512 640 768 1024; do run RLC_LEDUC_BLOCK=$b; done
for b in 64 256 512 1024; do run RLC_LEDUC_BLOCK=$b RLC_LEDUC_ONE_PER_SM=0; done
for b in 64 128 256; do run RLC_LEDUC_BLOCK=$b RLC_LEDUC_ONE_PER_SM=1; done
run RLC_LEDUC_BLOCK=512
