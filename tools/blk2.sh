#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { g=$1; shift; env "$@" python bench.py --game $g --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/blk2_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$g $*', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"; }
python -m pytest tests -m gpu -x -q -k "leduc" 2>&1 | tail -2
run leduc-holdem A=1
run leduc-holdem A=1
for n in 8192 32768 131072 100000; do g=leduc-holdem; python bench.py --game $g --envs $n --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/blk2_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('leduc envs $n', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"; RLC_LEDUC_BLOCK=64 python bench.py --game $g --envs $n --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/blk2_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('leduc envs $n block 64', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"; done
for b in 64 128 256 448 512; do run scout RLC_WROLLOUT_BLOCK=$b; done
RLC_WROLLOUT_BLOCK=512 python -m pytest tests -m gpu -x -q -k "scout_lanes" 2>&1 | tail -2
