#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "doudizhu or dmc or rl_ or kat" 2>&1 | tail -2
for r in 1 2; do python bench.py --game doudizhu --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/ddz_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('doudizhu', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; done
