#!/bin/bash
# DouDizhu after the reset rewrite (parity + timing) and the batch-size scaling of the thread-per-env rollouts
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "doudizhu" 2>&1 | tail -2
one() { python bench.py "$@" --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/misc_err.log | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), '%.3e' % d['value'])"; }
one --game doudizhu
one --game doudizhu
for n in 16384 32768 65536 131072; do one --game limit-holdem --envs $n; done
for n in 16384 65536; do RLC_LIMIT_FSM=0 one --game limit-holdem --envs $n; done
for n in 16384 32768 65536; do one --game uno --envs $n; done
for n in 8192 16384; do one --game scout --envs $n; one --game doudizhu --envs $n; done
