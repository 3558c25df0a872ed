#!/bin/bash
# usage: bash tools/gpu_one.sh TAG game [bench flags...]  -- refresh the evidence of ONE game inside pass TAG (same commands as gpu_round.sh)
set -u
TAG=$1; g=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --game $g --steps 50 --warmup 5 "$@" > $OUT/bench_${g}_$TAG.json 2> $OUT/bench_${g}_$TAG.err; echo "bench $g rc=$?"
CMD="python bench.py --game $g --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0"
$CMD > $OUT/plain_${g}_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${g}_$TAG.csv $CMD > $OUT/ncu_launches_${g}_$TAG.log 2>&1
echo "launch list $g rc=$?"
$CMD > $OUT/plain2_${g}_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout -s 3 -c 1 -f -o $OUT/prof_${g}_$TAG $CMD > $OUT/ncu_full_${g}_$TAG.log 2>&1
echo "full capture $g rc=$?"
