#!/bin/bash
# One GPU-box pass: parity tests, smoke, per-game bench lines, reference arm, ncu launch lists + full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag] [games to profile ...]
# gpurun merges back at most 64 MiB per call and a full capture is 6-17 MB: profile at most three games here and the rest
# with tools/prof_games.sh in a second call.
set -u
TAG=${1:-r01}
shift || true
PROF_GAMES=${*:-leduc-holdem doudizhu}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "bench default rc=$?"
for g in blackjack no-limit-holdem; do      # the two games outside BASELINE configs 2-5 (those ride in the default line's `configs`)
  python bench.py --game $g --steps 50 --warmup 5 --cpu-seconds 5 > $OUT/bench_${g}_$TAG.json 2> $OUT/bench_${g}_$TAG.err; echo "bench $g rc=$?"
done
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_$TAG.json 2>&1; echo "reference rc=$?"
for g in $PROF_GAMES; do
  CMD="python bench.py --game $g --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0"
  $CMD > $OUT/plain_${g}_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${g}_$TAG.csv $CMD > $OUT/ncu_launches_${g}_$TAG.log 2>&1
  echo "launch list $g rc=$?"
  $CMD > $OUT/plain2_${g}_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rollout -s 3 -c 1 -f -o $OUT/prof_${g}_$TAG $CMD > $OUT/ncu_full_${g}_$TAG.log 2>&1
  echo "full capture $g rc=$?"
done
