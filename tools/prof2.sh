# usage: bash tools/prof2.sh TAG game [game ...]   -- one `ncu --set full` capture of the rollout kernel per game
set -u
TAG=$1; shift
OUT=gpurun_out
for g in "$@"; do
  CMD="python bench.py --game $g --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0"
  $CMD > $OUT/plain_${g}_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rollout -s 3 -c 1 -f -o $OUT/prof_${g}_$TAG $CMD > $OUT/ncu_full_${g}_$TAG.log 2>&1
  echo "full capture $g rc=$?"
done
