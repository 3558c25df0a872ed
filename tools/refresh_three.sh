set -u
TAG=r01k
OUT=gpurun_out
for g in blackjack uno; do python bench.py --game $g --steps 50 --warmup 5 > $OUT/bench_${g}_$TAG.json 2> $OUT/bench_${g}_$TAG.err; echo "bench $g rc=$?"; done
python bench.py --game scout --steps 50 --warmup 5 --e2e-steps 3 --dmc-steps 10 > $OUT/bench_scout_$TAG.json 2> $OUT/bench_scout_$TAG.err; echo "bench scout rc=$?"
for g in blackjack uno scout; do
  CMD="python bench.py --game $g --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0"
  $CMD > $OUT/plain_${g}_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${g}_$TAG.csv $CMD > $OUT/ncu_launches_${g}_$TAG.log 2>&1
  $CMD > $OUT/plain2_${g}_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout -s 3 -c 1 -f -o $OUT/prof_${g}_$TAG $CMD > $OUT/ncu_full_${g}_$TAG.log 2>&1
  echo "profile $g rc=$?"
done
