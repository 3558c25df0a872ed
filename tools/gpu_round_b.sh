#!/bin/bash
# Second half of an evidence pass (gpurun merges back at most 64 MiB per call): the remaining ncu captures, then the bench lines again
# on the final build.  Usage: bash tools/gpu_round_b.sh TAG game [game ...]
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
bash tools/prof_games.sh $TAG "$@"
python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "bench default rc=$?"
