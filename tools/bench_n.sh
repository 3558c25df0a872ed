#!/bin/bash
# usage: bash tools/bench_n.sh TAG N  -- the driver's multi-GPU launch of bench.py on N GPUs of one box
TAG=$1; N=$2
OUT=gpurun_out; mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err
echo "bench n=$N rc=$?"; tail -c 600 $OUT/bench_n${N}_$TAG.json
