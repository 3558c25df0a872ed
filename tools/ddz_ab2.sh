set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r02q.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_r02q.log
for v in 0 3 0 3; do
  RLC_WROLLOUT_BULK=$v python bench.py --game doudizhu --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 2>> $OUT/ddz_ab_err.log | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bulk=$v', d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])"
done
