# usage: bash tools/sweep_epw.sh TAG game [game ...]  -- kernel-only bench line per envs-per-warp setting of the fused rollout
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for g in "$@"; do
  for epw in 32 16 8; do
    RLC_ROLLOUT_EPW=$epw python bench.py --game $g --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 0 --e2e-step-api-steps 0 \
      > $OUT/epw_${g}_${epw}_$TAG.json 2> $OUT/epw_${g}_${epw}_$TAG.err
    echo "$g epw=$epw rc=$? $(python -c "import json,sys; d=json.loads(open('$OUT/epw_${g}_${epw}_$TAG.json').read().strip().splitlines()[-1]); print('%.4f ms  %.3e steps/s  frac %.3f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))" 2>&1 | tail -1)"
  done
done
