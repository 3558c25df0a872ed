#!/bin/bash
# compact wire formats of UNO / DouDizhu / Scout: parity of the host expansion, then the e2e legs (dense vs compact)
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q -k "compact" 2>&1 | tail -3
for g in uno doudizhu scout; do
  python bench.py --game $g --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 --e2e-step-api-steps 0 2>> $OUT/compact_err.log | tail -n 1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$g dense e2e %.3e (%.1f GB/s)  compact %.3e (%.1f GB/s, %d B/step)' % (e['value'], e['pcie_gbs'], e['compact']['value'], e['compact']['pcie_gbs'], e['compact']['d2h_bytes_per_step']*1.0/(d['config']['envs_per_gpu']*d['config']['env_steps_per_launch_per_env'])))"
done
