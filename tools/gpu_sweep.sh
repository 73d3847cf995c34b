#!/bin/bash
# batch sweep of the config-D sampling step + determinism stress (final build)
set -u
mkdir -p gpurun_out
for b in 1 2 4 16; do
  timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline --train-steps 0 --ddim-steps 0 --trajectory-steps 0 --fp32-steps 0 2>/dev/null \
    | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $b', round(d['value'], 1), 'evals/s', round(d['ms_per_step'], 2), 'ms', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'conv frac', round(d['roofline']['frac'], 3), 'executed', round(d['roofline']['frac_executed'], 3))"
done
timeout 400 python tools/stress_determinism.py 8 2>&1 | tail -6
