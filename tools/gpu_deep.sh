#!/bin/bash
# deep operand ring for launches with <= 1 CTA per SM + batched partial-sum loads: parity, then batch 1 / 2 / 8 A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_switches.py -q -x -m gpu 2>&1 | tail -2
Q="--steps 8 --warmup 3 --no-cpu-baseline --train-steps 0 --ddim-steps 0 --trajectory-steps 0 --fp32-steps 0"
for v in 0 1; do for b in 1 2 8; do
  RHO_CONV_DEEP_RING=$v timeout 300 python bench.py --batch $b $Q --launch-table gpurun_out/lt_deep${v}_b$b.csv 2>/dev/null \
    | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deep=$v batch $b', round(d['value'], 1), 'evals/s', round(d['ms_per_step'], 2), 'ms', d['clocks']['sm_mhz'], 'conv frac', round(d['roofline']['frac'], 3))"
done; done
grep "512->512\|256->256\|gn_stats_fused" gpurun_out/lt_deep0_b1.csv gpurun_out/lt_deep1_b1.csv
