#!/bin/bash
# wide 1x1 convs through the halo kernel's streaming mode: parity + launch table A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_data.py -q -x -m gpu -k "conv or line_profiles" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_model.py -q -x -m gpu -k "forward" 2>&1 | tail -2
for v in 0 1; do
  echo "== RHO_HALO_WIDE_K1=$v"
  RHO_HALO_WIDE_K1=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --trajectory-steps 0 --fp32-steps 0 --ddim-steps 0 \
    --launch-table gpurun_out/lt_k1_$v.csv 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])"
  grep "conv1d_k1" gpurun_out/lt_k1_$v.csv
done
