#!/bin/bash
# halo conv kernel variants: parity + launch table
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "conv" 2>&1 | tail -2
for v in 1 2; do
  echo "== RHO_HALO_ISSUERS=$v"
  RHO_HALO_ISSUERS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --trajectory-steps 0 --fp32-steps 0 --ddim-steps 0 \
    --launch-table gpurun_out/lt_iss_$v.csv 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved_executed'])"
  grep "conv3d_k3_64->64\|conv3d_k3_128->128,\|conv3d_k3_192->64\|conv3d_k3_128->64\|conv3d_k1_128->64\|conv3d_k3_256->128" gpurun_out/lt_iss_$v.csv
done
