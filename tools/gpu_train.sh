#!/bin/bash
# training-step kernels: parity of the conv backward + per-kernel table of one step (batch 4)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "conv_fwd_bwd" 2>&1 | tail -2
for v in "RHO_WGRAD_HALO_LEAN=0" "RHO_WGRAD_HALO_LEAN=1"; do
  echo "== $v"; env $v timeout 300 python tools/prof_step.py train 4 2>&1 | grep -v Warn | head -14
done
