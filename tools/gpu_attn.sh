#!/bin/bash
set -u
mkdir -p gpurun_out
RHO_ATTN_BWD_V2_MIN_T=1 timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "test_attention_fwd_bwd" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "test_attention_fwd_bwd or conv_fwd_bwd" 2>&1 | tail -2
for v in "RHO_ATTN_BWD_LEAN=1" "RHO_ATTN_BWD_LEAN=2"; do echo "== $v"; env $v timeout 120 python tools/bench_attn.py 2>&1 | grep "bwd N=8\|bwd N=2 T=16384"; done
for v in "RHO_WGRAD_HALO_ISSUERS=1" "RHO_WGRAD_HALO_ISSUERS=2"; do
  echo "== $v"; env $v timeout 300 python tools/prof_step.py train 4 2>&1 | grep "total device\|attn_bwd2\|wgrad_halo\|conv_halo"
done
