#!/bin/bash
set -u
mkdir -p gpurun_out
RHO_ATTN_V2_MIN_T=1 timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "test_attention_fwd_bwd" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "test_attention_fwd_bwd" 2>&1 | tail -2
for v in "RHO_ATTN_LEAN=0" "RHO_ATTN_LEAN=1" "RHO_ATTN_LEAN=1 RHO_ATTN_POLY=2" "RHO_ATTN_LEAN=1 RHO_ATTN_POLY=4"; do echo "== $v"; env $v timeout 120 python tools/bench_attn.py 2>&1 | grep "fwd N=8\|fwd N=2 T=16384"; done
