#!/bin/bash
set -u
mkdir -p gpurun_out
RHO_ATTN_BWD_V2_MIN_T=1 timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "test_attention_fwd_bwd" 2>&1 | tail -3
for v in "RHO_ATTN_BWD_V2=1"; do echo "== $v"; env $v timeout 120 python tools/prof_attn.py 8 2>&1 | grep "ms x" | head -3; done
