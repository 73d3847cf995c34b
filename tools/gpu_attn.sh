#!/bin/bash
set -u
for v in "RHO_ATTN_BWD_DBG=0" "RHO_ATTN_BWD_DBG=1" "RHO_ATTN_BWD_DBG=2" "RHO_ATTN_BWD_DBG=3"; do echo "== $v"; env $v timeout 120 python tools/bench_attn.py 2>&1 | grep "bwd N=8"; done
