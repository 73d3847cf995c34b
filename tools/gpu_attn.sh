#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "RHO_ATTN_POLY=0" "RHO_ATTN_POLY=3" "RHO_ATTN_POLY=0 RHO_ATTN_PVFIRST=1" "RHO_ATTN_POLY=3 RHO_ATTN_PVFIRST=1"; do
  echo "== $v"; env $v timeout 120 python tools/bench_attn.py 2>&1 | grep "fwd N=8"
done
python tools/run_attn_fwd.py 8 3 && ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 1 -c 1 -o gpurun_out/r02_attn2_p0 -f python tools/run_attn_fwd.py 8 3 > gpurun_out/ncu_a.log 2>&1
RHO_ATTN_POLY=3 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 1 -c 1 -o gpurun_out/r02_attn2_p3 -f python tools/run_attn_fwd.py 8 3 > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
