#!/bin/bash
# End-of-round re-verification after the late changes (pointwise fold, deep ring, statistics kernel choice): full GPU suite, smoke(),
# default bench line + launch table, ncu launch list.  The --set full captures / per-launch traffic pass of tools/gpu_final.sh stay valid
# (the dominant kernels did not change).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --launch-table gpurun_out/r02_launch_table_b8.csv > gpurun_out/r02_bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_b8.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_ncu_launch_list_b8.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 --trajectory-steps 0 --fp32-steps 0 --ddim-steps 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
