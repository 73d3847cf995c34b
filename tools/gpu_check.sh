#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (B=2, B=8), ncu launch list and one --set full capture per hot kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --batch 2 --launch-table gpurun_out/launch_table_b2.csv > gpurun_out/bench_b2.json 2> gpurun_out/bench_b2.err; echo "bench b2 rc=$?"
python bench.py --batch 8 --steps 6 --no-cpu-baseline --launch-table gpurun_out/launch_table_b8.csv > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench b8 rc=$?"
cat gpurun_out/bench_b2.json gpurun_out/bench_b8.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ncu_launches_b2.csv \
  python bench.py --batch 2 --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 6 -o gpurun_out/ncu_conv_tc -f \
  python bench.py --batch 2 --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -c 1 -o gpurun_out/ncu_attn_tc -f \
  python bench.py --batch 2 --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ -s 10 -c 4 -o gpurun_out/ncu_gn -f \
  python bench.py --batch 2 --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn rc=$?"
ls -la gpurun_out
