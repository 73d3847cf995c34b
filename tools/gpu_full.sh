#!/bin/bash
# Full GPU check: parity suite, bench lines, ncu launch list + traffic capture of one eval at the bench batch.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|loss curve|config [AD]" gpurun_out/pytest_gpu.log | tail -12
python bench.py --launch-table gpurun_out/launch_table_b8.csv > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench rc=$?"
cat gpurun_out/bench_b8.json
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/ncu_eval_b8 -f \
  python tools/one_eval.py 8 > gpurun_out/ncu_eval.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/ncu_eval_b8.ncu-rep --page raw --csv > gpurun_out/ncu_eval_b8_raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/ncu_eval_b8_raw.csv gpurun_out/ncu_traffic_b8.json | head -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/ncu_launches_b8.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out | head -40
