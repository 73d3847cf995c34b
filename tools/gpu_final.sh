#!/bin/bash
# End-of-round evidence (round 2): full GPU parity suite, smoke(), default bench line, reference arm, launch table, ncu launch list,
# whole-eval DRAM traffic pass, --set full captures of the dominant kernels, training-step kernel breakdown.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --launch-table gpurun_out/r02_launch_table_b8.csv > gpurun_out/r02_bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_b8.json
RHO_BENCH_CPU_BUDGET_S=60 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_ncu_launch_list_b8.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0 --trajectory-steps 0 --fp32-steps 0 --ddim-steps 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --page raw --log-file gpurun_out/r02_ncu_eval_b8_per_launch.csv python tools/one_eval.py 8 > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
python tools/ncu_traffic.py gpurun_out/r02_ncu_eval_b8_per_launch.csv gpurun_out/r02_ncu_traffic_b8.json | head -16
for k in conv_tc_kernel conv_halo_kernel attn_tc2_kernel gn_apply8_kernel; do
  skip=8; [ $k = attn_tc2_kernel ] && skip=1
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s $skip -c 2 -o gpurun_out/r02_ncu_full_$k -f python tools/one_eval.py 8 > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu $k rc=$?"
  ncu -i gpurun_out/r02_ncu_full_$k.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${k}_raw.csv 2>/dev/null
done
python tools/prof_step.py train 4 > gpurun_out/r02_train_step_b4_kernels.txt 2>&1
du -sh gpurun_out
