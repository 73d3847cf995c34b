#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_scripts.py -m gpu -q -x -k "independent or bounded or script" 2>&1 | tail -40 > gpurun_out/pytest_fail.log; cat gpurun_out/pytest_fail.log | tail -40
# whole-eval DRAM traffic / tensor-pipe activity per kernel (4 metrics: 1-2 replay passes per launch)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --page raw --log-file gpurun_out/ncu_eval_b8_raw.csv python tools/one_eval.py 8 > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
python tools/ncu_traffic.py gpurun_out/ncu_eval_b8_raw.csv gpurun_out/ncu_traffic_b8.json | head -24
# --set full of the three dominant kernels (two launches each)
for k in conv_tc_kernel conv_halo_kernel attn_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 8 -c 2 -o gpurun_out/ncu_full_$k -f python tools/one_eval.py 8 > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu $k rc=$?"
  ncu -i gpurun_out/ncu_full_$k.ncu-rep --page raw --csv > gpurun_out/ncu_full_${k}_raw.csv 2>/dev/null
done
python tools/prof_step.py train 4 > gpurun_out/prof_train_b4.txt 2>&1; head -50 gpurun_out/prof_train_b4.txt
du -sh gpurun_out
