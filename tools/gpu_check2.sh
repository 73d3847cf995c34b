#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -x -m gpu -k "host_staged or reverse_process" 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --train-steps 0 --trajectory-steps 0 --fp32-steps 0 --ddim-steps 0 2>gpurun_out/bench_quick.err | cut -c1-900
tail -3 gpurun_out/bench_quick.err
