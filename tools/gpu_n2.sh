#!/bin/bash
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --batch 8 --train-batch 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python bench.py --gpus 1 --steps 6 --warmup 3 --batch 8 --train-batch 2 --no-cpu-baseline > gpurun_out/bench_n1_tb2.json 2> gpurun_out/bench_n1_tb2.err; echo "n1 rc=$?"
cat gpurun_out/bench_n1_tb2.json
python bench.py --gpus 1 --steps 6 --warmup 3 --batch 8 --train-batch 4 --no-cpu-baseline > gpurun_out/bench_n1_tb4.json 2> gpurun_out/bench_n1_tb4.err; echo "n1 rc=$?"
cat gpurun_out/bench_n1_tb4.json
RHO_BENCH_CPU_BUDGET_S=60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
