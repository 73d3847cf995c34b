#!/bin/bash
# two-GPU checks: NCCL data-parallel trainer (eager + CUDA-graph), sharded sampler, bench at N = 2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
  --trajectory-steps 20 --train-steps 20 --train-warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/bench_r02_n2.err; cut -c1-400 gpurun_out/bench_r02_n2.json
