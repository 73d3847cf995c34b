#!/bin/bash
# config E (128^3, batch 1 per GPU): sampling step + a 20-step trajectory through DDPM.generate(), N = 1 or torchrun N = $1
set -u
mkdir -p gpurun_out
N=${1:-1}
ARGS="--config E --steps 10 --warmup 3 --trajectory-steps 20 --train-steps 0 --fp32-steps 0 --ddim-steps 0 --no-cpu-baseline"
if [ "$N" = 1 ]; then
  timeout 600 python bench.py $ARGS > gpurun_out/r02_bench_cfgE_n1.json 2> gpurun_out/bench_cfgE_n1.err; echo "bench E n1 rc=$?"
else
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N $ARGS \
    > gpurun_out/r02_bench_cfgE_n$N.json 2> gpurun_out/bench_cfgE_n$N.err; echo "bench E n$N rc=$?"
fi
tail -3 gpurun_out/bench_cfgE_n$N.err; cut -c1-1200 gpurun_out/r02_bench_cfgE_n$N.json
