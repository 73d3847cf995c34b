#!/bin/bash
# N GPUs of one box (torchrun, one rank per GPU): config E (128^3, batch 1 per GPU) sampling + short trajectory, then config D sampling +
# training step (one CUDA graph incl. the bucketed NCCL all-reduces, replicas_identical)
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29517 bench.py --gpus $N --config E --steps 5 --warmup 3 --trajectory-steps 10 --train-steps 0 --fp32-steps 0 --ddim-steps 0 --no-cpu-baseline \
  > gpurun_out/r02_bench_cfgE_n$N.json 2> gpurun_out/bench_cfgE_n$N.err; echo "bench E n$N rc=$?"
tail -2 gpurun_out/bench_cfgE_n$N.err; cut -c1-700 gpurun_out/r02_bench_cfgE_n$N.json
timeout 400 $TR --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --trajectory-steps 0 --train-steps 10 --train-warmup 3 --fp32-steps 0 --ddim-steps 0 --no-cpu-baseline \
  > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench D n$N rc=$?"
tail -2 gpurun_out/bench_n$N.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["train"])
except Exception as e:
    print("no line:", e)
PY
