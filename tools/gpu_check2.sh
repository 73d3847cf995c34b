#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --trajectory-steps 0 --fp32-steps 0 --ddim-steps 0 2>gpurun_out/bench_quick.err > gpurun_out/bench_quick.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("evals/s", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "achieved_executed", "frac_executed", "whole_step_tflops")}, d["roofline"]["others"])
print("train", {k: d["train"][k] for k in ("value", "ms_per_step", "launch")}, d["train"]["roofline"]["frac"])
PY
tail -2 gpurun_out/bench_quick.err
