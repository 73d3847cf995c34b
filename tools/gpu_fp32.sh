#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fp32 or float32 or taps or reverse or golden or oracle" 2>&1 | tail -8
python - <<'PY'
import os, sys, time, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "rho-diffusion_b200")]
import bench
from rho_diffusion.b200 import ops, set_compute_dtype
cfg = bench.CONFIGS["D"]
dev = torch.device("cuda", 0)
ddpm, _ = bench.build_model(cfg, dev, torch.float32)
ddpm.eval()
x, y = bench.make_inputs(cfg, 2, 0)
x, y = x.to(dev), y.to(dev)
t = torch.full((2,), 500, dtype=torch.int64, device=dev)
outs = {}
for split in (True, False):
    ops._fp32_split = split
    with torch.inference_mode():
        o = ddpm.backbone(x, t, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            o = ddpm.backbone(x, t, y)
        e1.record(); torch.cuda.synchronize()
    outs[split] = o.double()
    print("fp32 eval, split engine" if split else "fp32 eval, CUDA-core engine", f"{e0.elapsed_time(e1) / 3:.1f} ms per batch-2 eval")
d = (outs[True] - outs[False])
print("split vs CUDA-core engine: rel L2", float(d.norm() / outs[False].norm()), "max abs", float(d.abs().max()))
PY
