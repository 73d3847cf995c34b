"""One eager UNet eval of config D between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
ddpm, _ = bench.build_model(bench.CONFIGS['D'], dev, torch.bfloat16)
ddpm.eval()
x, y = bench.make_inputs(bench.CONFIGS['D'], B, 0)
x, y = x.to(dev), y.to(dev)
tt = torch.full((B,), 500, dtype=torch.int64, device=dev)
with torch.inference_mode():
    for _ in range(2):
        ddpm.backbone(x, tt, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ddpm.backbone(x, tt, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
