"""Per-kernel device-time table of one training step / one sampling eval of config D (torch.profiler / CUPTI)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
import bench
from rho_diffusion.b200 import optimize

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
CFG = bench.CONFIGS[os.environ.get('RHO_PROF_CONFIG', 'D')]
ddpm, _ = bench.build_model(CFG, dev, torch.bfloat16)
x, y = bench.make_inputs(CFG, B, 0)
x, y = x.to(dev), y.to(dev)
if mode == "train":
    opt = ddpm.configure_optimizers(1)["optimizer"]
    ddpm, fused = optimize(ddpm, optimizer=opt, dtype=torch.bfloat16)
    ddpm.train()

    def step():
        fused.zero_grad()
        loss = ddpm.training_step([x, y])
        loss.backward()
        fused.step()
else:
    ddpm.eval()
    tt = torch.full((B,), 500, dtype=torch.int64, device=dev)

    def step():
        with torch.inference_mode():
            ddpm.backbone(x, tt, y)

for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(e.name[:70], [0, 0.0])
        r[0] += 1
        r[1] += e.device_time
tot = sum(r[1] for r in rows.values())
print(f"mode={mode} B={B} total device time {tot/1e3:.2f} ms")
for k, r in sorted(rows.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{r[1]/1e3:9.3f} ms {100*r[1]/tot:5.1f}% x{r[0]:4d}  {k}")

print("-- individual launches of the CUDA-core conv engines (us) --")
for e in prof.events():
    pats = os.environ.get("RHO_PROF_LIST", "simt,small").split(",")
    if e.device_type == torch.autograd.DeviceType.CUDA and any(q in e.name for q in pats):
        print(f"{e.device_time:10.1f}  {e.name[:90]}")
