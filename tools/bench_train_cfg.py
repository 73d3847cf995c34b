"""Training-step time (bf16, AdamW + EMA) and per-kernel breakdown of BASELINE configs A (3-D 32^3, batch 16), B (2-D 128^2 deep-galaxy,
batch 64) or C (1-D L = 1024, batch 512):  bench_train_cfg.py [A|B|C] [graph]   (RHO_TRACE_SIMT=1 lists convs off the tensor-core path)."""
import os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from oracle.synth import synth_input, synth_state_dict
from rho_diffusion.b200 import optimize
from rho_diffusion.diffusion import DDPM, LinearSchedule

dev = torch.device("cuda", 0)
space = OrderedDict([("s", [0.25, 0.5, 1.0]), ("m", [1.0]), ("t", [float(v) for v in range(400, 505, 5)]), ("c", [0.0])])
cfg = next((a for a in sys.argv[1:] if a in ("A", "B", "C")), "B")
common = dict(in_channels=1, out_channels=1, num_res_blocks=2, attention_resolutions=[16, 8], use_scale_shift_norm=True, num_heads=4,
              activation="SiLU", use_new_attention_order=False)
if cfg == "A":
    kw, shape, B, sp = dict(dims=3, model_channels=64, data_shape=[32, 32, 32], num_classes=20, **common), (1, 32, 32, 32), 16, None
elif cfg == "B":
    kw, shape, B, sp = dict(dims=2, model_channels=32, data_shape=[128, 128], num_classes=25, **common), (1, 128, 128), 64, space
else:
    kw, shape, B, sp = dict(dims=1, model_channels=64, data_shape=[1024], num_classes=20, **common), (1, 1024), 512, None
extra = dict(cond_fn="MultiEmbeddings", cond_fn_kwargs={"parameter_space": sp, "embedding_dim": 4 * kw["model_channels"]}) if sp else {}
ddpm = DDPM("UNetv2", dict(kw), LinearSchedule(500, 1e-3, 0.02), "MSELoss", timesteps=500, optimizer="AdamW", opt_kwargs={"lr": 1e-4}, **extra)
manifest = {k: tuple(v.shape) for k, v in ddpm.backbone.state_dict().items()}
ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
ddpm = ddpm.to(dev)
opt = ddpm.configure_optimizers(1)["optimizer"]
ddpm, fused = optimize(ddpm, optimizer=opt, dtype=torch.bfloat16)
ddpm.train()
x = (synth_input("b.x", (B, *shape), kind="rand") * 2 - 1).to(dev)
if sp:
    y = torch.tensor([[0.5, 1.0, 400.0 + 5 * (i % 21), 0.0] for i in range(B)], device=dev)
else:
    y = synth_input("cfg.y", (B, 4 * kw["model_channels"]), kind="rand").to(dev)


def step():
    fused.zero_grad()
    loss = ddpm.training_step([x, y])
    loss.backward()
    fused.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"config {cfg} train: batch {B}  {ms:.2f} ms/step  {B / ms * 1e3:.1f} samples/s")
if "graph" in sys.argv[1:]:
    from rho_diffusion.b200 import GraphedTrainStep
    gstep = GraphedTrainStep(ddpm, fused, x, y)
    for _ in range(3):
        gstep(x, y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        gstep(x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"config {cfg} train (whole-step CUDA graph): batch {B}  {ms:.2f} ms/step  {B / ms * 1e3:.1f} samples/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type.name == "CUDA":
        r = rows.setdefault(ev.name[:90], [0, 0.0]); r[0] += 1; r[1] += ev.device_time / 1e3
tot = sum(r[1] for r in rows.values())
print(f"device time {tot:.2f} ms")
for k, r in sorted(rows.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {r[1]:7.3f} ms {100 * r[1] / tot:5.1f}% x{r[0]:4d}  {k}")
