"""Run-to-run determinism stress at full size (config D): N repeated evals must be bit-identical, and two identical
training runs (K steps, same data / t / noise) must end with bit-identical weights."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
import bench
from rho_diffusion.b200 import optimize

dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for B in (2, 8):
    ddpm, _ = bench.build_model(bench.CONFIGS['D'], dev, torch.bfloat16)
    ddpm.eval()
    x, y = bench.make_inputs(bench.CONFIGS['D'], B, 0)
    x, y = x.to(dev), y.to(dev)
    tt = torch.full((B,), 437, dtype=torch.int64, device=dev)
    with torch.inference_mode():
        ref = ddpm.backbone(x, tt, y).clone()
        bad = sum(int(not torch.equal(ddpm.backbone(x, tt, y), ref)) for _ in range(reps))
    print(f"eval B={B}: {reps} repeats, {bad} differ", flush=True)
    del ddpm


def train_run():
    ddpm, _ = bench.build_model(bench.CONFIGS['D'], dev, torch.bfloat16)
    opt = ddpm.configure_optimizers(1)["optimizer"]
    ddpm, fused = optimize(ddpm, optimizer=opt, dtype=torch.bfloat16)
    ddpm.train()
    x, y = bench.make_inputs(bench.CONFIGS['D'], 2, 100)
    x, y = x.to(dev), y.to(dev)
    g = torch.Generator().manual_seed(1)
    losses = []
    for _ in range(3):
        t = torch.randint(0, 1000, (2,), generator=g)
        noise = torch.randn(x.shape, generator=g).to(dev)
        fused.zero_grad()
        loss = ddpm.training_step([x, y], t=t, noise=noise)
        loss.backward()
        fused.step()
        losses.append(float(loss.detach()))
    return losses, fused.flat_p.clone()


l1, p1 = train_run()
l2, p2 = train_run()
print("train losses", l1, l2, "weights equal:", bool(torch.equal(p1, p2)), flush=True)
