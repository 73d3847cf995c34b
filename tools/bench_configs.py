"""Sampling-step throughput (UNet denoise evals/s, CUDA-graph ancestral step, bf16) of the other BASELINE.json configs:
A (3-D 32^3, the shipped spherical-harmonics config), B (2-D 128^2 deep-galaxy, MultiEmbeddings), C (1-D L = 1024)."""
import os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from oracle.synth import synth_input, synth_state_dict
from rho_diffusion.b200 import set_compute_dtype
from rho_diffusion.b200.sampler import AncestralSampler
from rho_diffusion.diffusion import DDPM, LinearSchedule

dev = torch.device("cuda", 0)
space = OrderedDict([("s", [0.25, 0.5, 1.0]), ("m", [1.0]), ("t", [float(v) for v in range(400, 505, 5)]), ("c", [0.0])])
common = dict(in_channels=1, out_channels=1, num_res_blocks=2, attention_resolutions=[16, 8], use_scale_shift_norm=True, num_heads=4,
              activation="SiLU", use_new_attention_order=False)
CONFIGS = {
    "A 3-D 32^3 mc64": (dict(dims=3, model_channels=64, data_shape=[32, 32, 32], num_classes=20, **common), (1, 32, 32, 32), [32, 100], None),
    "B 2-D 128^2 mc32": (dict(dims=2, model_channels=32, data_shape=[128, 128], num_classes=25, **common), (1, 128, 128), [64, 256], space),
    "C 1-D 1024 mc64": (dict(dims=1, model_channels=64, data_shape=[1024], num_classes=20, **common), (1, 1024), [64, 512], None),
    "E 3-D 128^3 mc64": (dict(dims=3, model_channels=64, data_shape=[128, 128, 128], num_classes=20, **common), (1, 128, 128, 128), [1, 2], None),
}
for name, (kw, shape, batches, sp) in CONFIGS.items():
    if len(sys.argv) > 2 and not name.startswith(tuple(sys.argv[2:])):      # e.g. `bench_configs.py run E`
        continue
    T = 1000
    extra = dict(cond_fn="MultiEmbeddings", cond_fn_kwargs={"parameter_space": sp, "embedding_dim": 4 * kw["model_channels"]}) if sp else {}
    ddpm = DDPM("UNetv2", dict(kw), LinearSchedule(T, 1e-3, 0.02), "MSELoss", timesteps=T, optimizer="AdamW", opt_kwargs={"lr": 1e-4}, **extra)
    manifest = {k: tuple(v.shape) for k, v in ddpm.backbone.state_dict().items()}
    ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
    ddpm = set_compute_dtype(ddpm.to(dev).eval(), torch.bfloat16)
    alpha, beta, abar = ddpm._tables(dev)
    for B in batches:
        if sp:
            cond = torch.tensor([[0.5, 1.0, 400.0 + 5 * (i % 21), 0.0] for i in range(B)], device=dev)
        else:
            cond = synth_input("cfg.y", (B, 4 * kw["model_channels"]), kind="rand").to(dev)
        with torch.inference_mode():
            s = AncestralSampler(ddpm.backbone, (B, *shape), dev, alpha, beta, abar, cond_example=cond, use_cuda_graph=True)
            s._capture(seed=1)
            s.x.copy_(synth_input("cfg.x", (B, *shape)).to(dev))
            s.tt.fill_(T - 1); s.t_scalar.fill_(T - 1)
            for _ in range(3):
                s.graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            K = 20 if B * shape[-1] ** len(shape[1:]) < 1 << 21 else 4
            for _ in range(K):
                s.graph.replay()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
        print(f"{name}: batch {B:4d}  {ms:8.3f} ms/step  {B / ms * 1e3:9.1f} evals/s", flush=True)
        del s
    del ddpm

if len(sys.argv) > 1 and sys.argv[1] == "table":
    # per-kernel-shape table (CUDA events around every conv / GN / attention launch of one eager eval)
    from rho_diffusion.b200 import ops
    for name, (kw, shape, batches, sp) in CONFIGS.items():
        if not name.startswith(tuple(sys.argv[2:] or ["B", "C"])):
            continue
        T = 1000
        extra = dict(cond_fn="MultiEmbeddings", cond_fn_kwargs={"parameter_space": sp, "embedding_dim": 4 * kw["model_channels"]}) if sp else {}
        ddpm = DDPM("UNetv2", dict(kw), LinearSchedule(T, 1e-3, 0.02), "MSELoss", timesteps=T, optimizer="AdamW", opt_kwargs={"lr": 1e-4}, **extra)
        manifest = {k: tuple(v.shape) for k, v in ddpm.backbone.state_dict().items()}
        ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
        ddpm = set_compute_dtype(ddpm.to(dev).eval(), torch.bfloat16)
        B = batches[-1]
        cond = torch.tensor([[0.5, 1.0, 400.0 + 5 * (i % 21), 0.0] for i in range(B)], device=dev) if sp else \
            synth_input("cfg.y", (B, 4 * kw["model_channels"]), kind="rand").to(dev)
        x = synth_input("cfg.x", (B, *shape)).to(dev)
        tt = torch.full((B,), 500, dtype=torch.int64, device=dev)
        with torch.inference_mode():
            for _ in range(2):
                ddpm.backbone(x, tt, cond)
            torch.cuda.synchronize()
            ops.profiler = ops.KernelProfiler()
            ddpm.backbone(x, tt, cond)
            ops.profiler.summary()
            rows = {}
            for kind, nm, flops, nbytes, s0, s1 in ops.profiler.records:
                r = rows.setdefault((kind, nm), [0, 0.0, 0.0, 0.0])
                r[0] += 1; r[1] += s0.elapsed_time(s1); r[2] += flops; r[3] += nbytes
            ops.profiler = None
        tot = sum(r[1] for r in rows.values())
        print(f"== {name} batch {B}: {tot:.3f} ms in profiled kernels")
        for (kind, nm), r in sorted(rows.items(), key=lambda kv: -kv[1][1])[:18]:
            print(f"  {r[1]:7.3f} ms x{r[0]:3d}  {r[2] / (r[1] * 1e-3) / 1e12 if r[1] else 0:8.1f} TFLOP/s {r[3] / (r[1] * 1e-3) / 1e9 if r[1] else 0:8.1f} GB/s  {nm}")
