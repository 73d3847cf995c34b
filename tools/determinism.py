"""Run-to-run determinism of every op of one config-D eval: checksums of each conv / GroupNorm / attention output of two
identical evals, first divergence reported."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
import bench
from rho_diffusion.b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
ddpm, _ = bench.build_model(bench.CONFIGS['D'], dev, torch.bfloat16)
ddpm.eval()
x, y = bench.make_inputs(bench.CONFIGS['D'], B, 0)
x, y = x.to(dev), y.to(dev)
tt = torch.full((B,), 437, dtype=torch.int64, device=dev)
log = []

def checksum(t):
    v = t.detach().contiguous().view(torch.int16 if t.element_size() == 2 else torch.int32).to(torch.int64)
    return int(v.sum()), int((v * (torch.arange(v.numel(), device=v.device) % 8191).view(v.shape)).sum())

orig_conv, orig_gn, orig_attn = ops._conv_forward, ops.gn_forward, ops.AttentionFn.apply

def conv(x, x2, wp, bias, residual, spec, Cout, *a, **k):
    out = orig_conv(x, x2, wp, bias, residual, spec, Cout, *a, **k)
    log.append(("conv %s+%s->%d %s k%s s%s u%s" % (x.shape[-1], 0 if x2 is None else x2.shape[-1], Cout, tuple(x.shape[1:-1]), spec.kernel, spec.stride, spec.up), checksum(out)))
    return out

def gn(*a, **k):
    out = orig_gn(*a, **k)
    log.append(("gn %s" % (tuple(out[0].shape),), checksum(out[0]) + checksum(out[1])))
    return out

def attn(qkv, heads, order):
    out = orig_attn(qkv, heads, order)
    log.append(("attn %s" % (tuple(qkv.shape),), checksum(out)))
    return out

ops._conv_forward, ops.gn_forward = conv, gn
ops.AttentionFn.apply = staticmethod(attn)
runs = []
with torch.inference_mode():
    ddpm.backbone(x, tt, y)
    for r in range(3):
        log.clear()
        out = ddpm.backbone(x, tt, y)
        torch.cuda.synchronize()
        runs.append((list(log), checksum(out)))
print("ops per eval:", len(runs[0][0]))
for r in (1, 2):
    diffs = [(i, a[0]) for i, (a, b) in enumerate(zip(runs[0][0], runs[r][0])) if a != b]
    print(f"run {r} vs 0: final equal={runs[0][1] == runs[r][1]}  differing ops: {len(diffs)}")
    for i, name in diffs[:12]:
        print("   ", i, name)
