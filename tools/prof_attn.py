"""Per-kernel device times of one attention forward + backward at config D's shape (torch profiler)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
from rho_diffusion.b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
qkv = torch.randn(N, 16384, 3 * 4 * 64, device=dev).bfloat16().requires_grad_(True)
dy = torch.randn(N, 16384, 4 * 64, device=dev).bfloat16()
for _ in range(2):
    out = ops.attention(qkv, 4, 1); out.backward(dy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        out = ops.attention(qkv, 4, 1); out.backward(dy)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"{e.device_time_total / e.count / 1000:8.3f} ms x{e.count:3d}  {e.key[:110]}")
