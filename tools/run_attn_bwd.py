"""One attention forward + backward at config D's shape (for ncu captures of the backward kernels): python tools/run_attn_bwd.py [N]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from rho_diffusion.b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.manual_seed(0)
qkv = torch.randn(N, 16384, 3 * 4 * 64, device=dev).bfloat16().requires_grad_(True)
dy = torch.randn(N, 16384, 4 * 64, device=dev).bfloat16()
for _ in range(2):
    out = ops.attention(qkv, 4, 1)
    out.backward(dy)
torch.cuda.synchronize()
print("ok", float(qkv.grad.float().abs().mean()))
