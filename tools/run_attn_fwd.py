"""Forward attention calls only at config D's shape (for ncu captures): python tools/run_attn_fwd.py [N] [reps]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from rho_diffusion.b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
qkv = torch.randn(N, 16384, 3 * 4 * 64, device=dev).bfloat16()
with torch.no_grad():
    for _ in range(reps):
        out = ops.attention(qkv, 4, 1)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
