"""Error of the split-bf16 fp32 conv engine vs an fp64 reference as a function of the contraction length (is it the tensor core's
fp32 accumulation?): 3-D k = 3 convs with Cin = 32 ... 1024 on a 16^3 volume; the CUDA-core engine beside it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
import torch.nn.functional as F
from rho_diffusion.b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for cin in (32, 128, 512, 1024):
    x = torch.randn(1, cin, 8, 16, 16)
    w = torch.randn(64, cin, 3, 3, 3) / (27 * cin) ** 0.5
    ref = F.conv3d(x.double(), w.double(), padding=1)
    xc = x.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    holder = torch.nn.Module()
    spec = ops.ConvSpec.make(3, 3, 1, 1)
    res = {}
    with torch.no_grad():
        for split in (True, False):
            ops._fp32_split = split
            y = ops.conv(holder, xc, None, w.to(dev), None, None, spec, torch.float32)
            res[split] = y.permute(0, 4, 1, 2, 3).double().cpu()
    e = {k: float((v - ref).norm() / ref.norm()) for k, v in res.items()}
    print(f"Cin {cin:5d} (K = {27 * cin:6d}): split engine rel L2 {e[True]:.2e}   CUDA-core engine {e[False]:.2e}   mean signed err split {float((res[True] - ref).mean() / ref.abs().mean()):+.2e}")
