"""One conv layer (forward, bf16, channels-last) timed with CUDA events, for `ncu -k regex:conv` captures:
    one_conv.py dims N Cin Cout k spatial...     e.g.  one_conv.py 2 256 32 32 3 128 128      (RHO_FUSE_GN_STATS=0: no statistics)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from rho_diffusion.b200 import ops

dims, N, Cin, Cout, k = (int(v) for v in sys.argv[1:6])
sp = tuple(int(v) for v in sys.argv[6:6 + dims])
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(1)
x = torch.randn((N, *sp, Cin), generator=g).to(dev).to(torch.bfloat16)        # channels-last storage
w = (torch.randn((Cout, Cin, *([k] * dims)), generator=g) / (Cin * k ** dims) ** 0.5).to(dev)
b = torch.zeros(Cout, device=dev)
holder = torch.nn.Module()
spec = ops.ConvSpec.make(dims, k, 1, 1)
with torch.inference_mode():
    for _ in range(3):
        y = ops.conv(holder, x, None, w, b, None, spec, torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = ops.conv(holder, x, None, w, b, None, spec, torch.bfloat16)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
M = N
for s in sp:
    M *= s
fl = 2.0 * M * Cin * Cout * k ** dims
by = 2.0 * M * (Cin + Cout)
print(f"conv{dims}d {Cin}->{Cout} k{k} N={N} {sp}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.1f} GB/s (algorithmic)")
