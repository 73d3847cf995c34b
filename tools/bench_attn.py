"""Isolated timing of the attention kernels (forward and backward) at config D's shapes; CUDA events, L2-cold enough
(the working set per launch exceeds L2 at T = 16384, N >= 2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from rho_diffusion.b200 import ops

dev = torch.device("cuda", 0)
for (N, T, H, ch) in [(2, 16384, 4, 64), (8, 16384, 4, 64), (2, 4096, 4, 128)]:
    torch.manual_seed(0)
    qkv = torch.randn(N, T, 3 * H * ch, device=dev).bfloat16().requires_grad_(True)
    dy = torch.randn(N, T, H * ch, device=dev).bfloat16()
    def fwd():
        with torch.no_grad():
            return ops.attention(qkv, H, 1)
    for _ in range(3):
        fwd()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        fwd()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 4.0 * N * H * T * T * ch
    print(f"fwd N={N} T={T} ch={ch}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
    out = ops.attention(qkv, H, 1)
    for _ in range(2):
        out.backward(dy, retain_graph=True)
    e0.record()
    for _ in range(reps):
        out.backward(dy, retain_graph=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"bwd N={N} T={T} ch={ch}: {ms:.3f} ms  {2.5 * flops / ms / 1e9:.1f} TFLOP/s (2.5x fwd FLOP convention; 3.5x executed)")
