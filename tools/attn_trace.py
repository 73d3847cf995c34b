"""Hand-off timeline of one CTA of the head-dim-64 tcgen05 attention forward (rho_debug_attention_trace): per key tile, when
the row owners of the two query tiles saw S / published P and when the MMA issuer saw P / finished issuing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch
from rho_diffusion.b200 import _lib, ops
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
qkv = torch.randn(N, 16384, 3 * 4 * 64, device=dev).bfloat16()
buf = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
with torch.no_grad():
    ops.attention(qkv, 4, 1)
    _lib.lib().rho_debug_attention_trace(buf.data_ptr())
    ops.attention(qkv, 4, 1)
    torch.cuda.synchronize()
    _lib.lib().rho_debug_attention_trace(None)
t = buf.cpu().view(64, 16)
t0 = int(t[8, 0])
names = ["S0 seen", "P0 pub", "S1 seen", "P1 pub", "mma P0 seen", "mma 0 issued", "mma P1 seen", "mma 1 issued"]
print("tile  " + "  ".join(f"{n:>12s}" for n in names))
for j in range(8, 24):
    print(f"{j:4d}  " + "  ".join(f"{int(t[j, k]) - t0:12d}" for k in range(8)))
per = (t[40, 0] - t[8, 0]).item() / 32
print(f"period per key tile: {per:.0f} clk")
for a, b, what in ((0, 1, "softmax 0 active"), (2, 3, "softmax 1 active")):
    print(what, float((t[8:40, b] - t[8:40, a]).float().mean()))
print("WG0: P pub -> next S seen", float((t[9:41, 0] - t[8:40, 1]).float().mean()))
print("WG1: P pub -> next S seen", float((t[9:41, 2] - t[8:40, 3]).float().mean()))
print("P0 pub -> mma saw", float((t[8:40, 4] - t[8:40, 1]).float().mean()), " issue time", float((t[8:40, 5] - t[8:40, 4]).float().mean()))
print("P1 pub -> mma saw", float((t[8:40, 6] - t[8:40, 3]).float().mean()), " issue time", float((t[8:40, 7] - t[8:40, 6]).float().mean()))

if int(t[8, 8]) != 0:      # two-threads-per-row kernel: inner stamps of warp (i, c = 0, quad = 0)
    for i in (0, 1):
        b = 8 + 4 * i
        print(f"WG{i} inner: S seen -> ld done {float((t[8:40, b] - t[8:40, 2 * i]).float().mean()):.0f}, exp32 #1 {float((t[8:40, b + 1] - t[8:40, b]).float().mean()):.0f}, "
              f"o_done wait {float((t[8:40, b + 2] - t[8:40, b + 1]).float().mean()):.0f}, exp32 #2 + st wait {float((t[8:40, b + 3] - t[8:40, b + 2]).float().mean()):.0f}, "
              f"vote + arrive {float((t[8:40, 2 * i + 1] - t[8:40, b + 3]).float().mean()):.0f}")
