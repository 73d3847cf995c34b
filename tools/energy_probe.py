"""Where do the Joules of a config-D sampling step go?  Every batch-8 run sits on the 1000 W power cap (`sw_power_cap`), so the step
time is set by ENERGY per eval, not by the sum of stand-alone kernel times.  This probe loops one kernel class at a time for ~1.5 s
(CUDA-graph of 20 launches, replayed), samples board power and SM clocks through NVML every 20 ms, and reports time, power, clocks and
energy per launch next to the algorithmic TFLOP/s resp. GB/s -- then the same for the whole captured step.

    python tools/energy_probe.py [batch]          -> table on stdout (copy under profiles/)
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch  # noqa: E402
import pynvml  # noqa: E402
from rho_diffusion.b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.on, self.rows = True, []

    def run(self):
        while self.on:
            self.rows.append((time.perf_counter(), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                              pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.02)


def probe(name, fn, flops=0.0, nbytes=0.0, seconds=1.5, per_graph=20):
    with torch.inference_mode():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(per_graph):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        reps = max(int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)), 1)
        s = Sampler(); s.start()
        t_begin = time.perf_counter()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        t_end = time.perf_counter()
        s.on = False; s.join()
    ms = e0.elapsed_time(e1) / (reps * per_graph)
    # skip the first 40 % of the window: NVML power is a ~100 ms moving average
    rows = [r for r in s.rows if t_begin + 0.4 * (t_end - t_begin) <= r[0] <= t_end]
    watts = sum(r[1] for r in rows) / max(len(rows), 1)
    mhz = sum(r[2] for r in rows) / max(len(rows), 1)
    print(f"{name:44s} {ms * 1e3:9.1f} us  {watts:6.0f} W  {mhz:5.0f} MHz  {watts * ms:8.2f} mJ"
          + (f"  {flops / ms / 1e9:7.0f} TFLOP/s  {watts * ms * 1e-3 / flops * 1e12:5.2f} pJ/FLOP" if flops else "")
          + (f"  {nbytes / ms / 1e6:7.0f} GB/s  {watts * ms * 1e-3 / nbytes * 1e12 / 8:5.1f} pJ/bit" if nbytes else ""), flush=True)
    return ms, watts


def conv_case(cin, cout, sp, k=3, cin2=0):
    spec = ops.ConvSpec.make(3, k)
    x = torch.randn(B, *sp, cin, device=dev).bfloat16()
    x2 = torch.randn(B, *sp, cin2, device=dev).bfloat16() if cin2 else None
    w = (torch.randn(cout, cin + cin2, k, k, k, device=dev) / (27 * (cin + cin2)) ** 0.5)
    b = torch.zeros(cout, device=dev)
    owner = torch.nn.Module()
    m = B * sp[0] * sp[1] * sp[2]
    return (lambda: ops.conv(owner, x, x2, w, b, None, spec, torch.bfloat16)), 2.0 * m * cout * (cin + cin2) * k ** 3


print(f"config D shapes, batch {B}; idle power {pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0:.0f} W, "
      f"power limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0:.0f} W")
print(f"{'kernel class':44s} {'time':>12s} {'power':>8s} {'clock':>9s} {'energy':>11s}")
for cin, cout, sp, k, cin2 in [(64, 64, (64, 64, 64), 3, 0), (128, 64, (64, 64, 64), 3, 64), (128, 128, (64, 32, 32), 3, 0),
                               (256, 256, (64, 16, 16), 3, 0), (512, 512, (64, 8, 8), 3, 0), (128, 64, (64, 64, 64), 1, 0)]:
    fn, fl = conv_case(cin, cout, sp, k, cin2)
    probe(f"conv k{k} {cin + cin2}->{cout} at {sp[0]}x{sp[1]}x{sp[2]}", fn, flops=fl)
T, H, ch = 16384, 4, 64
qkv = torch.randn(B, T, 3 * H * ch, device=dev).bfloat16()
probe("attention fwd T=16384 ch=64 heads=4", lambda: ops.attention(qkv, H, 1), flops=4.0 * B * H * T * T * ch)
xg = torch.randn(B, 64, 64, 64, 64, device=dev).bfloat16()
gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
film = torch.randn(B, 128, device=dev)
probe("GroupNorm stats + apply (FiLM, SiLU) 64^3 x 64", lambda: ops.gn_forward(xg, None, gamma, beta, film, None, 1, torch.bfloat16),
      nbytes=3.0 * xg.numel() * 2)
a, b2 = torch.empty(1 << 29, device=dev, dtype=torch.bfloat16), torch.empty(1 << 29, device=dev, dtype=torch.bfloat16)
probe("device copy 1 GiB (torch copy_)", lambda: b2.copy_(a), nbytes=2.0 * a.numel() * 2, per_graph=4)
del a, b2
mm_a, mm_b = torch.randn(8192, 8192, device=dev).bfloat16(), torch.randn(8192, 8192, device=dev).bfloat16()
probe("cuBLAS bf16 GEMM 8192^3 (the roofline's peak)", lambda: torch.matmul(mm_a, mm_b), flops=2.0 * 8192 ** 3, per_graph=4)
del mm_a, mm_b

import bench  # noqa: E402
ddpm, _ = bench.build_model(bench.CONFIGS["D"], dev, torch.bfloat16)
ddpm.eval()
x, y = bench.make_inputs(bench.CONFIGS["D"], B, 0)
x, y = x.to(dev), y.to(dev)
tt = torch.full((B,), 500, dtype=torch.int64, device=dev)
probe("whole UNet eval (config D)", lambda: ddpm.backbone(x, tt, y), flops=5506.35e9 * B, seconds=3.0, per_graph=1)
