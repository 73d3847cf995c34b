#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native rho-diffusion hot path.

Metric (BASELINE.json): UNet denoise evals/s (batch x steps / s) of the 3-D 64^3 conditional UNet (attention at
64x16x16, "config D" of SURVEY.md section 8) inside the DDPM ancestral sampling loop, plus train samples/s of the same
model.  One "step" = one sampling step (UNet eval + fused ancestral update) over one batch of synthetic density fields.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config D|E] [--batch B] [--dtype bf16|fp32]

Legs of one line (rank 0 prints ONE JSON line):
  value        K replays of the captured sampling step, inputs resident in HBM (CUDA events, max over ranks)
  e2e          the same step with pinned-host -> device copies of x / labels and a device -> host read of x EVERY step
  trajectory   a FULL T-step trajectory through the public ``DDPM.generate()`` (host result), seconds per trajectory
  ddim         the GaussianDiffusionPipeline DDIM step (x0 prediction + dynamic thresholding) of scripts/inference.py
  fp32         evals/s of the exact fp32 engine (the reference's arithmetic: TF32 off)
  train        optimizer steps/s of DDPM.training_step + backward + fused AdamW (whole step as ONE CUDA graph at every N;
               the data-parallel gradient all-reduce is bucketed, overlapped with the backward and captured in the graph)
  roofline     conv implicit-GEMM class: algorithmic AND executed FLOP / CUDA-event time vs the measured sustained bf16 peak
  cpu_baseline the UNMODIFIED reference (oracle/_ref, kind "reference") or its restatement (kind "port") on the host cores

For N > 1 launch with torchrun (one rank per GPU); sampling shards the batch across ranks with no communication (weak
scaling: per-GPU batch fixed).  ``--impl reference`` times the reference's own CPU implementation of the same workload
(``DDPM.reverse_process`` of the unmodified reference over a K-step schedule) on the host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "rho-diffusion_b200")

import torch  # noqa: E402

# ---- workloads ---------------------------------------------------------------------
CONFIGS = {
    "D": dict(kwargs=dict(dims=3, in_channels=1, out_channels=1, model_channels=64, num_res_blocks=2, data_shape=[64, 64, 64],
                          attention_resolutions=[4], use_scale_shift_norm=True, num_heads=4, num_classes=20, activation="SiLU",
                          use_new_attention_order=True),
              batch=8, train_batch=4, train_tflop_per_sample=17.9,
              workload="3D conditional UNet 64^3 (mc64, nrb2, FiLM, 4 heads, attention at 64x16x16), DDPM ancestral sampling step, T=1000"),
    "E": dict(kwargs=dict(dims=3, in_channels=1, out_channels=1, model_channels=64, num_res_blocks=2, data_shape=[128, 128, 128],
                          attention_resolutions=[16, 8], use_scale_shift_norm=True, num_heads=4, num_classes=20, activation="SiLU",
                          use_new_attention_order=True),
              batch=1, train_batch=0, train_tflop_per_sample=None,
              workload="3D conditional UNet 128^3 (mc64, nrb2, FiLM, 4 heads, attention at 128x16x16 and 128x8x8), DDPM ancestral sampling step, T=1000"),
}
T_STEPS = 1000


def add_paths(product: bool):
    for p in ((ROOT, PKG) if product else (ROOT,)):
        if p not in sys.path:
            sys.path.insert(0, p)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []      # lines: (host time of receipt, csv row)
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # samples received inside the timed region (the sampler is started before the warm-up so that it is already
        # streaming); a region shorter than the sampling period falls back to the samples nearest to it
        lo = self.t_begin if self.t_begin is not None else 0.0
        hi = (self.t_end if self.t_end is not None else time.time()) + 0.05
        rows = [ln for (ts, ln) in self.lines if lo <= ts <= hi]
        if not rows and self.lines:
            mid = 0.5 * (lo + hi)
            rows = [ln for (_, ln) in sorted(self.lines, key=lambda r: abs(r[0] - mid))[:2]]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(batch):
    """DRAM bytes of the tcgen05 conv kernels of ONE eval (sum over their launches) from the committed `ncu` pass of
    tools/one_eval.py at the same batch (profiles/r02_ncu_traffic_b<batch>.json, else the round-1 file), or None."""
    for name in (f"r02_ncu_traffic_b{batch}.json", f"r01_ncu_traffic_b{batch}.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as fh:
                d = json.load(fh)
            out = d.get("conv_tcgen05")
            if out:
                out = dict(out, file="profiles/" + name)
            return out
    return None


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return world, rank, local


def make_inputs(cfg, batch, seed):
    from oracle.synth import synth_input
    x = synth_input(f"bench.x.{seed}", (batch, 1, *cfg["kwargs"]["data_shape"]))
    y = synth_input(f"bench.y.{seed}", (batch, 4 * cfg["kwargs"]["model_channels"]), kind="rand")
    return x, y


def build_model(cfg, device, dtype, seed=777):
    from oracle.synth import synth_state_dict
    from rho_diffusion.b200 import set_compute_dtype
    from rho_diffusion.diffusion import DDPM, LinearSchedule
    ddpm = DDPM("UNetv2", dict(cfg["kwargs"]), LinearSchedule(T_STEPS, 1e-3, 0.02), "MSELoss", timesteps=T_STEPS, optimizer="AdamW",
                opt_kwargs={"lr": 1e-4}, seed=seed)
    manifest = {k: tuple(v.shape) for k, v in ddpm.backbone.state_dict().items()}
    ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=777))   # random-init weights of the architecture
    return set_compute_dtype(ddpm.to(device), dtype), manifest


# --------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's own implementation of the path on the host cores -- ``DDPM.reverse_process`` of the UNMODIFIED
    reference (oracle/_ref, imported through oracle/ref_loader.py) over a K-step LinearSchedule, i.e. K denoise evals + K
    ancestral updates of the same UNet (same architecture, same synthetic weights and inputs as the B200 arm) at batch 1.
    Falls back to the restatement (oracle/torch_oracle.py, kind "port") only if the installed reference is missing."""
    world, rank, _ = dist_env()
    if rank != 0:
        return
    add_paths(product=False)
    cfg = CONFIGS[args.config]
    from oracle import ref_loader
    from oracle.synth import synth_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    batch = 1
    x, y = make_inputs(cfg, batch, 0)
    budget_s = float(os.environ.get("RHO_BENCH_CPU_BUDGET_S", "150"))
    kind = "reference" if ref_loader.available() else "port"
    if kind == "reference":
        R = ref_loader.load()
        sched1 = R.schedule.LinearSchedule(num_steps=1, device="cpu")
        pipe = R.DDPM(backbone="UNetv2", backbone_kwargs=dict(cfg["kwargs"]), schedule=sched1, loss_func="MSELoss", timesteps=1,
                      optimizer="AdamW", opt_kwargs={"lr": 1e-4})
        manifest = {k: tuple(v.shape) for k, v in pipe.backbone.state_dict().items()}
        pipe.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
        pipe.eval()
        t0 = time.perf_counter()
        pipe.reverse_process(x, conditions=y)                  # warm-up: one eval through the stock sampling loop
        t_warm = time.perf_counter() - t0
        steps = max(1, min(args.steps, int(budget_s / max(t_warm, 1e-3))))
        pipe.schedule = R.schedule.LinearSchedule(num_steps=steps, device="cpu")
        t0 = time.perf_counter()
        pipe.reverse_process(x, conditions=y)
        elapsed = time.perf_counter() - t0
        timed = steps
        what = f"DDPM.reverse_process of the unmodified reference over a {steps}-step LinearSchedule"
    else:
        add_paths(product=True)
        from oracle import torch_oracle as O
        from rho_diffusion.models.unet_v2 import UNet
        manifest = {k: tuple(v.shape) for k, v in UNet(**cfg["kwargs"]).state_dict().items()}
        sd = synth_state_dict(manifest, seed=777)
        spec = O.UNetSpec(**cfg["kwargs"])
        sched = O.linear_schedule(T_STEPS, 1e-3, 0.02)
        tt = torch.full((batch,), T_STEPS - 1, dtype=torch.long)
        timed, elapsed, t_begin = 0, 0.0, time.perf_counter()
        with torch.inference_mode():
            for i in range(1 + args.steps):
                if timed >= 1 and time.perf_counter() - t_begin > budget_s:
                    break
                t0 = time.perf_counter()
                eps = O.unet_forward(sd, spec, x, tt, y)
                x = O.p_sample_step(sched, x, eps, torch.zeros_like(x), T_STEPS - 1)
                if i >= 1:
                    timed += 1
                    elapsed += time.perf_counter() - t0
        what = "UNet eval + ancestral update of the torch-CPU restatement (oracle/torch_oracle.py)"
    value = batch * timed / elapsed
    sample = f"{timed} timed evals of batch {batch} (of {args.steps} requested; {budget_s:.0f} s CPU budget), fp32, {what}"
    line = {
        "impl": "reference", "metric": "UNet denoise evals/s", "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": timed,
        "warmup": 1, "ms_per_step": 1000.0 * elapsed / timed, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["workload"], "batch_per_step": batch},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_leg(args, budget_s=30.0):
    """Bounded sample of the same workload on the host cores: the reference arm in its own process (the reference package has
    the product's name, so the two cannot share an interpreter)."""
    env = dict(os.environ, RHO_BENCH_CPU_BUDGET_S=str(budget_s), WORLD_SIZE="1", RANK="0", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="")
    for k in ("MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID", "GROUP_RANK", "LOCAL_WORLD_SIZE"):
        env.pop(k, None)
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config, "--steps", "3",
                              "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
        line = json.loads(res.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": "evals/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": f"cpu baseline failed: {exc!r}"}


# --------------------------------------------------------------------------------------------------------
def executed_fraction(name: str) -> float:
    """Executed / algorithmic FLOP of a conv launch: the fused nearest-upsample conv runs 2 of the 3 taps per upsampled axis
    (pre-summed weights per output-parity class), i.e. 12 / 27 in 3-D (depth is not upsampled), 4 / 9 in 2-D, 2 / 3 in 1-D."""
    if "_up" not in name:
        return 1.0
    dims = int(name[4])
    return {3: 12.0 / 27.0, 2: 4.0 / 9.0, 1: 2.0 / 3.0}[dims]


def run_b200(args):
    add_paths(product=True)
    import torch.distributed as dist
    from rho_diffusion.b200 import _lib, ops, optimize
    from rho_diffusion.b200.sampler import AncestralSampler

    cfg = CONFIGS[args.config]
    world, rank, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B = args.batch if args.batch > 0 else cfg["batch"]
    shape = (B, 1, *cfg["kwargs"]["data_shape"])
    pk = peaks()

    ddpm, _ = build_model(cfg, device, dtype, seed=777)      # rank folded in through ddpm.global_rank (process group)
    ddpm.eval()
    x_host, y_host = make_inputs(cfg, B, rank)
    x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()
    x_out_pin = torch.empty_like(x_pin).pin_memory()
    alpha, beta, abar = ddpm._tables(device)
    cond_dev = y_host.to(device)
    sampler = AncestralSampler(ddpm.backbone, shape, device, alpha, beta, abar, cond_example=cond_dev, use_cuda_graph=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1 or v is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- capture (counts the kernels of one step through the C-ABI launch counter)
    with torch.inference_mode():
        c0 = _lib.launch_count()
        sampler._capture(seed=1234 + rank)
        launches_per_step = (_lib.launch_count() - c0) // 2      # warm-up pass + capture pass
        x_dev = x_host.to(device)
        sampler.x.copy_(x_dev)
        sampler.tt.fill_(T_STEPS - 1)
        sampler.t_scalar.fill_(T_STEPS - 1)

        # ---- device-resident leg: K replays of the captured step (inputs already in HBM; the working set of one eval,
        #      >1 GB of activations + 326 MB of bf16 weights, is far larger than the 126 MB L2, so no flush is needed)
        clocks = ClockSampler(local)
        clocks.start()
        for _ in range(args.warmup):
            sampler.graph.replay()
        barrier()
        clocks.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sampler.graph.replay()
        e1.record()
        barrier()
        clocks.mark_end()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop()

        # ---- end-to-end leg: host buffers in, host buffers out, EVERY step.  The PCIe copies run on a second stream through
        #      double-buffered device staging tensors, so step k's upload / step k-1's download overlap step k's kernels; the
        #      compute stream only does two device-to-device copies around the graph replay.
        #      (the staging pipeline is the product's: rho_diffusion.b200.sampler.HostStagedStepper)
        from rho_diffusion.b200.sampler import HostStagedStepper
        stepper = HostStagedStepper(sampler)

        def e2e_step():
            stepper.step(x_pin, y_pin, x_out_pin)

        sampler.tt.fill_(T_STEPS // 2)
        sampler.t_scalar.fill_(T_STEPS // 2)
        for _ in range(min(args.warmup, 3)):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            e2e_step()
        stepper.finish()                                               # the last download is inside the timed region
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        del stepper

        # ---- per-kernel roofline pass: one eager eval with CUDA events around every conv / GN / attention launch
        tt = torch.full((B,), 500, dtype=torch.int64, device=device)
        ddpm.backbone(x_dev, tt, cond_dev)      # untimed: fills this stream's allocator pool (no cudaMalloc in the timed pass)
        torch.cuda.synchronize()
        ops.profiler = ops.KernelProfiler()
        ddpm.backbone(x_dev, tt, cond_dev)
        prof = ops.profiler.summary()
        records = ops.profiler.records
        ops.profiler = None
        conv_exec_flops = 0.0
        rows = {}
        for kind, name, flops, nbytes, s0, s1 in records:
            r = rows.setdefault((kind, name), [0, 0.0, 0.0, 0.0, 0.0])
            ex = flops * (executed_fraction(name) if kind == "conv" else 1.0)
            r[0] += 1; r[1] += s0.elapsed_time(s1); r[2] += flops; r[3] += nbytes; r[4] += ex
            if kind == "conv":
                conv_exec_flops += ex
        if args.launch_table and rank == 0:
            with open(args.launch_table, "w") as fh:
                fh.write("kind,name,launches,total_ms,tflops_algorithmic,tflops_executed,gbs\n")
                for (kind, name), r in sorted(rows.items(), key=lambda kv: -kv[1][1]):
                    sec = r[1] * 1e-3
                    fh.write(f"{kind},{name},{r[0]},{r[1]:.4f},{r[2] / sec / 1e12 if sec else 0:.2f},{r[4] / sec / 1e12 if sec else 0:.2f},"
                             f"{r[3] / sec / 1e9 if sec else 0:.1f}\n")

    # ---- full trajectory through the public API: DDPM.generate() -> p_sample -> reverse_process (T-step CUDA-graph loop),
    #      batch-sharded across ranks, result read back to the host inside the timed region
    traj = None
    if args.trajectory_steps > 0:
        ddpm.sampling_batch_size = B * world
        ddpm.t_checkpoints = None
        labels_all = torch.cat([make_inputs(cfg, B, r)[1] for r in range(world)]).to(device)
        n_steps = args.trajectory_steps
        if n_steps != T_STEPS:
            from rho_diffusion.diffusion import LinearSchedule
            ddpm.schedule = LinearSchedule(n_steps, 1e-3, 0.02)
        barrier()
        w0 = time.perf_counter()
        out = ddpm.generate(parameter_space=labels_all)["denoised"]
        host = out.cpu()                                     # device -> host of the local slice (synchronises)
        barrier()
        traj_s = max_over_ranks(time.perf_counter() - w0)
        traj = {"metric": "seconds per full ancestral trajectory through DDPM.generate() (host result)", "value": traj_s, "unit": "s",
                "steps": n_steps, "batch_per_gpu": B, "global_batch": B * world, "evals_per_s": world * B * n_steps / traj_s,
                "includes": "CUDA-graph capture of the step, x_T noise, every step, device->host copy of the samples",
                "finite": bool(torch.isfinite(host).all()), "mean": float(host.mean()), "std": float(host.std())}
        if n_steps != T_STEPS:
            from rho_diffusion.diffusion import LinearSchedule
            ddpm.schedule = LinearSchedule(T_STEPS, 1e-3, 0.02)
        ddpm._samplers = {}

    # ---- x0-prediction DDIM sampler (GaussianDiffusionPipeline.generate, the loop scripts/inference.py runs): same UNet,
    #      plus the per-sample |x0| quantile and the fused DDIM update, one CUDA graph per step
    ddim = None
    if args.ddim_steps > 0:
        from rho_diffusion.b200.sampler import X0Sampler
        from rho_diffusion.diffusion.gaussian_diffusion import sampler_tables
        with torch.inference_mode():
            tab = sampler_tables(T_STEPS, device)
            xs = X0Sampler(ddpm.backbone, shape, device, tab, 0.9, cond_example=cond_dev, use_cuda_graph=True)
            xs._capture()
            xs.x.copy_(x_dev)
            xs.tt.fill_(T_STEPS - 1)
            xs.t_scalar.fill_(T_STEPS - 1)
            for _ in range(3):
                xs.graph.replay()
            barrier()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(args.ddim_steps):
                xs.graph.replay()
            h1.record()
            barrier()
            ddim = max_over_ranks(h0.elapsed_time(h1))
            del xs

    # ---- fp32 engines (the reference's own arithmetic, TF32 off): evals/s at a small batch of the exact CUDA-core engine and of the
    #      split-bf16 tensor-core engine (six bf16 products per fp32 product, fp32 accumulation flushed every 32 MMAs)
    fp32 = None
    if args.fp32_steps > 0 and dtype == torch.bfloat16:
        from rho_diffusion.b200 import set_compute_dtype
        Bf = min(B, args.fp32_batch)
        res, outs = {}, {}
        with torch.inference_mode():
            set_compute_dtype(ddpm, torch.float32)
            xf, yf, tf = x_dev[:Bf].contiguous(), cond_dev[:Bf].contiguous(), torch.full((Bf,), 500, dtype=torch.int64, device=device)
            for name, split in (("cuda_core", False), ("split_bf16_tcgen05", True)):
                ops.set_fp32_split(split)
                outs[name] = ddpm.backbone(xf, tf, yf).double()
                barrier()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for _ in range(args.fp32_steps):
                    ddpm.backbone(xf, tf, yf)
                p1.record()
                barrier()
                res[name] = max_over_ranks(p0.elapsed_time(p1))
            ops.set_fp32_split(False)
            set_compute_dtype(ddpm, dtype)
        diff = outs["split_bf16_tcgen05"] - outs["cuda_core"]
        fp32 = {"metric": "UNet denoise evals/s, exact fp32 engine (TF32 off, CUDA cores)", "unit": "evals/s", "batch_per_gpu": Bf, "steps": args.fp32_steps,
                "value": world * Bf * args.fp32_steps / (res["cuda_core"] * 1e-3), "ms_per_eval_batch": res["cuda_core"] / args.fp32_steps,
                "split_bf16": {"metric": "same eval on the split-bf16 tensor-core engine (RHO_FP32_SPLIT=1)",
                               "value": world * Bf * args.fp32_steps / (res["split_bf16_tcgen05"] * 1e-3),
                               "ms_per_eval_batch": res["split_bf16_tcgen05"] / args.fp32_steps,
                               "speedup": res["cuda_core"] / res["split_bf16_tcgen05"],
                               "rel_l2_vs_cuda_core": float(diff.norm() / outs["cuda_core"].norm())}}
        del outs

    # ---- training leg (samples/s): q_sample -> UNet fwd/bwd -> fused AdamW; N > 1: bucketed NCCL all-reduce overlapped with backward
    train = None
    train_steps = args.train_steps if cfg["train_batch"] else 0
    if train_steps > 0:
        Bt = args.train_batch if args.train_batch > 0 else cfg["train_batch"]
        xt, yt = make_inputs(cfg, Bt, 100 + rank)
        xt, yt = xt.to(device), yt.to(device)
        opt = ddpm.configure_optimizers(world)["optimizer"]
        ddpm, fused = optimize(ddpm, optimizer=opt, dtype=dtype)
        ddpm.reseed_per_rank(777)
        ddpm.train()

        def train_step():
            fused.zero_grad()
            loss = ddpm.training_step([xt, yt])
            loss.backward()
            fused.step()
            return loss

        mode = "eager"
        if args.train_graph:
            # the whole step (q_sample .. fused AdamW, all-reduce buckets included) as one replayed CUDA graph
            from rho_diffusion.b200 import GraphedTrainStep
            graphed = GraphedTrainStep(ddpm, fused, xt, yt)
            train_step = lambda: graphed(xt, yt)      # noqa: E731
            mode = "cuda_graph"
        for _ in range(args.train_warmup):
            train_step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(train_steps):
            loss = train_step()
        g1.record()
        barrier()
        ms_train = max_over_ranks(g0.elapsed_time(g1))
        identical = None
        if world > 1:
            # multi-GPU correctness carried by the line: every rank holds bit-identical parameters after the timed steps
            chk = fused.flat_p.double().sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            identical = bool(lo.item() == hi.item())
            assert identical, "data-parallel replicas diverged"
        train = {"ms": ms_train, "batch": Bt, "steps": train_steps, "loss": float(loss.detach()), "mode": mode, "identical": identical,
                 "collective": fused.collective_description() if world > 1 else "none"}
        if args.train_graph:
            graphed.close()          # a graph holding NCCL collectives must go before the process group does

    ms = max_over_ranks(ms)
    ms_e2e = max_over_ranks(ms_e2e)

    if rank == 0:
        conv = prof.get("conv", {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        traffic = ncu_traffic(B)
        attn = prof.get("attention", {"ms": 0.0, "flops": 0.0, "launches": 0})
        gn = prof.get("gn_apply", {"ms": 0.0, "bytes": 0.0, "launches": 0})
        total_prof_ms = sum(v["ms"] for v in prof.values())
        conv_sec = conv["ms"] * 1e-3
        conv_tflops = conv["flops"] / conv_sec / 1e12 if conv_sec > 0 else 0.0
        conv_tflops_exec = conv_exec_flops / conv_sec / 1e12 if conv_sec > 0 else 0.0
        peak = pk["bf16_tflops_sustained"] if dtype == torch.bfloat16 else None
        flops_per_eval = (conv["flops"] + attn["flops"]) / B
        roofline = {
            "kernel": "conv_nd implicit GEMM (forward)", "bound": "tensor", "achieved": conv_tflops, "peak": peak,
            "unit": "TFLOP/s", "frac": (conv_tflops / peak) if peak else None,
            "achieved_executed": conv_tflops_exec, "frac_executed": (conv_tflops_exec / peak) if peak else None,
            "flop_note": "achieved = algorithmic FLOP (reference formula, 27 taps for the fused-upsample convs) / CUDA-event time; "
                         "achieved_executed = FLOP the MMAs actually run (fused upsample: 12 of 27 taps)",
            "traffic": (traffic["dram_bytes"] / traffic["launches"]) if traffic else None,
            "traffic_note": ("DRAM read+write bytes per launch, averaged over the %d tcgen05 conv launches of one eval (ncu, %s); "
                             "algorithmic bytes per launch (inputs + weights + outputs once): %.3e"
                             % (traffic["launches"], traffic.get("file", "profiles/"), conv["bytes"] / max(conv["launches"], 1))) if traffic else None,
            "peak_source": pk["source"] + " (bf16 sustained)" if peak else "n/a for the fp32 CUDA-core engine",
            "launches_per_eval": conv["launches"], "share_of_eval_time": conv["ms"] / total_prof_ms if total_prof_ms else None,
            "algorithmic_gflop_per_eval_per_sample": flops_per_eval / 1e9,
            "whole_step_tflops": flops_per_eval * B / (ms / args.steps * 1e-3) / 1e12,
            "others": {
                "attention_tflops": attn["flops"] / (attn["ms"] * 1e-3) / 1e12 if attn["ms"] > 0 else 0.0,
                "attention_share": attn["ms"] / total_prof_ms if total_prof_ms else None,
                "gn_apply_gbs": gn["bytes"] / (gn["ms"] * 1e-3) / 1e9 if gn["ms"] > 0 else 0.0,
                "gn_apply_frac_of_hbm_peak": (gn["bytes"] / (gn["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"]) if gn["ms"] > 0 else None,
                "gn_share": (gn["ms"] + prof.get("gn_stats", {"ms": 0.0})["ms"]) / total_prof_ms if total_prof_ms else None,
            },
        }
        cpu = cpu_baseline_leg(args) if (world == 1 and not args.no_cpu_baseline) else None
        value = world * B * args.steps / (ms * 1e-3)
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
        line = {
            "metric": "UNet denoise evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"batch-sharded x{world}, no collective",
                       "l2": "working set per eval (>1 GB) exceeds the 126 MB L2; no explicit flush", "cuda_graph": True},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(x_pin.numel() * 4 + y_pin.numel() * 4),
                    "d2h_bytes_per_step": int(x_out_pin.numel() * 4)},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clk, "roofline": roofline,
        }
        if traj is not None:
            line["trajectory"] = traj
        if ddim is not None:
            line["ddim"] = {"metric": "UNet denoise evals/s, GaussianDiffusionPipeline DDIM step (x0 prediction + dynamic thresholding)",
                            "value": world * B * args.ddim_steps / (ddim * 1e-3), "unit": "evals/s", "steps": args.ddim_steps,
                            "ms_per_step": ddim / args.ddim_steps}
        if fp32 is not None:
            line["fp32"] = fp32
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if train is not None:
            sps = world * train["batch"] * train["steps"] / (train["ms"] * 1e-3)
            tf = cfg["train_tflop_per_sample"]
            line["train"] = {"metric": "train samples/s", "value": sps, "unit": "samples/s", "ms_per_step": train["ms"] / train["steps"],
                             "batch_per_gpu": train["batch"], "steps": train["steps"], "warmup": args.train_warmup, "last_loss": train["loss"],
                             "launch": train["mode"], "collective": train["collective"], "replicas_identical": train["identical"],
                             "roofline": {"bound": "tensor", "achieved": sps * tf / world, "peak": peak, "unit": "TFLOP/s",
                                          "frac": (sps * tf / world / peak) if peak else None,
                                          "algorithmic_tflop_per_sample": tf,
                                          "note": "per-GPU: 3 x conv / linear FLOP + 4 x attention FLOP (SURVEY 8d)"} if tf else None}
        emit(line)
    if world > 1:
        # the line is out: never let communicator teardown turn a finished run into a hang
        threading.Timer(60.0, lambda: os._exit(0)).start()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything libraries print (the NCCL version banner, torchrun notices)
    was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)           # fd 1 -> stderr for the rest of the process (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("RHO_BENCH_CONFIG", "D"), choices=sorted(CONFIGS),
                    help="D: 3-D 64^3 (the configuration the metric is quoted on); E: 3-D 128^3 sampling")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("RHO_BENCH_BATCH", "0")), help="sampling batch per GPU (0 = the config's)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--train-steps", type=int, default=int(os.environ.get("RHO_BENCH_TRAIN_STEPS", "50")))
    ap.add_argument("--train-warmup", type=int, default=int(os.environ.get("RHO_BENCH_TRAIN_WARMUP", "10")))
    ap.add_argument("--train-batch", type=int, default=int(os.environ.get("RHO_BENCH_TRAIN_BATCH", "0")))
    ap.add_argument("--trajectory-steps", type=int, default=int(os.environ.get("RHO_BENCH_TRAJECTORY_STEPS", str(T_STEPS))),
                    help="steps of the full-trajectory leg through DDPM.generate() (0 = skip; default: the full T = 1000)")
    ap.add_argument("--ddim-steps", type=int, default=int(os.environ.get("RHO_BENCH_DDIM_STEPS", "5")),
                    help="timed steps of the GaussianDiffusionPipeline DDIM sampler (0 = skip)")
    ap.add_argument("--fp32-steps", type=int, default=int(os.environ.get("RHO_BENCH_FP32_STEPS", "3")))
    ap.add_argument("--fp32-batch", type=int, default=2)
    ap.add_argument("--no-train-graph", dest="train_graph", action="store_false",
                    help="run the training leg eagerly instead of through GraphedTrainStep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--launch-table", default=None, help="write a per-kernel-shape table (CUDA-event times of one eager eval) to this CSV")
    args = ap.parse_args()
    if args.impl != "reference" and args.warmup < 3:
        args.warmup = 3   # timing hygiene: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
