"""The ancestral sampling loop as a replayed CUDA graph (replaces the host-driven loop of ddpm.py:195-227).

One graph = one UNet evaluation + the fused p_sample update (in-kernel Philox noise keyed by the step index) +
the timestep decrement.  The step index and the (B,) timestep vector live in device memory and are advanced by
a kernel inside the graph, so the same graph is replayed for every step; nothing is launched from Python in
between and there is no host synchronisation inside the loop.
"""
from __future__ import annotations

import torch

from . import ops


def shard_range(total: int, world: int, rank: int):
    """Contiguous batch shard of rank ``rank`` (sampling shards the batch, no communication)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class AncestralSampler:
    def __init__(self, backbone, shape, device, alpha, beta, alpha_bar, cond_example=None, use_cuda_graph=True,
                 external_noise=False, noise_scale=0.8, clamp=(-1.0, 1.0)):
        self.backbone = backbone
        self.shape = tuple(shape)
        self.device = device
        self.alpha, self.beta, self.alpha_bar = alpha.contiguous(), beta.contiguous(), alpha_bar.contiguous()
        self.T = int(alpha_bar.numel())
        self.noise_scale, self.clamp = noise_scale, clamp
        self.use_graph = bool(use_cuda_graph)
        with torch.inference_mode(False):      # state buffers outlive the (inference-mode) call that builds the sampler
            self.x = torch.zeros(self.shape, dtype=torch.float32, device=device)
            self.tt = torch.zeros((self.shape[0],), dtype=torch.int64, device=device)
            self.t_scalar = torch.zeros((1,), dtype=torch.int64, device=device)
            self.cond = torch.empty_like(cond_example).copy_(cond_example) if cond_example is not None else None
            self.z = torch.zeros(self.shape, dtype=torch.float32, device=device) if external_noise else None
            self.seed_dev = torch.zeros((1,), dtype=torch.int64, device=device)   # Philox key of the run, read by the step kernel
        self.graph = None
        self._graph_sig = None

    def _step(self, seed=0):
        eps = self.backbone(self.x, self.tt, self.cond) if self.cond is not None else self.backbone(self.x, self.tt)
        ops.p_sample_step(self.x, eps, self.z, self.t_scalar, self.alpha, self.beta, self.alpha_bar, self.noise_scale, self.clamp,
                          seed, seed_dev=self.seed_dev)
        ops.advance_timestep(self.t_scalar, self.tt)

    def _capture(self, seed=0):
        self.seed_dev.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)
        # warm-up on a side stream (packs weights, sets kernel attributes, sizes the workspace), then capture
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.t_scalar.fill_(1)
            self.tt.fill_(1)
            self._step(seed)
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        self.t_scalar.fill_(1)
        self.tt.fill_(1)
        with torch.cuda.graph(g):
            self._step(seed)
        # the eager warm-up above packed the conv weights; the graph reads those snapshots, so it is only valid for this
        # state of the parameters (optimizer steps, load_state_dict and EMA swaps change the signature -> re-capture)
        self.graph, self._graph_sig = g, ops.weights_signature(self.backbone)

    def run(self, x_init, cond=None, seed=0, noises=None, t_checkpoints=None, num_steps=None, progress=False):
        T = self.T if num_steps is None else int(num_steps)
        if self.use_graph and (self.graph is None or self._graph_sig != ops.weights_signature(self.backbone)):
            self._capture(seed)
        self.seed_dev.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)
        self.x.copy_(x_init.reshape(self.shape).to(torch.float32))
        if self.cond is not None:
            self.cond.copy_(cond)
        self.tt.fill_(T - 1)
        self.t_scalar.fill_(T - 1)
        buffer = None
        n_ckpt = 0
        if t_checkpoints is not None:
            n_ckpt = len(t_checkpoints)
            buffer = torch.zeros((self.shape[0], n_ckpt, *self.shape[1:]), dtype=torch.float32, device=self.device)
        per_ckpt = max(T // 10, 1)
        idx = 0
        for t in range(T - 1, -1, -1):
            if self.z is not None:
                if t > 1:
                    self.z.copy_(noises[t].to(self.device))
                else:
                    self.z.zero_()
            if self.use_graph:
                self.graph.replay()
            else:
                self._step(seed)
            if buffer is not None and t % per_ckpt == 0 and idx < n_ckpt:
                buffer[:, idx].copy_(self.x)
                idx += 1
        return {"buffer": buffer, "denoised": self.x.clone()}


class X0Sampler:
    """DDIM (eta = 0) loop of GaussianDiffusionPipeline.reverse_process (gaussian_diffusion.py:1029-1099) for an
    x0-predicting backbone: one graph = UNet eval + per-sample |x0| quantile (dynamic threshold) + fused DDIM update +
    timestep decrement.  With eta = 0 the trajectory is a deterministic function of x_T."""

    def __init__(self, backbone, shape, device, tables, percentile=0.9, cond_example=None, use_cuda_graph=True, eta=0.0):
        assert eta == 0.0, "the reference's reverse_process calls ddim_sample with its default eta = 0"
        self.backbone = backbone
        self.shape = tuple(shape)
        self.device = device
        self.tables = tables
        self.T = int(tables["alphas_cumprod"].numel())
        self.percentile = percentile
        self.use_graph = bool(use_cuda_graph)
        self.x = torch.zeros(self.shape, dtype=torch.float32, device=device)
        self.tt = torch.zeros((self.shape[0],), dtype=torch.int64, device=device)
        self.t_scalar = torch.zeros((1,), dtype=torch.int64, device=device)
        self.thresh = torch.ones((self.shape[0],), dtype=torch.float32, device=device)
        self.cond = cond_example.clone() if cond_example is not None else None
        self.graph = None
        self._graph_sig = None

    def _step(self):
        out = self.backbone(self.x, self.tt, self.cond) if self.cond is not None else self.backbone(self.x, self.tt)
        out = out.float().contiguous()
        ops.abs_quantile(out, self.percentile, floor_value=1.0, out=self.thresh)
        ops.x0_step(out, self.x, self.tables, mode=0, t_scalar=self.t_scalar, thresh=self.thresh)
        ops.advance_timestep(self.t_scalar, self.tt)

    def _capture(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.t_scalar.fill_(1)
            self.tt.fill_(1)
            self._step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        self.t_scalar.fill_(1)
        self.tt.fill_(1)
        with torch.cuda.graph(g):
            self._step()
        self.graph, self._graph_sig = g, ops.weights_signature(self.backbone)

    def run(self, x_init, cond=None, t_checkpoints=None, num_steps=None):
        T = self.T if num_steps is None else int(num_steps)
        if self.use_graph and (self.graph is None or self._graph_sig != ops.weights_signature(self.backbone)):
            self._capture()      # stale packed-weight snapshots after training / load_state_dict / EMA swap
        self.x.copy_(x_init.reshape(self.shape).to(torch.float32))
        if self.cond is not None:
            self.cond.copy_(cond)
        self.tt.fill_(T - 1)
        self.t_scalar.fill_(T - 1)
        buffer, n_ckpt, idx = None, 0, 0
        if t_checkpoints is not None:
            n_ckpt = len(t_checkpoints)
            buffer = torch.zeros((self.shape[0], n_ckpt, *self.shape[1:]), dtype=torch.float32, device=self.device)
        per_ckpt = T // n_ckpt if n_ckpt else T          # gaussian_diffusion.py:1048
        for t in range(T - 1, -1, -1):
            if self.use_graph:
                self.graph.replay()
            else:
                self._step()
            if buffer is not None and t % per_ckpt == 0 and idx < n_ckpt:
                buffer[:, idx].copy_(self.x)
                idx += 1
        return {"buffer": buffer, "denoised": self.x.clone()}


class HostStagedStepper:
    """One sampler step per call with HOST buffers on both sides (the serving form of the loop: x_t and the labels arrive from the
    host, x_{t-1} goes back to the host every step).  ``step(x_pinned, cond_pinned, out_pinned)`` uploads the inputs and downloads
    the result on a second stream through double-buffered device staging tensors, so step k's upload and step k-1's download
    overlap step k's kernels; the compute stream only does two device-to-device copies around the graph replay.  Call
    ``finish()`` (or synchronise) before reading the last ``out_pinned``.  ``sampler`` is an :class:`AncestralSampler` or an
    :class:`X0Sampler` whose step has been captured (``use_cuda_graph=True``) or not (eager steps)."""

    def __init__(self, sampler):
        self.sampler = sampler
        dev = sampler.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        with torch.inference_mode(False):
            self.stage_in = [torch.empty_like(sampler.x) for _ in range(2)]
            self.stage_cond = [torch.empty_like(sampler.cond) for _ in range(2)] if sampler.cond is not None else None
            self.stage_out = [torch.empty_like(sampler.x) for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]
        self.ev_out_free = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        self.h2d_bytes = sampler.x.numel() * 4 + (sampler.cond.numel() * sampler.cond.element_size() if sampler.cond is not None else 0)
        self.d2h_bytes = sampler.x.numel() * 4

    def step(self, x_host, cond_host=None, out_host=None):
        s = self.sampler
        k = self.k & 1
        self.k += 1
        main = torch.cuda.current_stream(s.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_free[k])               # the compute stream has consumed this staging slot
            self.stage_in[k].copy_(x_host, non_blocking=True)         # host -> device (pinned)
            if self.stage_cond is not None:
                self.stage_cond[k].copy_(cond_host, non_blocking=True)
            self.ev_in[k].record(self.copy_stream)
        main.wait_event(self.ev_in[k])
        s.x.copy_(self.stage_in[k], non_blocking=True)
        if self.stage_cond is not None:
            s.cond.copy_(self.stage_cond[k], non_blocking=True)
        self.ev_free[k].record(main)
        if s.use_graph:
            s.graph.replay()
        else:
            s._step()
        if out_host is not None:
            main.wait_event(self.ev_out_free[k])                       # the download of step k-2 has left this slot
            self.stage_out[k].copy_(s.x, non_blocking=True)
            self.ev_done[k].record(main)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.ev_done[k])
                out_host.copy_(self.stage_out[k], non_blocking=True)   # device -> host (pinned) of this step's result
                self.ev_out_free[k].record(self.copy_stream)

    def finish(self):
        """Makes the current stream wait for the outstanding download."""
        torch.cuda.current_stream(self.sampler.device).wait_stream(self.copy_stream)
