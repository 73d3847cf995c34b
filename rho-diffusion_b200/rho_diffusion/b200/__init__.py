"""B200 backend plug-in: the replacement for the reference's IPEX / XPU / oneCCL hooks.

The reference accelerates a model in three places: ``model, optimizer = ipex.optimize(model,
optimizer=optimizer)`` (scripts/training_ddp.py:182-183, scripts/inference.py:161-163), the Lightning XPU
accelerator / strategies (rho_diffusion/xpu.py) and ``dist.init_process_group("ccl")`` (rho_diffusion/utils.py:45-81).
This module offers the same call shapes on B200:

* :func:`optimize` -- ``model = optimize(model)`` or ``model, optimizer = optimize(model, optimizer=optimizer)``;
  selects the compute dtype (bf16 tensor-core path by default, fp32 exact path on request), flattens parameters
  and gradients into contiguous buffers and swaps the optimizer for the fused AdamW(+EMA) kernel with a single
  NCCL all-reduce of the flat gradient buffer per step;
* :func:`ddp_setup` -- ``(world, rank, local_rank)`` over NCCL (``rho_diffusion.utils.ddp_setup``).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch import nn

from rho_diffusion.utils import ddp_setup  # noqa: F401  (re-export: the backend's process-group entry point)

from . import _lib, ops  # noqa: F401

__all__ = ["optimize", "ddp_setup", "FusedAdamW", "GraphedTrainStep", "set_compute_dtype", "is_available"]


def is_available() -> bool:
    import os
    return os.path.exists(_lib.LIB_PATH) and torch.cuda.is_available()


def set_compute_dtype(model: nn.Module, dtype: torch.dtype, engine: int = 0) -> nn.Module:
    """Marks every sub-module with the activation storage dtype (float32: exact CUDA-core path, TF32 never used;
    bfloat16: tcgen05 tensor-core path with fp32 accumulation)."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("compute dtype must be torch.float32 or torch.bfloat16")
    for m in model.modules():
        m._rho_dtype = dtype
        m._rho_engine = engine
    return model


class FusedAdamW:
    """torch.optim.AdamW semantics over flat fp32 buffers, fused with the EMA update, one kernel per step.

    * parameters are re-pointed at slices of one contiguous buffer (values preserved; ``state_dict()`` of the
      model is unaffected), gradients accumulate into slices of a second one;
    * with a process group, ``step()`` first sums the flat gradient buffer across ranks with ONE NCCL all-reduce
      and folds the 1/world mean into the update (the intended DDP semantics of scripts/training_ddp.py:171-176,
      which the reference loses by calling ``model.module.training_step``, SURVEY quick facts);
    * ``ema_decay > 0`` keeps a flat shadow copy with the warm-up schedule of rho_diffusion/ema.py:33.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, ema_decay: float = 0.0,
                 process_group=None, align: int = 64):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdamW: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.RhoB200Error("FusedAdamW needs CUDA parameters (there is no CPU path)")
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(self.defaults, params=self.params)]
        self.ema_decay = float(ema_decay)
        self.process_group = process_group
        offsets, total = [], 0
        for p in self.params:
            offsets.append(total)
            total += (p.numel() + align - 1) // align * align
        self.offsets, self.total = offsets, total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(self.params, offsets):
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p.detach())
                p.data = view
        self.flat_shadow = self.flat_p.clone() if self.ema_decay > 0 else None
        self._bind_grads()
        self.step_id = 0
        self.step_dev = None          # set by enable_device_step(): the step counter lives on the device (graph replay)
        ops.bump_weights_epoch()

    @classmethod
    def from_torch(cls, optimizer: torch.optim.Optimizer, **kw) -> "FusedAdamW":
        if not isinstance(optimizer, (torch.optim.AdamW,)):
            raise NotImplementedError(f"optimize(): only AdamW (the reference's optimizer) has a fused B200 kernel, got {type(optimizer).__name__}")
        if len(optimizer.param_groups) != 1:
            raise NotImplementedError("optimize(): a single parameter group is supported (as built by configure_optimizers)")
        g = optimizer.param_groups[0]
        if g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("amsgrad / maximize are not supported by the fused kernel")
        return cls(g["params"], lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"], **kw)

    def _bind_grads(self):
        self._views = [self.flat_g[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]

    def zero_grad(self, set_to_none: bool = True):
        """Gradients are dropped (``p.grad = None``): the first backward then hands its weight-gradient tensor to
        ``p.grad`` without a kernel, instead of one accumulate launch per parameter tensor (329 for the 64^3 UNet);
        :meth:`step` gathers them into the flat buffer with a handful of multi-tensor copies."""
        for p in self.params:
            p.grad = None

    def _gather_grads(self):
        dst, src, missing = [], [], []
        for p, v in zip(self.params, self._views):
            if p.grad is None:
                missing.append(v)
            elif p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad if p.grad.dtype == torch.float32 else p.grad.float())
        if src:
            torch._foreach_copy_(dst, src)
        if missing:
            torch._foreach_zero_(missing)

    def shadow_state_dict(self, model: nn.Module) -> dict:
        """EMA weights keyed like ``model.state_dict()`` (for saving the averaged checkpoint)."""
        if self.flat_shadow is None:
            raise RuntimeError("EMA is disabled (ema_decay == 0)")
        by_ptr = {p.data_ptr(): (off, p) for p, off in zip(self.params, self.offsets)}
        out = {}
        for k, v in model.state_dict().items():
            hit = by_ptr.get(v.data_ptr())
            out[k] = self.flat_shadow[hit[0]:hit[0] + v.numel()].view_as(v).clone() if hit else v.clone()
        return out

    def enable_device_step(self):
        """Moves the step counter to device memory so that ``step()`` can be captured in a CUDA graph and replayed: the
        kernel increments it and derives the bias corrections / EMA warm-up factor itself."""
        if self.step_dev is None:
            self.step_dev = torch.zeros(16, dtype=torch.int32, device=self.flat_p.device)
            self.step_dev[0] = self.step_id
        return self

    @property
    def steps_taken(self) -> int:
        return int(self.step_dev[0]) if self.step_dev is not None else self.step_id

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        self._gather_grads()
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
            if world > 1:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.process_group)
        self.step_id += 1
        d_t = self.ema_decay * (1.0 - math.exp(-self.step_id / 2000.0)) if self.flat_shadow is not None else 0.0
        ops.adamw_ema_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_shadow, self.step_id, g["lr"], g["betas"],
                           g["eps"], g["weight_decay"], ema_decay_t=d_t, grad_scale=1.0 / world, step_dev=self.step_dev,
                           ema_decay=self.ema_decay)


class GraphedTrainStep:
    """One whole training step -- q_sample, UNet forward, loss, backward, gradient gather (and all-reduce), fused
    AdamW(+EMA) -- captured ONCE in a CUDA graph and replayed: ~1500 kernel launches and the Python autograd walk per step
    become one ``cudaGraphLaunch``.  The per-step inputs live in static buffers: the batch, the labels, the timesteps (drawn
    on the CPU generator like ``random_timesteps``, abstract_diffusion.py:163-169) and the noise (one Philox launch outside
    the graph, so every step sees fresh noise).

        step = GraphedTrainStep(model, optimizer, example_data, example_labels)
        for data, labels in loader:
            loss = step(data, labels)          # a 0-d device tensor, valid until the next call

    Semantics are those of ``optimizer.zero_grad(); loss = model.training_step([data, labels]); loss.backward();
    optimizer.step()`` (tests/test_gpu_model.py::test_graphed_train_step_matches_eager)."""

    def __init__(self, pipeline, optimizer: "FusedAdamW", data: torch.Tensor, labels: torch.Tensor = None, warmup: int = 2):
        self.pipeline, self.optimizer = pipeline, optimizer
        dev = data.device
        self.data = data.clone()
        self.labels = labels.clone() if labels is not None else None
        self.t = torch.zeros((data.shape[0],), dtype=torch.int64, device=dev)
        self.noise = torch.zeros(data.shape, dtype=torch.float32, device=dev)
        self._t_pin = torch.zeros((data.shape[0],), dtype=torch.int64).pin_memory()
        optimizer.enable_device_step()
        self._draw()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):           # allocator pools, packed-weight caches, kernel attributes
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.steps = warmup                   # optimizer steps already taken by the warm-up

    def _batch(self):
        return [self.data, self.labels] if self.labels is not None else self.data

    def _body(self):
        self.optimizer.zero_grad()
        loss = self.pipeline.training_step(self._batch(), t=self.t, noise=self.noise)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _draw(self):
        self._t_pin.copy_(self.pipeline.random_timesteps(self.data.shape[0]))
        self.t.copy_(self._t_pin, non_blocking=True)
        fresh = self.pipeline.noise(self.data)
        self.noise.copy_(fresh)

    def __call__(self, data: torch.Tensor, labels: torch.Tensor = None, t: torch.Tensor = None, noise: torch.Tensor = None):
        self.data.copy_(data, non_blocking=True)
        if self.labels is not None:
            self.labels.copy_(labels, non_blocking=True)
        if t is None or noise is None:
            self._draw()
        if t is not None:
            self.t.copy_(t, non_blocking=True)
        if noise is not None:
            self.noise.copy_(noise, non_blocking=True)
        self.graph.replay()
        self.steps += 1
        ops.bump_weights_epoch()              # eager users of the modules (sampling, eval) must re-pack the new weights
        return self.loss


def optimize(model: nn.Module, optimizer=None, dtype: torch.dtype = torch.bfloat16, ema_decay: float = 0.0, engine: int = 0,
             **_ignored):
    """Drop-in for ``ipex.optimize(model, optimizer=optimizer)``: returns ``model`` or ``(model, optimizer)``.

    ``model`` may be the pipeline (``DDPM``), its backbone, or a ``DistributedDataParallel`` wrapper around either
    (unwrapped: gradient averaging is done by :class:`FusedAdamW` on the flat buffer, one all-reduce per step).
    """
    _lib.lib()  # fail loudly if the CUDA extension is missing
    inner = model.module if isinstance(model, nn.parallel.DistributedDataParallel) else model
    if any(p.device.type != "cuda" for p in inner.parameters()):
        raise _lib.RhoB200Error("optimize(): move the model to a CUDA device first (there is no CPU path)")
    set_compute_dtype(inner, dtype, engine)
    if optimizer is None:
        return inner
    fused = optimizer if isinstance(optimizer, FusedAdamW) else FusedAdamW.from_torch(optimizer, ema_decay=ema_decay)
    return inner, fused
