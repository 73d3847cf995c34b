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

__all__ = ["optimize", "ddp_setup", "FusedAdamW", "set_compute_dtype", "is_available"]


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
                           g["eps"], g["weight_decay"], ema_decay_t=d_t, grad_scale=1.0 / world)


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
