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

__all__ = ["optimize", "ddp_setup", "FusedAdamW", "GraphedTrainStep", "set_compute_dtype", "set_dropout_seed", "is_available"]


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


def set_dropout_seed(model: nn.Module, seed: int) -> nn.Module:
    """Keys the dropout masks of every sub-module by ``seed`` (pipeline seed + 7919 * rank) and the module's position, so that
    layers and ranks draw independent masks (the offset within a module is the call index and the optimizer step)."""
    for idx, m in enumerate(model.modules()):
        m._rho_dropout_seed = (int(seed) * 1000003 + idx) & 0x7FFFFFFFFFFFFFFF
    return model


def partition_buckets(offsets, total: int, n_buckets: int):
    """Cuts the flat buffer ``[0, total)`` into at most ``n_buckets`` contiguous ranges of ~equal size at parameter boundaries
    (``offsets``: start of every parameter, ascending).  Returns ``(ranges, bucket_of)``: ``[(lo, hi), ...]`` covering
    ``[0, total)`` without gaps and the bucket index of every parameter."""
    n = len(offsets)
    n_b = max(1, min(int(n_buckets), n))
    target = total / n_b
    ranges, bucket_of, lo, b = [], [], 0, 0
    for i in range(n):
        end = offsets[i + 1] if i + 1 < n else total
        bucket_of.append(b)
        if (end >= target * (b + 1) and b < n_b - 1) or i == n - 1:
            ranges.append((lo, end))
            lo, b = end, b + 1
    return ranges, bucket_of


class FusedAdamW:
    """torch.optim.AdamW semantics over flat fp32 buffers, fused with the EMA update, one kernel per step.

    * parameters are re-pointed at slices of one contiguous buffer (values preserved; ``state_dict()`` of the model is
      unaffected); ``p.grad`` of every parameter IS a slice of a second flat buffer, and the backward kernels write the
      parameter gradients straight into it (:class:`ops.GradSink`): no per-parameter accumulate launches, no gather copies;
    * with a process group the flat gradient buffer is split into ``buckets`` contiguous buckets (aligned to parameter
      boundaries, ~equal bytes).  The backward produces gradients from the END of the buffer to its start; as soon as the
      last gradient of a bucket has been enqueued, that bucket's NCCL all-reduce (SUM) is launched asynchronously, so the
      exchange overlaps the rest of the backward -- the bucketed, overlapped reduction the reference asks
      ``DistributedDataParallel`` for (scripts/training_ddp.py:171-176) and then loses by calling ``model.module.training_step``
      (SURVEY quick facts).  ``step()`` waits for the buckets and folds the 1/world mean into the update.  All of it is
      stream-ordered, so the whole step including the collectives is capturable in one CUDA graph (GraphedTrainStep);
    * ``ema_decay > 0`` keeps a flat shadow copy with the warm-up schedule of rho_diffusion/ema.py:33.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, ema_decay: float = 0.0,
                 process_group=None, align: int = 64, buckets: int = 8):
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
        self.step_id = 0
        self.step_dev = None          # set by enable_device_step(): the step counter lives on the device (graph replay)
        # ---- buckets: contiguous [lo, hi) ranges of the flat buffers, cut at parameter boundaries
        self.bucket_ranges, self._bucket_of = partition_buckets(offsets, total, buckets)
        self._bucket_size = [0] * len(self.bucket_ranges)
        for bi in self._bucket_of:
            self._bucket_size[bi] += 1
        self._bind_grads()
        self.zero_grad()
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

    # -- gradients ---------------------------------------------------------------------------------------------------------
    def _bind_grads(self):
        self._views = [self.flat_g[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        self._sinks = []
        for p, v, b in zip(self.params, self._views, self._bucket_of):
            p.grad = v                                   # p.grad IS the flat slice (tests / users read gradients there)
            p._rho_sink = ops.GradSink(v, self, b)
            self._sinks.append(p._rho_sink)

    def _world(self) -> int:
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def zero_grad(self, set_to_none: bool = True):
        """The first backward kernel that produces a parameter's gradient OVERWRITES its slice, so only the slices that no
        kernel wrote in the previous step are zeroed (unused parameters; normally none) -- a gradient that reaches such a
        parameter through plain autograd accumulation (``p.grad += g``) then adds onto zero.  The rest is per-step bookkeeping
        (which sinks were written, which buckets are complete / in flight)."""
        stale = [s.view for s in self._sinks if not s.touched]
        if stale and len(stale) < len(self._sinks):
            torch._foreach_zero_(stale)
        elif stale:
            self.flat_g.zero_()
        for s in self._sinks:
            s.touched = False
        self._pending = list(self._bucket_size)
        self._works = []
        self._reduced = [False] * len(self.bucket_ranges)

    def _reduce_bucket(self, b: int):
        lo, hi = self.bucket_ranges[b]
        self._reduced[b] = True
        self._works.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.process_group, async_op=True))

    def _grad_ready(self, b: int):
        """Called by a GradSink right after the kernel that wrote its gradient was enqueued."""
        self._pending[b] -= 1
        if self._pending[b] == 0 and self._world() > 1:
            self._reduce_bucket(b)

    def collective_description(self) -> str:
        mb = [(hi - lo) * 4 / 1e6 for lo, hi in self.bucket_ranges]
        return (f"NCCL all-reduce(SUM) of the flat fp32 gradient buffer in {len(mb)} buckets ({min(mb):.0f}-{max(mb):.0f} MB), each launched "
                "async as soon as its last gradient is enqueued (overlaps the backward); 1/world folded into the AdamW kernel")

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
        ops.dropout_step_dev = self.step_dev      # dropout masks are keyed by the device step count from now on
        return self

    def _snapshot(self):
        return (self.flat_p.clone(), self.flat_m.clone(), self.flat_v.clone(),
                None if self.flat_shadow is None else self.flat_shadow.clone(), self.step_id,
                None if self.step_dev is None else self.step_dev.clone())

    def _restore(self, snap):
        p, m, v, sh, step_id, step_dev = snap
        self.flat_p.copy_(p); self.flat_m.copy_(m); self.flat_v.copy_(v)
        if sh is not None:
            self.flat_shadow.copy_(sh)
        self.step_id = step_id
        if step_dev is not None:
            self.step_dev.copy_(step_dev)
        ops.bump_weights_epoch()

    @property
    def steps_taken(self) -> int:
        return int(self.step_dev[0]) if self.step_dev is not None else self.step_id

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        # slices no kernel wrote this step were zeroed by zero_grad (or hold a gradient accumulated by plain autograd)
        world = self._world()
        if world > 1:
            for b in range(len(self.bucket_ranges)):
                if not self._reduced[b]:
                    self._reduce_bucket(b)
            for w in self._works:
                w.wait()                                  # the compute stream waits for the NCCL stream (no host sync)
            self._works = []
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
    optimizer.step()`` (tests/test_gpu_model.py::test_graphed_train_step_matches_eager).  In a multi-process job the bucketed
    NCCL all-reduces launched from inside the backward are part of the captured graph (fork / join onto NCCL's stream), so
    every rank replays the same step, collectives included."""

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
        # the warm-up runs REAL steps (allocator pools, packed-weight caches, kernel attributes): weights, moments, EMA shadow and
        # the step counters are put back afterwards, so the first replay is optimizer step 1 on the caller's parameters
        snap = optimizer._snapshot()
        noise_calls = getattr(pipeline, "_noise_calls", None)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        optimizer._restore(snap)
        del snap
        if noise_calls is not None:
            pipeline._noise_calls = noise_calls
        self.steps = 0

    def close(self):
        """Releases the captured graph.  A graph that holds NCCL collectives must be released BEFORE the process group is
        destroyed (destroying the communicator with such a graph alive hangs)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

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
    set_dropout_seed(inner, int(getattr(inner, "seed", 0)) + 7919 * int(getattr(inner, "global_rank", 0)))
    if optimizer is None:
        return inner
    fused = optimizer if isinstance(optimizer, FusedAdamW) else FusedAdamW.from_torch(optimizer, ema_decay=ema_decay)
    return inner, fused
