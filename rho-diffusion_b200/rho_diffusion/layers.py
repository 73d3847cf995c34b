"""n-D layer factories and normalisation (reference ``rho_diffusion/layers.py``), executed by the B200 kernels.

Every module keeps the reference's parameter names, shapes and initialisation (they subclass the torch
modules the reference instantiates, so ``state_dict()`` is interchangeable), but ``forward`` launches the
hand-written sm_100a kernels through the C ABI instead of ATen.  Modules accept the reference's
``(N, C, *spatial)`` layout when called on their own; inside ``UNetv2`` the fused channels-last entry points
(``forward_cl``) are used and the layout conversion happens once at the model boundary.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn
from torch.nn.parameter import Parameter

from rho_diffusion.registry import registry

__all__ = ["RMSNorm", "GroupNorm32", "conv_nd", "avg_pool_nd", "mean_flat", "zero_module", "normalization", "checkpoint"]


@dataclass
class ExecCtx:
    """Execution options threaded through the fused channels-last forward."""
    dtype: torch.dtype = torch.float32   # storage dtype of activations (fp32 exact path or bf16 tensor-core path)
    engine: int = 0                      # RHO_CONV_AUTO


def _ops():
    from rho_diffusion.b200 import ops
    return ops


def module_ctx(module: nn.Module) -> ExecCtx:
    return ExecCtx(dtype=getattr(module, "_rho_dtype", torch.float32), engine=getattr(module, "_rho_engine", 0))


@registry.register_layer("RMSNorm")
class RMSNorm(nn.Module):
    """Registered for API parity (layers.py:37-68); never instantiated by UNetv2, so it has no kernel."""

    def __init__(self, input_dim: int, eps: float = 1e-8, bias: bool = False) -> None:
        super().__init__()
        self.input_dim, self.eps, self.has_bias = input_dim, eps, bias
        self.scale = Parameter(torch.ones(input_dim))
        if bias:
            self.bias = Parameter(torch.zeros(input_dim))

    def forward(self, data: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("RMSNorm is outside the B200 hot path (unused by UNetv2 / DDPM)")


@registry.register_layer("GroupNorm32")
class GroupNorm32(nn.GroupNorm):
    """GroupNorm computed in fp32 and cast back (layers.py:71-74)."""

    def forward_cl(self, x, x2=None, film=None, pre_add=None, act=0, ctx: ExecCtx = None):
        ctx = ctx or module_ctx(self)
        return _ops().GnActFn.apply(x, x2, self.weight, self.bias, film, pre_add, act, ctx.dtype, self.num_groups, self.eps)

    def forward(self, x):
        ops = _ops()
        ctx = module_ctx(self)
        h = ops.to_channels_last(x, ctx.dtype)
        return ops.to_channels_first(self.forward_cl(h, ctx=ctx), x.dtype)


class _ConvMixin:
    """forward / forward_cl shared by the three conv classes."""

    def _spec(self, up=1):
        ops = _ops()
        dims = len(self.kernel_size)
        if self.groups != 1 or any(d != 1 for d in self.dilation) or self.padding_mode != "zeros":
            raise NotImplementedError("conv_nd on B200 supports groups=1, dilation=1, zero padding")
        if isinstance(self.padding, str) or any(p != k // 2 for p, k in zip(self.padding, self.kernel_size)):
            raise NotImplementedError(f"conv_nd on B200 supports 'same' padding k//2 only (got {self.padding} for k={self.kernel_size})")
        return ops.ConvSpec.make(dims, self.kernel_size, self.stride, up)

    def forward_cl(self, x, x2=None, residual=None, up=1, y_f32=False, ctx: ExecCtx = None):
        ctx = ctx or module_ctx(self)
        return _ops().conv(self, x, x2, self.weight, self.bias, residual, self._spec(up), ctx.dtype, ctx.engine, y_f32)

    def forward(self, x):
        ops = _ops()
        ctx = module_ctx(self)
        h = ops.to_channels_last(x, ctx.dtype)
        return ops.to_channels_first(self.forward_cl(h, ctx=ctx), x.dtype)


class Conv1d(_ConvMixin, nn.Conv1d):
    pass


class Conv2d(_ConvMixin, nn.Conv2d):
    pass


class Conv3d(_ConvMixin, nn.Conv3d):
    pass


@registry.register_layer("conv_nd")
def conv_nd(dims, *args, **kwargs):
    """Create a 1D, 2D, or 3D convolution module (layers.py:77-88)."""
    try:
        return {1: Conv1d, 2: Conv2d, 3: Conv3d}[dims](*args, **kwargs)
    except KeyError:
        raise ValueError(f"unsupported dimensions: {dims}") from None


class AvgPoolNd(nn.Module):
    """nn.AvgPool{1,2,3}d with kernel_size == stride and no padding (the only form the UNet builds, unet_v2.py:163-165)."""

    def __init__(self, dims, kernel_size, stride=None, padding=0, **kwargs):
        super().__init__()
        stride = kernel_size if stride is None else stride
        k = (kernel_size,) * dims if isinstance(kernel_size, int) else tuple(kernel_size)
        s = (stride,) * dims if isinstance(stride, int) else tuple(stride)
        if k != s or padding not in (0, (0,) * dims) or any(kwargs.get(a) for a in ("ceil_mode", "divisor_override")):
            raise NotImplementedError("avg_pool_nd on B200 supports kernel_size == stride, no padding, floor mode")
        self.dims, self.kernel_size, self.stride = dims, k, s

    def forward_cl(self, h, ctx: ExecCtx = None):
        return _ops().avg_pool(h, self.kernel_size)

    def forward(self, x):
        ops = _ops()
        ctx = module_ctx(self)
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x, ctx.dtype), ctx), x.dtype)


@registry.register_layer("avg_pool_nd")
def avg_pool_nd(dims, *args, **kwargs):
    """Create a 1D, 2D, or 3D average pooling module (layers.py:90-102)."""
    if dims not in (1, 2, 3):
        raise ValueError(f"unsupported dimensions: {dims}")
    return AvgPoolNd(dims, *args, **kwargs)


@registry.register_layer("mean_flat")
def mean_flat(tensor):
    """Mean over all non-batch dimensions (layers.py:105-110); host-side helper, not on the DDPM path."""
    return tensor.mean(dim=list(range(1, len(tensor.shape))))


def zero_module(module):
    """Zero out the parameters of a module and return it (layers.py:113-119)."""
    for p in module.parameters():
        p.detach().zero_()
    return module


def normalization(channels):
    """GroupNorm32(32, channels) (layers.py:122-129)."""
    return GroupNorm32(32, channels)


def checkpoint(func, inputs, params, flag):
    """API-compatible with layers.py:153-199.  The reference recomputes activations in backward to save memory
    (always for attention, where it stores a T x T matrix); the B200 kernels keep only O(T) state (flash-style
    log-sum-exp) and recompute inside the backward kernels, so no Python-level re-execution is needed and the
    result is numerically the same computation."""
    return func(*inputs)
