"""Legacy v1 UNet (``"UNet"``, ``UNetBlock2d`` / ``UNetBlock3d``) on the B200 kernels.

Same constructors, registry names and ``state_dict`` layout as the reference (``rho_diffusion/models/unet.py:30-269``): a stack of
stride-1 blocks ``act(conv1) -> act(conv2) (+ residual_conv(x)) + time embedding -> GroupNorm(8) -> act`` without resampling or
attention, the up path built from ``ConvTranspose`` layers and the skip concat.  Execution:

* activations are channels-last (bf16 or fp32); the skip concat is never materialised (two-source conv);
* a stride-1 ``ConvTranspose(k, padding=k//2)`` IS a convolution with the ``[Cin, Cout, *k]`` weight transposed and flipped:
  one small kernel builds that weight (``rho_convt_weight``, its own adjoint for the gradient) and the layer runs on the conv
  engines like every other conv;
* ``h + residual_conv(x)`` is the conv epilogue's residual add; the time-embedding add is the GroupNorm kernel's ``pre_add``;
  the two bare activations are one element-wise kernel each (``rho_activation``).

Reference quirk: the block adds ``time_pe[(...,) + (None,) * 2]`` (unet.py:130-132), which only broadcasts correctly against
2-D data; with ``UNetBlock3d`` the reference raises a shape error (or silently mis-broadcasts when N == C == D).  Here the
embedding is added per sample and channel for 2-D and 3-D alike.
"""
from __future__ import annotations

import torch
from torch import nn

from rho_diffusion.layers import Conv2d, Conv3d, ExecCtx, module_ctx
from rho_diffusion.models.common import SinusoidalPositionEmbedding
from rho_diffusion.registry import registry

__all__ = ["UNetBlock2d", "UNetBlock3d", "UNet"]


def _ops():
    from rho_diffusion.b200 import ops
    return ops


class _ConvTransposeMixin:
    """forward / forward_cl of a stride-1, 'same'-padded ConvTranspose through the conv engines."""

    def _spec(self):
        dims = len(self.kernel_size)
        if (self.groups != 1 or any(d != 1 for d in self.dilation) or any(s != 1 for s in self.stride) or any(o != 0 for o in self.output_padding)
                or any(p != k // 2 for p, k in zip(self.padding, self.kernel_size)) or any(k % 2 == 0 for k in self.kernel_size)):
            raise NotImplementedError("ConvTranspose on B200 supports the v1 UNet's form only: stride 1, odd kernel, padding k//2, groups=1")
        return _ops().ConvSpec.make(dims, self.kernel_size, 1, 1)

    def forward_cl(self, x, x2=None, residual=None, ctx: ExecCtx = None):
        ops = _ops()
        ctx = ctx or module_ctx(self)
        w = ops.ConvTWeightFn.apply(self.weight)               # [Cout, Cin, *k], flipped
        return ops.conv_uncached(x, x2, w, self.bias, residual, self._spec(), ctx.dtype, ctx.engine)

    def forward(self, x, output_size=None):
        ops = _ops()
        ctx = module_ctx(self)
        h = ops.to_channels_last(x, ctx.dtype)
        return ops.to_channels_first(self.forward_cl(h, ctx=ctx), x.dtype)


class ConvTranspose2d(_ConvTransposeMixin, nn.ConvTranspose2d):
    pass


class ConvTranspose3d(_ConvTransposeMixin, nn.ConvTranspose3d):
    pass


class AbstractUNetBlock(nn.Module):
    """Convolution building block with time-embedding injection and optional residual branch (unet.py:30-136)."""

    __conv_class__ = None
    __transpose_class__ = None

    def __init__(self, in_channels: int, out_channels: int, time_embedding_dim: int, is_up: bool = False, kernel_size: int = 3,
                 stride: int = 1, padding: int = 1, groups: int = 8, activation="GELU", residual: bool = False) -> None:
        super().__init__()
        conv, convt = self.__conv_class__, self.__transpose_class__
        kw = dict(out_channels=out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        self.time_embedding_readout = nn.Linear(time_embedding_dim, out_channels)
        if is_up:
            self.conv1 = conv(in_channels=2 * in_channels, **kw)
            self.conv2 = convt(in_channels=out_channels, **kw)
        else:
            self.conv1 = conv(in_channels=in_channels, **kw)
            self.conv2 = conv(in_channels=out_channels, **kw)
        if residual:
            self.residual_conv = convt(in_channels=2 * in_channels, **kw) if is_up else conv(in_channels=in_channels, **kw)
        else:
            self.residual_conv = None
        self.norm = None if groups == 0 else nn.GroupNorm(groups, out_channels)
        if isinstance(activation, str):
            activation = registry.get("activations", activation)()
        self.activation = activation
        self._act = _ops().act_code(activation)

    def forward_cl(self, x, x2, t, ctx: ExecCtx):
        """``x`` (and the optional skip ``x2``: their channel concat is the block input) channels-last; ``t`` [N, time_dim] fp32."""
        ops = _ops()
        lin = self.time_embedding_readout
        time_pe = ops.linear(t, lin.weight, lin.bias)                              # [N, Cout] fp32
        if self.residual_conv is not None:
            (xa, x), (x2a, x2) = ops.fork(x), ops.fork(x2)                         # main path and residual branch
        else:
            xa, x2a = x, x2
        h = ops.activation(self.conv1.forward_cl(xa, x2a, ctx=ctx), self._act)
        h = ops.activation(self.conv2.forward_cl(h, ctx=ctx), self._act)
        if self.residual_conv is not None:
            h = self.residual_conv.forward_cl(x, x2, residual=h, ctx=ctx)          # h + residual_conv(x)
        if self.norm is None:
            # no kernel fuses "add the embedding, activate" without the norm; the reference never builds this (groups = 8 always)
            raise NotImplementedError("UNetBlock with groups=0 has no B200 path")
        return ops.GnActFn.apply(h, None, self.norm.weight, self.norm.bias, None, time_pe, self._act, ctx.dtype, self.norm.num_groups,
                                 self.norm.eps)

    def forward(self, x, t):
        ops = _ops()
        ctx = module_ctx(self)
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x, ctx.dtype), None, t, ctx), x.dtype)


@registry.register_layer("UNetBlock2d")
class UNetBlock2d(AbstractUNetBlock):
    __conv_class__ = Conv2d
    __transpose_class__ = ConvTranspose2d


@registry.register_layer("UNetBlock3d")
class UNetBlock3d(AbstractUNetBlock):
    __conv_class__ = Conv3d
    __transpose_class__ = ConvTranspose3d


@registry.register_model("UNet")
class UNet(nn.Module):
    """v1 UNet for diffusion (unet.py:152-269): ``forward(data, t)``."""

    def __init__(self, block_type, input_channels: int, down_channels=[64, 128, 256], up_channels=[256, 128, 64],
                 time_embedding_dim: int = 32, kernel_size: int = 3, padding: int = 1, activation="ReLU", residual: bool = True) -> None:
        super().__init__()
        if isinstance(block_type, str):
            block_type = registry.get("layers", block_type)
        self.time_mlp = nn.Sequential(SinusoidalPositionEmbedding(time_embedding_dim), nn.Linear(time_embedding_dim, time_embedding_dim))
        layer_type = Conv3d if block_type == UNetBlock3d else Conv2d
        self.input_conv = layer_type(in_channels=input_channels, out_channels=down_channels[0], kernel_size=3, stride=1, padding=1)
        self.output_conv = layer_type(up_channels[-1], input_channels, kernel_size=1, stride=1, padding=0)

        def block(cin, cout, is_up):
            return block_type(in_channels=cin, out_channels=cout, time_embedding_dim=time_embedding_dim, is_up=is_up,
                              kernel_size=kernel_size, padding=padding, activation=activation, residual=residual)

        self.downsample = nn.ModuleList([block(down_channels[i], down_channels[i + 1], False) for i in range(len(down_channels) - 1)])
        self.upsample = nn.ModuleList([block(up_channels[i], up_channels[i + 1], True) for i in range(len(up_channels) - 1)])
        self.block_type = block_type

    @property
    def expected_dim(self) -> int:
        return 3 if self.block_type == UNetBlock2d else 4

    def forward(self, data: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        ops = _ops()
        ctx = module_ctx(self)
        lin = self.time_mlp[1]
        time_pe = ops.linear(self.time_mlp[0](t), lin.weight, lin.bias)
        pes = list(ops.fork(time_pe, len(self.downsample) + len(self.upsample)))
        x = self.input_conv.forward_cl(ops.to_channels_last(data, ctx.dtype), ctx=ctx)
        skips = []
        for blk in self.downsample:
            x = blk.forward_cl(x, None, pes.pop(), ctx)
            x, skip = ops.fork(x)
            skips.append(skip)
        for blk in self.upsample:
            x = blk.forward_cl(x, skips.pop(), pes.pop(), ctx)                     # torch.cat((x, skip), dim=1), never materialised
        o = self.output_conv.forward_cl(x, y_f32=True, ctx=ctx)
        return ops.to_channels_first(o, data.dtype if data.dtype in (torch.float32, torch.bfloat16) else torch.float32)
