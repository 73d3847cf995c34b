"""Conditional n-D UNet ("UNetv2") on the B200 kernels.

Same constructor, ``forward(x, timesteps, y=None)`` contract, registry name and ``state_dict`` key names /
tensor layouts as the reference model (``rho_diffusion/models/unet_v2.py:439-732``), so reference
checkpoints load with ``strict=True``.  Execution is different by design:

* activations are channels-last (bf16 or fp32) between layers; the (N, C, *spatial) layout exists only at the
  model boundary;
* each ResBlock is 2 GroupNorm-stat launches + 2 fused GN(+FiLM)+SiLU launches + 2 implicit-GEMM convs whose
  epilogues add bias and the residual; the skip concat (``torch.cat``, unet_v2.py:729) is never materialised
  (two-source K loop in GN and conv), nearest upsampling (unet_v2.py:124-131) is folded into the conv gather;
* attention is one flash-style kernel on the channels-last qkv, the T x T weights never touch HBM.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from rho_diffusion.layers import ExecCtx, avg_pool_nd, conv_nd, normalization, zero_module
from rho_diffusion.models.common import sinosoidal_position_embedding
from rho_diffusion.registry import registry


def _ops():
    from rho_diffusion.b200 import ops
    return ops


class TimestepBlock(nn.Module):
    """Any module whose forward takes the timestep embedding as a second argument (unet_v2.py:76-85)."""

    def forward(self, x, emb):
        raise NotImplementedError


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Sequential that routes ``emb`` to the children that take it (unet_v2.py:88-100)."""

    def forward_cl(self, h, emb, ctx: ExecCtx):
        for layer in self:
            if isinstance(layer, TimestepBlock):
                h = layer.forward_cl(h, emb, ctx)
            elif isinstance(layer, (Upsample, Downsample, AttentionBlock)):
                h = layer.forward_cl(h, ctx)
            else:  # plain conv (input_blocks[0])
                h = layer.forward_cl(h, ctx=ctx)
        return h

    def forward(self, x, emb):
        ops = _ops()
        ctx = ExecCtx(dtype=getattr(self, "_rho_dtype", torch.float32))
        h = self.forward_cl(ops.to_channels_last(x, ctx.dtype), emb, ctx)
        return ops.to_channels_first(h, x.dtype)


class Upsample(nn.Module):
    """Nearest x2 upsampling (3-D: inner two dims only) followed by an optional 3^n conv (unet_v2.py:103-134);
    the upsample is a gather inside the conv kernel, the 4x larger tensor is never written."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=1)

    def forward_cl(self, h, ctx: ExecCtx):
        assert h.shape[-1] == self.channels
        if not self.use_conv:
            return _ops().upsample_nearest(h, (1, 2, 2) if self.dims == 3 else (2,) * self.dims)
        up = (1, 2, 2) if self.dims == 3 else 2
        return self.conv.forward_cl(h, up=up, ctx=ctx)

    def forward(self, x):
        ops = _ops()
        ctx = ExecCtx(dtype=getattr(self, "_rho_dtype", torch.float32))
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x, ctx.dtype), ctx), x.dtype)


class Downsample(nn.Module):
    """Stride-2 3^n conv (3-D: stride (1,2,2)) (unet_v2.py:137-169)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        stride = 2 if dims != 3 else (1, 2, 2)
        if use_conv:
            self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=stride, padding=1)
        else:
            assert self.channels == self.out_channels
            self.op = avg_pool_nd(dims, kernel_size=stride, stride=stride)

    def forward_cl(self, h, ctx: ExecCtx):
        assert h.shape[-1] == self.channels
        return self.op.forward_cl(h, ctx=ctx)

    def forward(self, x):
        ops = _ops()
        ctx = ExecCtx(dtype=getattr(self, "_rho_dtype", torch.float32))
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x, ctx.dtype), ctx), x.dtype)


class ResBlock(TimestepBlock):
    """GN-act-conv -> (+emb | FiLM) -> GN-act-dropout-conv(zero-init) + skip (unet_v2.py:172-293)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False, activation=nn.SiLU()):
        super().__init__()
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.updown = up or down
        self._act = _ops_act(activation)

        self.in_layers = nn.Sequential(normalization(channels), activation, conv_nd(dims, channels, self.out_channels, 3, padding=1))
        if up:
            self.h_upd = Upsample(channels, False, dims)
            self.x_upd = Upsample(channels, False, dims)
        elif down:
            self.h_upd = Downsample(channels, False, dims)
            self.x_upd = Downsample(channels, False, dims)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(
            activation, nn.Linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm else self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels), activation, nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward_cl(self, h, emb, ctx: ExecCtx):
        """``h`` is a channels-last tensor or a pair (h, skip) standing for their channel concat."""
        ops = _ops()
        x, x2 = h if isinstance(h, tuple) else (h, None)
        gn1, conv1 = self.in_layers[0], self.in_layers[2]
        (xg, x), (x2g, x2) = ops.fork(x), ops.fork(x2)          # both feed the main path AND the skip: gradients summed by our kernel
        a = gn1.forward_cl(xg, x2g, act=self._act, ctx=ctx)
        if self.updown:
            # unet_v2.py:263-268: resample the activated features AND the skip input, then convolve
            assert x2 is None
            a = self.h_upd.forward_cl(a, ctx)
            x = self.x_upd.forward_cl(x, ctx)
        hh = conv1.forward_cl(a, ctx=ctx)
        lin = self.emb_layers[1]
        e = ops.linear(emb, lin.weight, lin.bias, None, act_in=self._act)
        gn2, conv2 = self.out_layers[0], self.out_layers[3]
        if self.use_scale_shift_norm:
            a2 = gn2.forward_cl(hh, film=e, act=self._act, ctx=ctx)
        else:
            a2 = gn2.forward_cl(hh, pre_add=e, act=self._act, ctx=ctx)
        if self.training and self.dropout > 0:
            a2 = ops.dropout(a2, self.dropout, seed=getattr(self, "_rho_dropout_seed", 0))      # unet_v2.py:239
        if isinstance(self.skip_connection, nn.Identity):
            assert x2 is None
            res = x
        else:
            res = self.skip_connection.forward_cl(x, x2, ctx=ctx)
        return conv2.forward_cl(a2, residual=res, ctx=ctx)

    def forward(self, x, emb):
        ops = _ops()
        ctx = ExecCtx(dtype=getattr(self, "_rho_dtype", torch.float32))
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x, ctx.dtype), emb, ctx), x.dtype)


def _ops_act(activation) -> int:
    return _ops().act_code(activation)


class AttentionBlock(nn.Module):
    """x + proj_out(attention(qkv(GN(x)))) over the flattened spatial positions (unet_v2.py:296-342)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False, use_new_attention_order=False):
        super().__init__()
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0, (
                f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}")
            self.num_heads = channels // num_head_channels
        self.use_checkpoint = use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        # split order: new = qkv chunks first then heads (QKVAttention), legacy = heads first (QKVAttentionLegacy)
        self.attention = QKVAttention(self.num_heads) if use_new_attention_order else QKVAttentionLegacy(self.num_heads)
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))

    def forward_cl(self, h, ctx: ExecCtx):
        shape = h.shape
        xf = _ops().carry_stats(h.reshape(shape[0], -1, shape[-1]), h)            # (N, T, C) view
        xn, xf = _ops().fork(xf)                           # GroupNorm input and residual
        a = self.norm.forward_cl(xn, ctx=ctx)
        qkv = self.qkv.forward_cl(a, ctx=ctx)              # (N, T, 3C)
        o = self.attention.forward_cl(qkv)
        out = self.proj_out.forward_cl(o, residual=xf, ctx=ctx)
        return _ops().carry_stats(out.reshape(shape), out)

    def forward(self, x):
        ops = _ops()
        ctx = ExecCtx(dtype=getattr(self, "_rho_dtype", torch.float32))
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x, ctx.dtype), ctx), x.dtype)


class _QKVBase(nn.Module):
    order = 0

    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads

    def forward_cl(self, qkv):
        """qkv: channels-last (N, T, 3*H*ch) -> (N, T, H*ch)."""
        assert qkv.shape[-1] % (3 * self.n_heads) == 0
        return _ops().attention(qkv, self.n_heads, self.order)

    def forward(self, qkv):
        """Reference layout: (N, 3*H*ch, T) -> (N, H*ch, T)."""
        ops = _ops()
        dtype = getattr(self, "_rho_dtype", torch.float32)
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(qkv, dtype)), qkv.dtype)


class QKVAttentionLegacy(_QKVBase):
    """Heads are split before q/k/v (unet_v2.py:365-397)."""
    order = 0


class QKVAttention(_QKVBase):
    """q/k/v are split before heads (unet_v2.py:400-436)."""
    order = 1


@registry.register_model("UNetv2")
class UNet(nn.Module):
    """The full UNet with attention and timestep embedding; arguments as in unet_v2.py:471-494."""

    def __init__(self, data_shape, in_channels: int, model_channels: int, out_channels: int, num_res_blocks: int,
                 attention_resolutions=[16, 8], dropout: float = 0, channel_mult=(1, 2, 4, 8), conv_resample: bool = True,
                 dims: int = 2, num_classes=None, cond_fn=None, use_checkpoint: bool = False, use_fp16: bool = False,
                 num_heads: int = 1, num_head_channels: int = -1, num_heads_upsample: int = -1,
                 use_scale_shift_norm: bool = False, resblock_updown: bool = False, use_new_attention_order: bool = False,
                 activation=nn.SiLU()):
        super().__init__()
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.data_shape = data_shape
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32   # reference attribute (unet_v2.py:511)
        self._rho_dtype = torch.bfloat16 if use_fp16 else torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.dims = dims

        emb_dim = model_channels * 4
        if isinstance(activation, str):
            activation = registry.get("activations", activation)()
        self._act = _ops_act(activation)
        self.time_embed = nn.Sequential(nn.Linear(model_channels, emb_dim), activation, nn.Linear(emb_dim, emb_dim))
        self.cond_fn = cond_fn
        if self.num_classes is not None:
            self.label_emb = None

        def res(ch_in, ch_out, **updown):
            return ResBlock(ch_in, emb_dim, dropout, out_channels=ch_out, dims=dims, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, activation=activation, **updown)

        def attn(ch, heads):
            return AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=heads, num_head_channels=num_head_channels,
                                  use_new_attention_order=use_new_attention_order)

        ch = input_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, ch, 3, padding=1))])
        self._feature_size = ch
        skip_chans = [ch]
        ds = 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, int(mult * model_channels))]
                ch = int(mult * model_channels)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(
                    res(ch, ch, down=True) if resblock_updown else Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
                self._feature_size += ch

        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch, num_heads), res(ch, ch))
        self._feature_size += ch

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = skip_chans.pop()
                layers = [res(ch + ich, int(model_channels * mult))]
                ch = int(model_channels * mult)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(res(ch, ch, up=True) if resblock_updown else Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch

        self.out = nn.Sequential(normalization(ch), activation, zero_module(conv_nd(dims, input_ch, out_channels, 3, padding=1)))

    # -- execution ------------------------------------------------------------------------------------
    def _embedding(self, timesteps, y):
        ops = _ops()
        te0, te2 = self.time_embed[0], self.time_embed[2]
        pe = sinosoidal_position_embedding(timesteps, self.model_channels)
        h = ops.linear(pe, te0.weight, te0.bias)
        cond = None
        if self.num_classes is not None:
            emb_dim = te2.weight.shape[0]
            if y.dim() == 1:
                assert y.shape == (timesteps.shape[0],)
                cond = self.cond_fn(y)
            elif y.dim() == 2:
                if tuple(y.shape) == (timesteps.shape[0], emb_dim):
                    cond = y.to(h.device)          # labels already are embeddings (unet_v2.py:710-715)
                else:
                    assert y.shape[0] == timesteps.shape[0]
                    cond = self.cond_fn(y)         # multi-categorical labels (unet_v2.py:716-719)
        return ops.linear(h, te2.weight, te2.bias, add=cond, act_in=self._act)

    def forward(self, x, timesteps, y=None):
        """x: (N, C, *spatial); timesteps: (N,) integer; y: labels iff the model is class-conditional."""
        assert (y is not None) == (self.num_classes is not None), (
            "must specify y if and only if the model is class-conditional")
        ops = _ops()
        ctx = ExecCtx(dtype=self._rho_dtype, engine=getattr(self, "_rho_engine", 0))
        emb = self._embedding(timesteps, y)
        n_blocks = len(self.input_blocks) + 1 + len(self.output_blocks)
        embs = list(ops.fork(emb, n_blocks))               # one alias per consumer block (gradients summed by our add kernel)
        h = ops.to_channels_last(x, ctx.dtype)
        hs = []
        for module in self.input_blocks:
            h = module.forward_cl(h, embs.pop(), ctx)
            h, skip = ops.fork(h)                          # next block + long skip
            hs.append(skip)
        h = self.middle_block.forward_cl(h, embs.pop(), ctx)
        for module in self.output_blocks:
            h = module.forward_cl((h, hs.pop()), embs.pop(), ctx)
        a = self.out[0].forward_cl(h, act=self._act, ctx=ctx)
        o = self.out[2].forward_cl(a, y_f32=True, ctx=ctx)   # eps-hat leaves the network in fp32
        return ops.to_channels_first(o, x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32)
