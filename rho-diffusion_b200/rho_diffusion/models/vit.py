"""Diffusion vision transformer (``"VisionTransformer"``) on the B200 kernels.

Same constructors, registry name and ``state_dict`` layout as the reference (``rho_diffusion/models/vit.py:31-372``): a
kernel = stride = patch convolution embeds 1-D / 2-D / 3-D data into a token sequence, a sinusoidal position embedding is added,
``transformer_depth`` pre-norm blocks (diffusion-time embedding -> LayerNorm -> ``nn.MultiheadAttention`` -> LayerNorm -> MLP) follow,
and a Linear plus a kernel = stride ``ConvTranspose`` map the tokens back to the input shape.  Execution:

* tokens are channels-last ``[N, T, E]`` (bf16 or fp32); every Linear over tokens (patch embedding, q/k/v and output projections,
  the MLP, the output projection and the transposed conv's matrix) is a 1x1 ``conv_nd`` on the implicit-GEMM engines with bias
  and residual fused in its epilogue; the patch convolution is a permutation (``rho_patchify``) followed by such a Linear;
* attention is the flash-style kernel of the UNet (``in_proj`` already produces the q | k | v, then heads layout of ``QKVAttention``),
  so the ``T x T`` weights never touch HBM -- ``attn_weights`` in the block's result is therefore ``None`` (the reference returns the
  head-averaged weights, which ``VisionTransformer.forward`` ignores);
* LayerNorm (with the time-embedding add of ``data + t_embed`` fused in) is one row-wise kernel forward and three backward.

``dropout > 0`` (attention-probability dropout inside ``nn.MultiheadAttention`` and the two MLP dropouts) has no kernel: training
with it raises; in ``eval()`` mode dropout is the identity as in the reference.
"""
from __future__ import annotations

from collections.abc import Mapping

import torch
from torch import nn, Tensor

from rho_diffusion.layers import ExecCtx, module_ctx
from rho_diffusion.models.common import SinusoidalPositionEmbedding
from rho_diffusion.registry import registry

__all__ = ["PatchEmbedding", "AttentionBlock", "VisionTransformer"]


def _ops():
    from rho_diffusion.b200 import ops
    return ops


class PatchEmbedding(nn.Module):
    """Embeds an image / volume / series into a sequence of patches (vit.py:31-126)."""

    def __init__(self, num_channels: int, patch_size: int, embedding_dim: int, data_dims: int) -> None:
        super().__init__()
        self.num_channels = num_channels
        self.patch_size = patch_size
        self.embedding_dim = embedding_dim
        self.data_dims = data_dims
        self._conv_type = registry.get("nn", f"Conv{self.data_dims}d")
        self.conv_shaper = self._conv_type(num_channels, embedding_dim, kernel_size=patch_size, stride=patch_size)

    @property
    def data_dims(self) -> int:
        return self._data_dims

    @data_dims.setter
    def data_dims(self, value: int) -> None:
        assert 3 >= value > 0, "data_dims must be between 1 and 3 for convolution."
        self._data_dims = value

    @property
    def input_shape(self) -> str:
        return {3: "n l h w d", 2: "n l h w", 1: "n l h"}[self.data_dims]

    @property
    def output_shape(self) -> str:
        return {3: "n (h w d) l", 2: "n (h w) l", 1: "n h l"}[self.data_dims]

    @property
    def ein_string(self) -> str:
        return f"{self.input_shape} -> {self.output_shape}"

    @property
    def stored_shape(self):
        """Shape of the patch grid (conv output without batch and channels)."""
        return self._stored_shape

    @stored_shape.setter
    def stored_shape(self, value) -> None:
        self._stored_shape = value

    def forward_cl(self, data: Tensor, ctx: ExecCtx) -> Tensor:
        """(N, C, *spatial) -> channels-last tokens (N, T, E)."""
        ops = _ops()
        p = (self.patch_size,) * self.data_dims
        self.stored_shape = tuple(s // self.patch_size for s in data.shape[2:])
        tok = ops.patchify(data, p, ctx.dtype)
        w = self.conv_shaper.weight
        return ops.linear_tokens(self.conv_shaper, tok, w.reshape(w.shape[0], -1), self.conv_shaper.bias, dtype=ctx.dtype, engine=ctx.engine)

    def forward(self, data: Tensor) -> Tensor:
        ops = _ops()
        out = self.forward_cl(data, module_ctx(self))
        return out if out.dtype == data.dtype else ops.cast(out, data.dtype)


class AttentionBlock(nn.Module):
    """Time-conditioned transformer block (vit.py:129-189)."""

    def __init__(self, embed_dim: int, hidden_dim: int, num_heads: int, dropout: float = 0.0, activation="GELU", time_dim: int = 128,
                 **attn_kwargs):
        super().__init__()
        self.norm_1 = nn.LayerNorm(embed_dim)
        self.norm_2 = nn.LayerNorm(embed_dim)
        attn_kwargs.setdefault("batch_first", True)
        self.attention_layer = nn.MultiheadAttention(embed_dim, num_heads, dropout, **attn_kwargs)
        mha = self.attention_layer
        if not mha.batch_first or not mha._qkv_same_embed_dim or mha.bias_k is not None or mha.add_zero_attn:
            raise NotImplementedError("VisionTransformer on B200 supports nn.MultiheadAttention(batch_first=True) self-attention without "
                                      "kdim / vdim / add_bias_kv / add_zero_attn")
        activation = registry.get("activations", activation) if isinstance(activation, str) else type(activation)
        self.linear_block = nn.Sequential(nn.Linear(embed_dim, hidden_dim), activation(), nn.Dropout(dropout), nn.Linear(hidden_dim, embed_dim),
                                          nn.Dropout(dropout))
        self.time_transform = nn.Sequential(SinusoidalPositionEmbedding(time_dim), nn.Linear(time_dim, embed_dim, bias=False), activation())
        self.dropout = dropout
        self._act = _ops().act_code(self.linear_block[1])

    def forward_cl(self, data: Tensor, t: Tensor, ctx: ExecCtx) -> Tensor:
        ops = _ops()
        if self.training and self.dropout > 0:
            raise NotImplementedError("VisionTransformer on B200: dropout > 0 has no kernel in training mode (attention-probability dropout); "
                                      "build the model with dropout=0.0 or call eval()")
        tt = self.time_transform
        t_embed = ops.activation(ops.linear(tt[0](t), tt[1].weight, None), self._act)                   # [N, E] fp32
        norm_data = ops.layernorm(data, self.norm_1.weight, self.norm_1.bias, self.norm_1.eps, pre_add=t_embed)   # norm_1(data + t_embed)
        nq, nres = ops.fork(norm_data)
        mha = self.attention_layer
        qkv = ops.linear_tokens(mha, nq, mha.in_proj_weight, mha.in_proj_bias, dtype=ctx.dtype, engine=ctx.engine)
        o = ops.attention(qkv, mha.num_heads, 1)                                                          # q | k | v first, then heads
        attn_residual = ops.linear_tokens(mha.out_proj, o, mha.out_proj.weight, mha.out_proj.bias, residual=nres, dtype=ctx.dtype,
                                          engine=ctx.engine)                                              # norm_data + attn_o
        ar, ares = ops.fork(attn_residual)
        n2 = ops.layernorm(ar, self.norm_2.weight, self.norm_2.bias, self.norm_2.eps)
        l1, l2 = self.linear_block[0], self.linear_block[3]
        m = ops.activation(ops.linear_tokens(l1, n2, l1.weight, l1.bias, dtype=ctx.dtype, engine=ctx.engine), self._act)
        return ops.linear_tokens(l2, m, l2.weight, l2.bias, residual=ares, dtype=ctx.dtype, engine=ctx.engine)

    def forward(self, data: Tensor, t: Tensor) -> dict:
        ops = _ops()
        ctx = module_ctx(self)
        x = data if data.dtype == ctx.dtype else ops.cast(data, ctx.dtype)
        out = self.forward_cl(x, t, ctx)
        return {"output": out if out.dtype == data.dtype else ops.cast(out, data.dtype), "attn_weights": None}


@registry.register_model("VisionTransformer")
class VisionTransformer(nn.Module):
    """Vision transformer specialised for diffusion (vit.py:192-372): ``forward(input_data, t)``."""

    def __init__(self, patch_size: int, input_shapes, num_channels: int, embedding_dim: int, hidden_dim: int, activation,
                 transformer_depth: int = 8, pos_embedding_dim: int = 128, time_embedding_dim: int = 128, max_seq_length: int = 20_000,
                 dropout: float = 0.2, num_heads: int = 16, attention_kwargs: Mapping | None = {}) -> None:
        super().__init__()
        data_dims = len(input_shapes)
        self.input_shapes = input_shapes
        self.patch_embedder = PatchEmbedding(num_channels, patch_size, embedding_dim, data_dims)
        self.max_seq_length = max_seq_length
        self.transformer_blocks = nn.ModuleList([
            AttentionBlock(embedding_dim, hidden_dim, num_heads, dropout, activation, time_embedding_dim, **(attention_kwargs or {}))
            for _ in range(transformer_depth)])
        self.output_projection = nn.Linear(embedding_dim, hidden_dim, bias=False)
        conv_type = registry.get("nn", f"ConvTranspose{self.data_dims}d")
        self.output_conv = conv_type(hidden_dim, num_channels, kernel_size=self.patch_size, stride=self.patch_size)
        act_cls = registry.get("activations", activation) if isinstance(activation, str) else type(activation)
        self.pos_embedding = nn.Sequential(SinusoidalPositionEmbedding(pos_embedding_dim), nn.Linear(pos_embedding_dim, embedding_dim), act_cls())
        self._act = _ops().act_code(self.pos_embedding[2])

    @property
    def input_shape_dim_names(self):
        return [key for _, key in zip(self.input_shapes, ["j", "k", "l"])]

    @property
    def patch_size(self) -> int:
        return self.patch_embedder.patch_size

    @property
    def data_dims(self) -> int:
        return self.patch_embedder.data_dims

    @property
    def stored_shape(self):
        return self.patch_embedder.stored_shape

    @property
    def stored_shape_ein(self):
        return {key: value for key, value in zip(["j", "k", "l"], self.stored_shape)}

    def forward(self, input_data: Tensor, t: Tensor) -> Tensor:
        ops = _ops()
        ctx = module_ctx(self)
        h = self.patch_embedder.forward_cl(input_data, ctx)                                              # [N, T, E]
        pe = self.pos_embedding
        idx = torch.arange(h.shape[1], device=h.device)
        pos = ops.activation(ops.linear(pe[0](idx), pe[1].weight, pe[1].bias), self._act)               # [T, E] fp32
        h = ops.add_bcast(h, pos)
        for block in self.transformer_blocks:
            h = block.forward_cl(h, t, ctx)
        h = ops.linear_tokens(self.output_projection, h, self.output_projection.weight, None, dtype=ctx.dtype, engine=ctx.engine)
        # kernel = stride ConvTranspose: tokens @ W.view(hidden, C p^d) + bias repeated per patch element, then the inverse permutation
        oc = self.output_conv
        reps = self.patch_size ** self.data_dims
        w = ops.ConvTWeightFn.apply(oc.weight.reshape(oc.weight.shape[0], -1))                          # [C p^d, hidden]
        b = ops.BiasExpandFn.apply(oc.bias, reps) if oc.bias is not None else None
        y = ops.linear_tokens(None, h, w, b, dtype=ctx.dtype, engine=ctx.engine, y_f32=True, cached=False)
        grid = self.stored_shape
        field = (input_data.shape[0], oc.out_channels, *[g * self.patch_size for g in grid])
        out_dtype = input_data.dtype if input_data.dtype in (torch.float32, torch.bfloat16) else torch.float32
        return ops.unpatchify(y, (self.patch_size,) * self.data_dims, field, out_dtype)
