"""TEST INFRASTRUCTURE -- the named UNet / DDPM parity cases shared by the golden generator and tests.

Each case is (UNet kwargs, batch, conditioning kind).  Shapes are small enough for the CPU oracle to
finish in seconds; together they cover n = 1, 2, 3, both attention split orders, FiLM and additive
embedding, all three conditioning paths (none / same-shape embedding / MultiEmbeddings), odd channel
counts (3) and the reference's own test fixture shape (tests/models/test_unet.py:28-56).
"""
from __future__ import annotations

UNET_CASES = {
    # the reference's own smoke-test fixture: unconditional 2-D, heads=1, additive embedding, legacy order
    "ref2d": dict(
        kwargs=dict(data_shape=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=2),
        x_shape=(2, 3, 24, 16), cond=None,
    ),
    # 3-D FiLM, 4 heads, new attention order, sha-style same-shape label embedding (config A family)
    "film3d": dict(
        kwargs=dict(data_shape=[8, 16, 16], in_channels=1, model_channels=32, out_channels=1, num_res_blocks=1,
                    attention_resolutions=[2, 4], channel_mult=[1, 2, 4], dims=3, num_classes=20, num_heads=4,
                    use_scale_shift_norm=True, use_new_attention_order=True, activation="SiLU"),
        x_shape=(2, 1, 8, 16, 16), cond="same_shape",
    ),
    # 1-D FiLM, 2 heads, legacy order, MultiEmbeddings conditioning (config B/C family)
    "multi1d": dict(
        kwargs=dict(data_shape=[64], in_channels=1, model_channels=32, out_channels=1, num_res_blocks=1,
                    attention_resolutions=[4, 8], channel_mult=[1, 2, 4, 4], dims=1, num_classes=25, num_heads=2,
                    use_scale_shift_norm=True, use_new_attention_order=False, activation="SiLU"),
        x_shape=(3, 1, 64), cond="multi",
        parameter_space={"s": [0.25, 0.5, 1.0], "m": [1.0], "t": [400.0, 405.0, 410.0, 415.0]},
    ),
    # 2-D deep-galaxy family (config B) at reduced resolution: FiLM, 4 heads, 64-multiple channels at depth
    "galaxy2d": dict(
        kwargs=dict(data_shape=[32, 32], in_channels=1, model_channels=32, out_channels=1, num_res_blocks=2,
                    attention_resolutions=[16, 8], dims=2, num_classes=25, num_heads=4,
                    use_scale_shift_norm=True, use_new_attention_order=True, activation="SiLU"),
        x_shape=(2, 1, 32, 32), cond="multi",
        parameter_space={"s": [0.25, 0.5, 1.0], "m": [1.0], "t": [400.0, 405.0, 410.0], "c": [0.0]},
    ),
    # the optional structure of the reference UNet: resblock_updown=True (ResBlocks that pool / upsample, unet_v2.py:218-227,
    # 263-268) with the registered SymmetricLog activation (activations.py:35-61), 3-D (depth is kept)
    "updown3d": dict(
        kwargs=dict(data_shape=[4, 16, 16], in_channels=1, model_channels=32, out_channels=1, num_res_blocks=1,
                    attention_resolutions=[2], channel_mult=[1, 2, 2], dims=3, num_classes=20, num_heads=2,
                    use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=True, activation="SymmetricLog"),
        x_shape=(2, 1, 4, 16, 16), cond="same_shape",
    ),
    # conv_resample=False: average-pool Downsample / conv-less nearest Upsample (unet_v2.py:124-131, 163-165), additive
    # embedding, GELU, 2-D with 3 input channels
    "pool2d": dict(
        kwargs=dict(data_shape=[12, 16], in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1,
                    attention_resolutions=[2], channel_mult=[1, 2], conv_resample=False, dims=2, num_heads=2,
                    use_scale_shift_norm=False, activation="GELU"),
        x_shape=(2, 3, 12, 16), cond=None,
    ),
}

TIMESTEPS = {"ref2d": [0, 7], "film3d": [3, 999], "multi1d": [1, 500, 250], "galaxy2d": [17, 420], "updown3d": [5, 640],
             "pool2d": [11, 902]}


def cond_values(case: str):
    """Label rows for the 'multi' cases (rows of the parameter space)."""
    if case == "multi1d":
        return [[0.5, 1.0, 405.0], [1.0, 1.0, 400.0], [0.25, 1.0, 415.0]]
    if case == "galaxy2d":
        return [[0.25, 1.0, 410.0, 0.0], [1.0, 1.0, 400.0, 0.0]]
    return None


def loss_curve_inputs(step: int, shape, T: int, n_data: int = 8):
    """Deterministic (data batch index, t, noise) of training step ``step`` of the 1k-step loss-curve fixture -- shared by
    the golden generator (oracle/make_golden.py::gen_loss_curve) and tests/test_gpu_fullsize.py.  Uses the CPU generator
    through functions captured at import (the golden generator patches ``torch.randint`` / ``torch.randn_like``)."""
    g = _torch.Generator().manual_seed(100003 * (step + 1))
    t = _RANDINT(0, T, (shape[0],), generator=g)
    noise = _RANDN(shape, generator=g)
    return step % n_data, t, noise


import torch as _torch  # noqa: E402

_RANDINT, _RANDN = _torch.randint, _torch.randn


# ---- the two remaining registered backbones ---------------------------------------------------------------------
# v1 UNet (models/unet.py): stride-1 blocks, ConvTranspose up path, GroupNorm(8); VisionTransformer (models/vit.py): patch
# embedding, LayerNorm, nn.MultiheadAttention, MLP, kernel = stride ConvTranspose.  "model" is the registry name.
BACKBONE_CASES = {
    # 2-D, ReLU (the constructor default), residual branches (ConvTranspose on the up path)
    "unet1_2d": dict(model="UNet", kwargs=dict(block_type="UNetBlock2d", input_channels=3, down_channels=[32, 64, 128],
                                               up_channels=[128, 64, 32], time_embedding_dim=32, activation="ReLU", residual=True),
                     x_shape=(2, 3, 16, 24), t=[3, 700]),
    # 2-D, GELU, no residual branch, wider time embedding, 64-multiple channels (tensor-core shapes in bf16)
    "unet1_2d_gelu": dict(model="UNet", kwargs=dict(block_type="UNetBlock2d", input_channels=1, down_channels=[64, 128],
                                                    up_channels=[128, 64], time_embedding_dim=64, activation="GELU", residual=False),
                          x_shape=(2, 1, 32, 32), t=[0, 999]),
    # 3-D blocks: N == down_channels[k] == depth would be needed for the reference's `time_pe[..., None, None]` add to broadcast
    # (unet.py:130-132); the generator records whether the reference runs at all
    "unet1_3d": dict(model="UNet", kwargs=dict(block_type="UNetBlock3d", input_channels=1, down_channels=[32, 64],
                                               up_channels=[64, 32], time_embedding_dim=32, activation="SiLU", residual=True),
                     x_shape=(2, 1, 4, 8, 8), t=[5, 250]),
    # ViT 2-D: 4x4 patches of 3-channel 16x16 images, 4 heads of 16
    "vit_2d": dict(model="VisionTransformer", kwargs=dict(patch_size=4, input_shapes=[16, 16], num_channels=3, embedding_dim=64, hidden_dim=128,
                                                          activation="GELU", transformer_depth=2, pos_embedding_dim=32, time_embedding_dim=32,
                                                          dropout=0.0, num_heads=4),
                   x_shape=(2, 3, 16, 16), t=[3, 700]),
    # ViT 3-D: 2x2x2 patches, one head of 64 (the tcgen05 attention shape in bf16), SiLU
    "vit_3d": dict(model="VisionTransformer", kwargs=dict(patch_size=2, input_shapes=[4, 8, 8], num_channels=1, embedding_dim=64, hidden_dim=64,
                                                          activation="SiLU", transformer_depth=1, pos_embedding_dim=64, time_embedding_dim=64,
                                                          dropout=0.0, num_heads=1),
                   x_shape=(2, 1, 4, 8, 8), t=[11, 420]),
    # ViT 1-D: patches of 8 over a series of 128, 2 heads of 32
    "vit_1d": dict(model="VisionTransformer", kwargs=dict(patch_size=8, input_shapes=[128], num_channels=2, embedding_dim=64, hidden_dim=96,
                                                          activation="GELU", transformer_depth=2, pos_embedding_dim=32, time_embedding_dim=48,
                                                          dropout=0.0, num_heads=2),
                   x_shape=(3, 2, 128), t=[1, 500, 999]),
}


def spectroscopy_records():
    """Three synthetic SpectroscopyDataset records (spectroscopy.py:103-109: group str(index) -> 'transitions' [2, n] = centre
    frequencies / log10 intensities, plus the parameters the spectrum was generated from), as '<index>/<name>' arrays."""
    import numpy as np
    rng = np.random.RandomState(20240607)
    rec = {}
    for i, n in enumerate((40, 17, 1)):
        centers = rng.uniform(500.0, 34000.0, size=n)            # some fall outside the default 1000 .. 32000 grid
        logint = rng.uniform(-12.0, -1.0, size=n)                # some beyond the [-10, -2] clip
        if n == 1:
            centers[0] = 16000.0
        rec[f"{i}/transitions"] = np.stack([centers, logint]).astype(np.float64)
        rec[f"{i}/A"] = np.array([3000.0 + 10 * i])
        rec[f"{i}/B"] = np.array([2000.0 + 5 * i])
        rec[f"{i}/C"] = np.array([1500.0 + i])
    return rec
