"""The fallback paths behind the environment switches (README "Switches") stay parity-green: each switch is read once per process, so
every variant reruns a slice of the kernel parity suite (conv forward / dgrad / wgrad on the shapes of each engine, the fused
statistics test, attention, and a UNet golden) in a fresh interpreter with the switch set."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# conv cases: per-tap tcgen05 (14, 16), strided / upsample (19, 20, 21), halo 64 / 128 / concat (30, 31, 33), config-D pairs (36, 44),
# 32-channel halo (45, 46), 1x1 on 32-channel blocks (50, 51), wide 1x1 convs of the attention blocks (55, 56, 57, 58)
CONV_IDS = (14, 16, 19, 20, 21, 30, 31, 33, 36, 44, 45, 46, 50, 51) + (55, 56, 57, 58)
SELECT = "(test_conv_fwd_bwd and (" + " or ".join(f"[{i}-dtype1]" for i in CONV_IDS) + ")) " \
         "or test_conv_epilogue_groupnorm_statistics or (test_attention_fwd_bwd and dtype1) or test_conv_engines_agree_on_random_shapes"

VARIANTS = [
    {"RHO_CONV_HALO": "0"},
    {"RHO_WGRAD_HALO": "0"},
    {"RHO_CONV_TMA_STORE": "0"},
    {"RHO_FUSE_GN_STATS": "0"},
    {"RHO_ATTN_POLY": "2"},
    {"RHO_ATTN_SPLIT": "1"},
    {"RHO_ATTN_V2": "0"},               # first-generation attention forward (64-key tiles, P through shared memory) for every shape
    {"RHO_ATTN_V2_MIN_T": "1"},         # second-generation forward (P in TMEM, two query tiles per CTA) forced for every head-dim-64 shape
    {"RHO_ATTN_V2_MIN_T": "1", "RHO_ATTN_POLY": "0"},   # ... with every exponential on the MUFU
    {"RHO_ATTN_BWD_V2": "0"},           # first-generation attention backward (32-column tiles, P / dS through shared memory)
    {"RHO_ATTN_BWD_V2_MIN_T": "1"},     # second-generation backward (64-column tiles, P / dS in TMEM) forced for every head-dim-64 shape
    {"RHO_HALO_WIDE_K1": "1"},          # wide 1x1 convs (attention qkv / proj_out) through the halo kernel's streaming mode
    {"RHO_CONV_CLUSTER": "1"},          # CTA pairs sharing the weight tile through TMA multicast (per-tap kernel, BLOCK_N >= 128)
    {"RHO_CONV_DEEP_RING": "0"},        # small launches (<= 1 CTA per SM: every test shape) on the 96 KB ring of the full-size launches
    {"RHO_HALO_ISSUERS": "1", "RHO_WGRAD_HALO_ISSUERS": "1"},   # one MMA issuer thread in the halo conv / weight-gradient kernels
    {"RHO_FP32_SPLIT": "1"},            # fp32 inference convs as six bf16 tensor-core products with flushed fp32 accumulation
]


@pytest.mark.parametrize("env", VARIANTS, ids=[",".join(f"{k}={v}" for k, v in e.items()) for e in VARIANTS])
def test_switched_paths_keep_parity(env):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    e = dict(os.environ, **env)
    select = SELECT
    if "RHO_FUSE_GN_STATS" in env or "RHO_CONV_TMA_STORE" in env:
        # the statistics test asserts that the side channel exists: not with the fusion off, nor without the TMA-store epilogue of the
        # per-tap kernel, which is where its statistics are computed
        select = select.replace("or test_conv_epilogue_groupnorm_statistics ", "")
    if "RHO_CONV_HALO" in env:              # case9 expects statistics from a 32-channel conv: only the halo kernel fuses them at N = 32
        select = select.replace("test_conv_epilogue_groupnorm_statistics", "(test_conv_epilogue_groupnorm_statistics and not case9)")
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_ops.py"), os.path.join(ROOT, "tests", "test_gpu_model.py"),
           "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-k", select + " or test_unet_forward_bf16_within_tolerance"]
    if "RHO_FP32_SPLIT" in env:            # the fp32 goldens (<= 1e-5 against the unmodified reference) and the per-block taps on the split engine
        cmd[-1] = "test_unet_forward_fp32_matches_reference or test_per_block_taps_fp32 or test_split_bf16"
    r = subprocess.run(cmd, env=e, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
