"""TEST INFRASTRUCTURE -- deterministic synthetic weights / inputs shared by the golden generator,
the parity tests and the bench.

A fresh reference UNet outputs exactly 0 (the ResBlock out conv, attention ``proj_out`` and the
final conv are zero-initialised, layers.py:113-119), so every comparison against it would be
vacuous.  Instead of relying on torch's module-construction RNG order, weights are a pure function
of (state-dict key, shape, seed): both the reference run that produced ``tests/golden/*`` and the
tests that load the same weights into the CUDA model call :func:`synth_state_dict`.
"""
from __future__ import annotations

import math
import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**31 - 1))


def synth_tensor(key: str, shape, seed: int = 777) -> torch.Tensor:
    g = _gen(seed, key)
    shape = tuple(shape)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r * (1.0 / math.sqrt(fan_in))
    if key.endswith("weight"):  # GroupNorm gain
        return 1.0 + 0.1 * r
    return 0.1 * r  # biases


def synth_state_dict(manifest: dict, seed: int = 777) -> dict:
    """``manifest`` maps state-dict key -> shape (e.g. ``{k: tuple(v.shape) for k, v in sd.items()}``)."""
    return {k: synth_tensor(k, shp, seed) for k, shp in manifest.items()}


def synth_input(name: str, shape, seed: int = 777, kind: str = "randn") -> torch.Tensor:
    g = _gen(seed, "input:" + name)
    if kind == "randn":
        return torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if kind == "rand":
        return torch.rand(tuple(shape), generator=g, dtype=torch.float32)
    raise ValueError(kind)
