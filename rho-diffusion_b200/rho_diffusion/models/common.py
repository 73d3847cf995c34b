"""Sinusoidal timestep embedding (reference ``rho_diffusion/models/common.py:27-80``) on the B200 kernels."""
from __future__ import annotations

import torch
from torch import nn

from rho_diffusion.registry import registry


def omega_table(dim: int, wavelength: int = 10000) -> torch.Tensor:
    """omega_i = wavelength^(2i/dim), built on the host in float32 exactly as common.py:37-38 builds it."""
    assert dim % 2 == 0, "`dim` should be dividable by 2."
    i = torch.arange(dim // 2)
    return torch.pow(wavelength, 2 * i / dim)


_omega_cache = {}


def sinosoidal_position_embedding(t: torch.Tensor, dim: int, wavelength: int = 10000) -> torch.Tensor:
    """t -> [sin(t/w_0), cos(t/w_0), sin(t/w_1), cos(t/w_1), ...] (interleaved), float32 ``(len(t), dim)``."""
    from rho_diffusion.b200 import ops

    key = (dim, wavelength, t.device)
    if key not in _omega_cache:
        with torch.inference_mode(False):
            _omega_cache[key] = omega_table(dim, wavelength).to(t.device)
    return ops.timestep_embedding(t, dim, _omega_cache[key])


@registry.register_layer("SinusoidalPositionEmbedding")
class SinusoidalPositionEmbedding(nn.Module):
    """Module form of :func:`sinosoidal_position_embedding` (common.py:46-80)."""

    def __init__(self, dim: int, wavelength: int = 10000) -> None:
        super().__init__()
        self.dim = dim
        self.wavelength = wavelength

    def forward(self, t: torch.Tensor):
        return sinosoidal_position_embedding(t, self.dim, self.wavelength)
