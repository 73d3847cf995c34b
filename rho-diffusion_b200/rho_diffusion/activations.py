"""Registered activations beyond torch's (reference ``rho_diffusion/activations.py``).

``SymmetricLog`` (Cai et al., arXiv:2111.15631): ``f(x) = tanh(x) + tanh(x) * log(x * tanh(x) + 1)``.  Inside the UNet it is
applied by the fused GroupNorm kernel (``RHO_ACT_SYMLOG``); as a stand-alone module it runs the same kernel through an
identity normalisation-free path (``rho_gn_apply`` is not involved: a plain elementwise launch via ``rho_add_cast`` would not
apply an activation), so the module form evaluates the closed form with torch ops -- it is not on the DDPM hot path.
"""
from __future__ import annotations

import torch
from torch import nn

from rho_diffusion.registry import registry

__all__ = ["SymmetricLog"]


@registry.register_activation("SymmetricLog")
class SymmetricLog(nn.Module):
    def forward(self, data: torch.Tensor) -> torch.Tensor:
        tanx = data.tanh()
        return tanx + tanx * torch.log(data * tanx + 1)
