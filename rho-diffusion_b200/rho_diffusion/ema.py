"""Exponential moving average of model weights (reference ``rho_diffusion/ema.py:29-79``)."""
from __future__ import annotations

import copy
import math
from collections import OrderedDict

import torch
from torch import nn


class ExponentialMovingAverage(nn.Module):
    """Shadow model with the warm-up decay ``d_t = decay * (1 - exp(-step / 2000))``;
    ``update()``: ``shadow -= (1 - d_t) * (shadow - param)``, buffers copied.

    When the model was prepared by ``rho_diffusion.b200.optimize`` the EMA is folded into the fused AdamW
    kernel (one pass over flat buffers); this class is the stand-alone form with one launch per tensor.
    """

    def __init__(self, model, decay=0.9999) -> None:
        super().__init__()
        self.model = model
        self.ema_model = copy.deepcopy(model).eval()
        self.decay = decay
        self.decay_func = lambda x: decay * (1 - math.exp(-x / 2000))
        self.step_id = 0
        self.current_ema_frac = 0.0
        for p in self.ema_model.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def update(self):
        from rho_diffusion.b200 import ops

        self.step_id += 1
        frac = self.current_ema_frac = self.decay_func(self.step_id)
        params = OrderedDict(self.model.named_parameters())
        shadow = OrderedDict(self.ema_model.named_parameters())
        assert params.keys() == shadow.keys()
        for name, p in params.items():
            ops.ema_update(shadow[name], p.detach(), 1.0 - frac)
        buffers = OrderedDict(self.model.named_buffers())
        shadow_buffers = OrderedDict(self.ema_model.named_buffers())
        assert buffers.keys() == shadow_buffers.keys()
        for name, b in buffers.items():
            shadow_buffers[name].copy_(b)

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)

    def denoise_process(self, *args, **kwargs):
        return self.ema_model.denoise_process(*args, **kwargs)
