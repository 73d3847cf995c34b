"""Base class of the diffusion pipelines (reference ``rho_diffusion/diffusion/abstract_diffusion.py``).

The reference derives from ``lightning.pytorch.LightningModule``; Lightning is not a dependency here -- the
pipeline is a plain ``nn.Module`` that provides the few Lightning attributes the reference code and the
training script use (``log``, ``hparams``, ``device``, ``global_rank``, ``current_epoch``).
"""
from __future__ import annotations

import math
import types
from collections.abc import Mapping
from inspect import getfullargspec
from logging import getLogger
from typing import Any, Union

import torch
from torch import Tensor, nn
from torch.optim import AdamW

from rho_diffusion.diffusion import schedule as _schedule
from rho_diffusion.registry import registry

__all__ = ["AbstractDiffusionPipeline"]


class AbstractDiffusionPipeline(nn.Module):
    def __init__(self, backbone, backbone_kwargs: dict, schedule, timesteps: Union[int, Tensor] = 1000, cond_fn=None,
                 cond_fn_kwargs: dict = None, optimizer=None, opt_kwargs: Union[Mapping[str, Any], None] = {}):
        super().__init__()
        if isinstance(backbone, str):
            backbone = registry.get("models", backbone)
        self.backbone = backbone(**backbone_kwargs)
        self.backbone_kwargs = backbone_kwargs
        if isinstance(cond_fn, str):
            self.backbone.cond_fn = registry.get("layers", cond_fn)(**(cond_fn_kwargs or {}))
        # the reference leaves the *string* "AdamW" here when optimizer is None and then crashes in
        # configure_optimizers (abstract_diffusion.py:72-76); the evident intent is the registry entry
        if optimizer is None:
            optimizer = "AdamW"
        if isinstance(optimizer, str):
            optimizer = registry.get("optimizers", optimizer)
        self.optimizer = optimizer
        self.schedule = schedule
        self.metrics = nn.ModuleDict()   # the reference's PSNR metric is only logged; see DDPM(log_psnr=...)
        self.hparams = types.SimpleNamespace(model_kwargs=backbone_kwargs, opt_kwargs=dict(opt_kwargs or {}))
        self.timesteps = timesteps
        self._global_rank = None          # None: follow torch.distributed (see the property below)
        self.current_epoch = 0
        self._python_logger = getLogger(self.__class__.__name__)

    # -- Lightning-compatible surface ------------------------------------------------------------------
    def log(self, *args, **kwargs) -> None:
        """No-op outside a Lightning Trainer (as in the reference when driven by scripts/training_ddp.py)."""

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def global_rank(self) -> int:
        """Lightning's ``global_rank``.  Without a Trainer it follows the process group (0 when none is initialised), so that
        every rank of a torchrun job keys its Philox noise differently (``noise``: seed + 7919 * rank) and the ranks of a
        batch-sharded ``generate`` do not produce duplicate samples; assigning to it pins a value."""
        if self._global_rank is not None:
            return self._global_rank
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
        return 0

    @global_rank.setter
    def global_rank(self, value) -> None:
        self._global_rank = None if value is None else int(value)

    # -- optimizer ---------------------------------------------------------------------------------------
    def configure_optimizers(self, mpi_world_size: int = 1) -> dict:
        """AdamW defaults filled from ``inspect.getfullargspec(AdamW)``, ``lr *= sqrt(world)`` and the optimizer
        built over ``self.parameters()`` (abstract_diffusion.py:86-148; the LR scheduler it builds is discarded there
        too).  Unlike the reference the stored ``hparams.opt_kwargs`` are not mutated, so calling this twice does not
        compound the learning-rate scaling."""
        opt_kwargs = dict(self.hparams.opt_kwargs)
        spec = getfullargspec(AdamW)
        names = [a for a in spec.args if a not in ("self", "params")]
        for key, value in zip(names, spec.defaults or ()):
            opt_kwargs.setdefault(key, value)
        if self.optimizer is not AdamW:
            allowed = set(getfullargspec(self.optimizer).args) | set(getfullargspec(self.optimizer).kwonlyargs)
            opt_kwargs = {k: v for k, v in opt_kwargs.items() if k in allowed}
        else:
            allowed = set(spec.args) | set(spec.kwonlyargs)
        opt_kwargs = {k: v for k, v in opt_kwargs.items() if k in allowed and k != "lr_schedule"}
        opt_kwargs["lr"] = opt_kwargs["lr"] * math.sqrt(mpi_world_size)
        return {"optimizer": self.optimizer(self.parameters(), **opt_kwargs)}

    # -- schedule helpers --------------------------------------------------------------------------------
    @property
    def schedule(self):
        return self._schedule

    @schedule.setter
    def schedule(self, value) -> None:
        self._schedule = value

    def random_timesteps(self, num_steps: int) -> Tensor:
        """``randint(0, T)`` drawn from the **CPU** generator, like abstract_diffusion.py:163-169.  Under torchrun the
        training script reseeds that generator with ``seed + rank`` AFTER the model is built (identical weight init, different
        timesteps per rank): ``reseed_per_rank``."""
        return torch.randint(low=0, high=self.timesteps, size=(num_steps,))

    def reseed_per_rank(self, seed: int) -> int:
        """Call once after construction in a multi-process job: the CPU generator behind ``random_timesteps`` (and the
        data-loader shuffles) continues from ``seed + rank``.  Returns the rank used."""
        rank = self.global_rank
        torch.manual_seed(int(seed) + rank)
        return rank

    def reshape_timesteps(self, data: Tensor, t: Tensor) -> Tensor:
        return t.view((-1, *((1,) * (data.ndim - 1))))

    def get_schedule_parameters_at_time(self, data: Tensor, t: torch.LongTensor) -> dict:
        """alpha, beta, alpha-bar, sigma gathered at ``t`` and shaped for broadcasting (abstract_diffusion.py:194-220).
        Provided for API parity; the sampling kernel gathers from the device tables itself."""
        out = {}
        for key in ("alpha_t", "beta_t", "alpha_bar_t", "sigma_t"):
            out[key] = self.reshape_timesteps(data, self.schedule[key].to(data.device)[t])
        return out

    def forward_process(self, data: Tensor, t=None):
        raise NotImplementedError

    def reverse_process(self, *args, **kwargs):
        raise NotImplementedError

    @staticmethod
    def make_image_grid(batched_image, filename: str = None):
        """Writes a PNG grid when torchvision is importable (abstract_diffusion.py:240-276); visualisation only."""
        try:
            from torchvision.utils import make_grid, save_image
        except ImportError:  # pragma: no cover
            return None
        if isinstance(batched_image, list):
            batched_image = torch.stack(batched_image)
        batched_image = batched_image.detach().float().cpu()
        if batched_image.ndim != 4:
            return None   # volumes / profiles: nothing to plot (the reference asserts here)
        grid = make_grid(batched_image, normalize=True)
        if filename is not None:
            save_image(grid, fp=filename)
        return grid
