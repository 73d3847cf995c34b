"""DDPM pipeline (reference ``rho_diffusion/diffusion/ddpm.py``) on the B200 kernels.

Same constructor / method surface as the reference's ``DDPM``; what changes is how the work is issued:

* ``forward_process``  : one fused kernel (``rho_q_sample``) instead of a gather + ~6 elementwise launches;
* ``training_step``    : fused q_sample -> UNet -> fused MSE(+grad); no per-step host sync (the reference's NaN
  guard, ddpm.py:268-272, is opt-in via ``nan_guard=True``);
* ``reverse_process``  : the T-step ancestral loop is ONE captured CUDA graph (UNet eval + fused update with
  in-kernel Philox noise + timestep decrement) replayed T times; the step index lives in device memory.

Reference semantics that are kept on purpose: noise scale ``0.8*sqrt(beta_t)``, ``z = 0`` for
``t <= 1``, the UNet is still evaluated at ``t = 0`` and its output discarded, clamp to [-1, 1] after every step,
``len(alpha_bar_t)`` steps (T+1 for the cosine schedule).
"""
from __future__ import annotations

from collections.abc import Iterable, Mapping
from typing import Any, Union

import torch
from torch import Tensor, nn

from rho_diffusion.diffusion.abstract_diffusion import AbstractDiffusionPipeline
from rho_diffusion.registry import registry
from rho_diffusion.utils import sample_from_discrete_parameter_space, save_model_checkpoint

__all__ = ["DDPM"]


def _ops():
    from rho_diffusion.b200 import ops
    return ops


class _FusedMSE(nn.Module):
    """nn.MSELoss() semantics (mean over all elements) with the fused loss+gradient kernel."""

    def forward(self, pred, target):
        return _ops().MseLossFn.apply(pred, target)


class DDPM(AbstractDiffusionPipeline):
    def __init__(self, backbone, backbone_kwargs: dict, schedule, loss_func, timesteps: Union[int, Tensor] = 1000,
                 cond_fn: str = None, cond_fn_kwargs: dict = None, optimizer=None, opt_kwargs: Union[Mapping[str, Any], None] = {},
                 t_checkpoints=None, sampling_batch_size=10, sample_every_n_epochs=5, sample_parameter_space=None,
                 save_checkpoint_every_n_epochs=10, labels=None, nan_guard: bool = False, seed: int = 777):
        # `labels=` is what scripts/training_ddp.py:159 passes (the reference constructor rejects it); accepted, unused
        super().__init__(backbone=backbone, backbone_kwargs=backbone_kwargs, schedule=schedule, timesteps=timesteps,
                         cond_fn=cond_fn, cond_fn_kwargs=cond_fn_kwargs, optimizer=optimizer, opt_kwargs=opt_kwargs)
        if isinstance(loss_func, str):
            loss_func = registry.get("nn", loss_func)
        if isinstance(loss_func, type):
            loss_func = loss_func()
        if type(loss_func) is nn.MSELoss and loss_func.reduction == "mean":
            loss_func = _FusedMSE()
        elif not isinstance(loss_func, _FusedMSE):
            raise NotImplementedError(f"loss {type(loss_func).__name__}: only nn.MSELoss (the shipped configs' loss) has a B200 kernel")
        self.loss_func = loss_func
        self.t_checkpoints = t_checkpoints
        self.sampling_batch_size = sampling_batch_size
        self.sample_every_n_epochs = sample_every_n_epochs
        self.sample_parameter_space = sample_parameter_space
        self.save_weights_every_n_epochs = save_checkpoint_every_n_epochs
        self.nan_guard = nan_guard
        self.seed = int(seed)
        self._noise_calls = 0
        self._samplers = {}

    # -- noise ---------------------------------------------------------------------------------------------
    def noise(self, data: Tensor) -> Tensor:
        """Standard normal noise shaped like ``data`` (ddpm.py:101-102): counter-based Philox4x32-10 keyed by
        (seed + rank, call index), so runs are reproducible and graph-capturable."""
        self._noise_calls += 1
        return _ops().randn(data.shape, data.device, self.seed + 7919 * self.global_rank, self._noise_calls)

    def _tables(self, device):
        sch = self.schedule
        if sch.device != device:
            sch.to(device)
        with sch:  # float32 view of the tables
            return sch["alpha_t"], sch["beta_t"], sch["alpha_bar_t"]

    # -- forward (noising) process ----------------------------------------------------------------------------
    def forward_process(self, data: Tensor, t: Tensor | None = None, noise: Tensor | None = None) -> list[Tensor]:
        """x_t = sqrt(abar_t) x_0 + sqrt(1 - abar_t) eps; returns ``[x_t, eps]`` (ddpm.py:104-130)."""
        self.schedule.dtype = data.dtype
        if t is None:
            t = self.random_timesteps(data.size(0))
        t = t.reshape(-1).to(data.device)
        if noise is None:
            noise = self.noise(data)
        _, _, alpha_bar = self._tables(data.device)
        return [_ops().q_sample(data, noise, t, alpha_bar), noise]

    # -- training ------------------------------------------------------------------------------------------
    def training_step(self, batch: Iterable[Any], t: Tensor | None = None, noise: Tensor | None = None):
        """Noise-prediction loss for one batch (ddpm.py:231-288).  ``batch``: ``[data, labels]``, a dict with ``data`` /
        ``label`` keys, or a bare tensor.  ``t`` / ``noise`` may be supplied for shared-noise parity runs."""
        if isinstance(batch, (list, tuple)):
            data, labels = batch
        elif isinstance(batch, dict):
            data, labels = batch.get("data"), batch.get("label")
        else:
            data, labels = batch, None
        self.data_shape = data.shape
        self.data_dtype = data.dtype
        if t is None:
            t = self.random_timesteps(data.size(0))
        t = t.to(data.device)
        x_data, noise = self.forward_process(data, t, noise)
        if self.nan_guard and bool(x_data.isnan().any()):
            raise FloatingPointError("Noised data contains NaNs. Check your noise scheduler.")
        pred_noise = self.backbone(x_data, t, labels) if labels is not None else self.backbone(x_data, t)
        loss = self.loss_func(pred_noise, noise)
        self.log("train_loss", loss, prog_bar=True)
        return loss

    def forward(self, batch):
        return self.backbone(batch)

    # -- reverse (sampling) process ---------------------------------------------------------------------------
    @torch.inference_mode()
    def reverse_process(self, x_T: Tensor, conditions=None, t_checkpoints=None, *, x_init: Tensor | None = None,
                        noises: dict | None = None, use_cuda_graph: bool = True, num_steps: int | None = None) -> dict:
        """Ancestral sampling, Algorithm 2 of Ho et al. with the reference's constants (ddpm.py:132-229).

        ``x_T`` only provides shape / dtype / device (the reference draws ``randn_like(x_T)``).  ``x_init`` and
        ``noises`` (``{t: z_t}``) replace the internally generated noise for shared-noise parity runs.
        Returns ``{"buffer", "denoised"}`` like the reference.
        """
        from rho_diffusion.b200.sampler import AncestralSampler

        batch = x_T.size(0)
        device = x_T.device
        self.schedule.dtype = x_T.dtype
        if conditions is None:
            cc = None
        elif isinstance(conditions, int):
            cc = torch.full((batch,), conditions, device=device, dtype=torch.long)
        elif isinstance(conditions, str) and conditions == "auto":
            cc = torch.randint(0, 10, (batch,)).to(device)
        elif isinstance(conditions, torch.Tensor):
            cc = conditions.to(device)
        else:
            cc = torch.tensor(conditions).to(device)
        alpha, beta, alpha_bar = self._tables(device)
        key = (tuple(x_T.shape), None if cc is None else (tuple(cc.shape), cc.dtype), bool(use_cuda_graph), noises is not None)
        sampler = self._samplers.get(key)
        if sampler is None:
            sampler = AncestralSampler(self.backbone, tuple(x_T.shape), device, alpha, beta, alpha_bar, cond_example=cc,
                                       use_cuda_graph=use_cuda_graph, external_noise=noises is not None)
            self._samplers = {key: sampler}   # keep one (graphs pin their memory pool)
        if x_init is None:
            x_init = self.noise(x_T)
        self._noise_calls += 1
        return sampler.run(x_init, cc, seed=self.seed + 7919 * self.global_rank + 104729 * self._noise_calls, noises=noises,
                           t_checkpoints=t_checkpoints, num_steps=num_steps, progress=self.global_rank == 0)

    def on_train_epoch_end(self) -> None:
        if self.current_epoch > 0 and self.sample_every_n_epochs > 0 and self.current_epoch % self.sample_every_n_epochs == 0:
            self.eval()
            self.generate()
        if (self.current_epoch > 0 and self.save_weights_every_n_epochs > 0
                and self.current_epoch % self.save_weights_every_n_epochs == 0):
            self.eval()
            self.save_model_weights()

    def p_sample(self, parameter_space, random=False, shard: bool = True):
        """Builds the sample shape and the conditions and runs the reverse process (ddpm.py:319-355).  With a
        process group initialised and ``shard=True`` every rank samples its own contiguous slice of the batch
        (no communication); the returned dict holds the local slice."""
        import torch.distributed as dist

        if hasattr(self, "data_shape"):
            shape = [int(s) for s in self.data_shape]
            shape[0] = self.sampling_batch_size
        else:
            data_shape = self.backbone_kwargs["data_shape"]
            data_shape = [data_shape] if isinstance(data_shape, int) else list(data_shape)
            shape = [self.sampling_batch_size, self.backbone_kwargs["out_channels"]] + data_shape
            self.data_dtype = torch.float32
        device = self.device
        if isinstance(parameter_space, Tensor):
            # extension: explicit condition rows (e.g. same-shape label embeddings, one per sample) instead of a discrete space
            assert parameter_space.shape[0] >= shape[0], "generate(): fewer condition rows than sampling_batch_size"
            cond = parameter_space[: shape[0]].to(device)
        else:
            cond = sample_from_discrete_parameter_space(parameter_space, shape[0], random=random, device=device) \
                if parameter_space is not None else None
        if shard and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from rho_diffusion.b200.sampler import shard_range
            lo, hi = shard_range(shape[0], dist.get_world_size(), dist.get_rank())
            shape[0] = hi - lo
            cond = cond[lo:hi] if cond is not None else None
        sample = torch.zeros(shape, dtype=self.data_dtype, device=device)
        results = self.reverse_process(x_T=sample, conditions=cond, t_checkpoints=self.t_checkpoints)
        self.make_image_grid(results["denoised"], filename="output_%d.png" % self.current_epoch)
        return results

    def generate(self, parameter_space=None, random=False):
        if parameter_space is None:
            parameter_space = self.sample_parameter_space
        return self.p_sample(parameter_space=parameter_space, random=random)

    def save_model_weights(self):
        print("saving model checkpoints...")
        save_model_checkpoint(self.backbone, "model.pth")

    def validation_step(self, batch) -> float:
        return 0
