"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference (intel/rho-diffusion) on CPU.

Sources: ``/root/reference`` where it exists (the build container), else the installed copy under
``oracle/_ref`` that ``oracle/build_ref.py`` makes (git-ignored; it travels to the GPU box with the
working tree).  Used by ``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/``, by the live cross-check in ``tests/test_oracle_vs_reference.py`` and by the CPU arm of
``bench.py`` (``--impl reference`` / ``cpu_baseline``, always in its own process).  Nothing in the product
path imports this module.

The reference cannot be imported as a package (its ``__init__`` pulls lightning, diffusers,
matplotlib, h5py, geomloss -- all absent here), so the sub-modules on the hot path are imported
one by one from their real location through stub parent packages, with ~30 lines of shims for the
third-party bases they subclass.  No reference source is copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("RHO_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "rho_diffusion")):
    REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # pip-installed copy (oracle/build_ref.py)
REF_PKG = os.path.join(REF_ROOT, "rho_diffusion")

_SUBPACKAGES = ("", ".models", ".diffusion", ".data", ".metrics")


def available() -> bool:
    return os.path.isdir(REF_PKG)


def _stub(name, path=None, **attrs):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _LightningModuleShim(nn.Module):
    """Stand-in for lightning.pytorch.LightningModule (only what ddpm.py touches)."""

    def __init__(self):
        super().__init__()
        self.global_rank = 0
        self.current_epoch = 0

    def save_hyperparameters(self, d):
        self.hparams = types.SimpleNamespace(**d)

    def log(self, *a, **k):
        pass

    @property
    def device(self):
        return next(self.parameters()).device


class _PSNRShim(nn.Module):
    def forward(self, a, b):
        return torch.tensor(0.0)


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path symbols (imported from /root/reference)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if "rho_diffusion" in sys.modules and not getattr(sys.modules["rho_diffusion"], "__rho_ref_stub__", False):
        raise RuntimeError(
            "a different 'rho_diffusion' package is already imported in this process; "
            "the reference loader must run in its own process (see oracle/make_golden.py)"
        )
    for sub in _SUBPACKAGES:
        m = _stub("rho_diffusion" + sub, REF_PKG + sub.replace(".", "/"))
        m.__rho_ref_stub__ = True
    pl = _stub("lightning.pytorch", LightningModule=_LightningModuleShim, seed_everything=torch.manual_seed)
    _stub("lightning", pytorch=pl)
    _stub("torchmetrics")
    _stub("torchmetrics.image", PeakSignalNoiseRatio=_PSNRShim)
    if "matplotlib" not in sys.modules:
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
    _stub("diffusers")
    _stub("diffusers.schedulers")
    _stub("diffusers.schedulers.scheduling_utils", SchedulerMixin=object)

    ns = types.SimpleNamespace()
    ns.registry = importlib.import_module("rho_diffusion.registry").registry
    ns.layers = importlib.import_module("rho_diffusion.layers")
    ns.activations = importlib.import_module("rho_diffusion.activations")      # registers SymmetricLog
    ns.common = importlib.import_module("rho_diffusion.models.common")
    ns.unet_v2 = importlib.import_module("rho_diffusion.models.unet_v2")
    ns.conditioning = importlib.import_module("rho_diffusion.models.conditioning")
    ns.parameter_space = importlib.import_module("rho_diffusion.data.parameter_space")
    ns.schedule = importlib.import_module("rho_diffusion.diffusion.schedule")
    sys.modules["rho_diffusion.diffusion"].schedule = ns.schedule
    ns.utils = importlib.import_module("rho_diffusion.utils")
    ns.ddpm = importlib.import_module("rho_diffusion.diffusion.ddpm")
    ns.ema = importlib.import_module("rho_diffusion.ema")
    ns.gaussian = importlib.import_module("rho_diffusion.diffusion.gaussian_diffusion")
    ns.GaussianDiffusionPipeline = ns.gaussian.GaussianDiffusionPipeline
    ns.UNet = ns.unet_v2.UNet
    ns.DDPM = ns.ddpm.DDPM
    ns.MultiEmbeddings = ns.conditioning.MultiEmbeddings
    ns.ExponentialMovingAverage = ns.ema.ExponentialMovingAverage
    _loaded = ns
    return ns


def randomize_zero_modules(model: nn.Module, seed: int = 778, std: float = 0.02) -> None:
    """The reference zero-initialises the ResBlock out conv, attention proj_out and the final conv
    (layers.py:113-119), so a fresh model outputs exactly 0.  Goldens are captured after every
    all-zero parameter tensor is re-drawn from N(0, std^2)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * std)
