"""rho_diffusion -- B200-native implementation of the reference's DDPM hot path (drop-in package name).

Importing the package registers the models / layers / schedules under the reference's registry names.  The B200
backend (``rho_diffusion.b200``) is imported lazily; like the reference's XPU plug-in (``__init__.py:27-35``) a
missing backend is reported and the import continues (configs, registry and schedules work without it), but any
attempt to *run* a layer without the CUDA extension raises.
"""
from __future__ import annotations

__version__ = "0.1.2+b200.r1"
__all__ = []

from rho_diffusion import registry as _registry  # noqa: F401,E402
from rho_diffusion import activations, layers, models, data, diffusion, ema, config, utils  # noqa: F401,E402

try:
    from rho_diffusion import b200  # noqa: F401
except (ModuleNotFoundError, ImportError, OSError) as error:  # pragma: no cover
    print(error)
