"""Noise schedules (reference ``rho_diffusion/diffusion/schedule.py``).

The beta / alpha / alpha-bar / sigma tables are built once on the host in float64 and stored as float32 -- the
same arithmetic, in the same order, as schedule.py:141-214, so the tables are bit-identical to the reference's
(checked against golden vectors in tests/).  They live on the requested device (``device=`` is honoured and
``.to(device)`` really moves them; the reference's ``to`` is a no-op, schedule.py:133-138).
"""
from __future__ import annotations

import math
from copy import deepcopy

import torch
from torch import Tensor

from rho_diffusion.registry import registry

__all__ = ["LinearSchedule", "CosineBetaSchedule", "SigmoidSchedule"]

_KEYS = ("alpha_t", "beta_t", "alpha_bar_t", "sigma_t")


def _shift_right(ab32: Tensor) -> Tensor:
    """alpha-bar of the previous step, 1 at t = 0 (``offset_alpha_bar_t``)."""
    return torch.cat([torch.ones(1, dtype=ab32.dtype, device=ab32.device), ab32[:-1]])


class AbstractSchedule:
    """Table container with the reference's accessors: ``schedule["alpha_bar_t"]``, ``.dtype`` cast on read,
    ``offset_alpha_bar_t``, ``state`` / ``step`` / ``reset`` bookkeeping and the float32 context manager."""

    def __init__(self) -> None:
        self._dtype = torch.float32

    # -- dtype handling (values are cast on access, schedule.py:35-46, 54-92)
    @property
    def dtype(self) -> torch.dtype:
        return getattr(self, "_dtype", None) or torch.float32

    @dtype.setter
    def dtype(self, value) -> None:
        self._dtype = value or torch.float32

    def _table(self, key: str) -> Tensor:
        return getattr(self, "_" + key).type(self.dtype)

    def __getitem__(self, key: str) -> Tensor:
        return getattr(self, key)

    def __enter__(self):
        self._saved_dtype = self.dtype
        self.dtype = torch.float32

    def __exit__(self, *exc):
        self.dtype = self._saved_dtype

    @property
    def offset_alpha_bar_t(self) -> Tensor:
        return _shift_right(self.alpha_bar_t)

    # -- step bookkeeping (schedule.py:48-52, 94-124)
    @property
    def index(self) -> int:
        return self._index

    @index.setter
    def index(self, value: int) -> None:
        self._index = value

    def state(self, index=None) -> dict:
        index = index or self.index
        return {k: getattr(self, k)[index] for k in _KEYS}

    @property
    def last_state(self) -> dict:
        return self._last_state

    @last_state.setter
    def last_state(self, state) -> None:
        self._last_state = deepcopy(state)

    def reset(self) -> None:
        self.index = 0
        self.last_state = {}

    def step(self) -> None:
        if not hasattr(self, "_index"):
            self.reset()
        else:
            self.last_state = self.state()
            self.index += 1

    def to(self, device) -> "AbstractSchedule":
        for k in _KEYS:
            setattr(self, "_" + k, getattr(self, "_" + k).to(device))
        return self

    @property
    def device(self):
        return self._beta_t.device

    def __len__(self) -> int:
        return len(self._alpha_bar_t)

    def _store(self, device, **tables) -> None:
        for k, v in tables.items():
            setattr(self, "_" + k, v.to(torch.float32).to(device))


def _make_prop(key):
    def getter(self):
        return self._table(key)

    def setter(self, value):
        setattr(self, "_" + key, value)

    return property(getter, setter)


for _k in _KEYS:
    setattr(AbstractSchedule, _k, _make_prop(_k))


@registry.register_schedule("LinearSchedule")
class LinearSchedule(AbstractSchedule):
    """beta linear in [scale*beta_1, scale*beta_T], scale = 1000/T (schedule.py:141-168; note beta_1 defaults to 1e-3)."""

    def __init__(self, num_steps: int, beta_1: float = 1.0e-3, beta_T: float = 0.02, device="cpu") -> None:
        super().__init__()
        scale = 1000 / num_steps
        beta = torch.linspace(scale * beta_1, scale * beta_T, num_steps, dtype=torch.float64)
        alpha = 1.0 - beta
        alpha_bar = alpha.cumprod(0)
        prev32 = _shift_right(alpha_bar.to(torch.float32))      # the reference shifts the float32 table
        sigma = torch.sqrt((1 - prev32) / (1 - alpha_bar) * beta)
        self._store(device, beta_t=beta, alpha_t=alpha, alpha_bar_t=alpha_bar, sigma_t=sigma)


@registry.register_schedule("CosineBetaSchedule")
class CosineBetaSchedule(AbstractSchedule):
    """Nichol & Dhariwal cosine schedule on T+1 points (schedule.py:171-214): ``len(schedule) == num_steps + 1`` and
    ``sigma_t[0]`` is NaN (0/0), exactly as in the reference."""

    def __init__(self, num_steps: int, offset: float = 0.008, device="cpu") -> None:
        super().__init__()
        t = torch.linspace(0.0, num_steps, num_steps + 1, dtype=torch.float64) / num_steps
        alpha_bar = torch.cos((t + offset) / (1 + offset) * math.pi * 0.5).pow(2.0)
        alpha_bar = alpha_bar / alpha_bar[0]
        ab32 = alpha_bar.to(torch.float32).clamp_(0.0, 1.0)
        prev32 = _shift_right(ab32)
        beta = (1 - alpha_bar / prev32).clip_(0.0001, 0.9999)
        sigma = torch.sqrt((1 - prev32) / (1 - alpha_bar) * beta)
        self._store(device, beta_t=beta, alpha_t=1 - beta, alpha_bar_t=ab32, sigma_t=sigma)


@registry.register_schedule("SigmoidSchedule")
class SigmoidSchedule(AbstractSchedule):
    def __init__(self, num_steps: int, offset: float = 0.008) -> None:
        super().__init__()
        raise NotImplementedError("SigmoidSchedule is not yet implemented.")
