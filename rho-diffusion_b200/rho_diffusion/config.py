"""JSON experiment configs with the reference's schema (``rho_diffusion/config.py:36-110``).

The reference validates with pydantic **1.10.17** (pyproject.toml:44).  That dependency is not required here;
the schema, the ``number_cast`` validator and -- deliberately -- pydantic-v1's coercion rules for
``kwargs: dict[str, Union[str, float, list]]`` are restated directly:

* v1 tries ``str`` first and accepts bool / int / float for it, so ``"use_new_attention_order": false``
  becomes the **string** ``'False'`` (truthy!) and numbers become numeric strings, which ``number_cast``
  then turns back into int / float.  The shipped example configs therefore select ``QKVAttention``
  (the "new" order) in the reference; a checkpoint trained there only reproduces with the same choice.
  ``RHO_CONFIG_BOOL_MODE=native`` (or ``ExperimentConfig.from_json(path, bool_mode="native")``) keeps JSON
  booleans as Python bools instead.
* unknown keys (``training.np``, ``training.benchmark_mode``) are ignored; missing required fields and
  wrongly typed sections raise ``ConfigValidationError`` (a ``ValueError``).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Union

from rho_diffusion.utils import number_cast_dict


class ConfigValidationError(ValueError):
    pass


_MISSING = object()


def _coerce_v1_scalar(v, bool_mode):
    """pydantic-v1 ``Union[str, float]``: str first (bools/numbers are stringified), then float."""
    if isinstance(v, bool):
        return str(v) if bool_mode == "pydantic_v1" else v
    if isinstance(v, (int, float)):
        return str(v)
    if isinstance(v, str):
        return v
    raise ConfigValidationError(f"value {v!r} is not a str, number or list")


class _Model:
    """Minimal typed record: ``_fields = {name: (caster, default)}``; attribute and ``dict()`` access."""
    _fields: dict = {}

    def __init__(self, **data):
        if not isinstance(data, dict):
            raise ConfigValidationError(f"{type(self).__name__}: expected an object")
        for name, (caster, default) in self._fields.items():
            if name in data:
                try:
                    value = caster(data[name])
                except ConfigValidationError:
                    raise
                except Exception as exc:  # noqa: BLE001
                    raise ConfigValidationError(f"{type(self).__name__}.{name}: {exc}") from exc
            elif default is _MISSING:
                raise ConfigValidationError(f"{type(self).__name__}.{name}: field required")
            else:
                value = default
            setattr(self, name, value)

    def dict(self) -> dict:
        return {k: (v.dict() if isinstance(v, _Model) else v) for k, v in ((n, getattr(self, n)) for n in self._fields)}

    def __repr__(self) -> str:
        return " ".join(f"{n}={getattr(self, n)!r}" for n in self._fields)

    def __eq__(self, other):
        return type(self) is type(other) and self.dict() == other.dict()


def _strict_str(v):
    if isinstance(v, (str, int, float)) and not isinstance(v, bool):
        return str(v)
    raise ConfigValidationError(f"str type expected, got {type(v).__name__}")


def _int(v):
    if isinstance(v, bool) or (isinstance(v, float) and not v.is_integer()):
        raise ConfigValidationError(f"value is not a valid integer: {v!r}")
    return int(v)


def _path_or_none(v):
    return None if v is None else (v if isinstance(v, Path) else str(v))


_bool_mode_default = lambda: os.environ.get("RHO_CONFIG_BOOL_MODE", "pydantic_v1")  # noqa: E731


class ComponentConfig(_Model):
    """``{name, kwargs}`` (config.py:36-48)."""

    def __init__(self, bool_mode: str | None = None, **data):
        mode = bool_mode or _bool_mode_default()

        def kwargs_caster(kw):
            if not isinstance(kw, dict):
                raise ConfigValidationError("ComponentConfig.kwargs: value is not a valid dict")
            out = {}
            for k, v in kw.items():
                out[str(k)] = [_coerce_v1_scalar(x, mode) for x in v] if isinstance(v, list) else _coerce_v1_scalar(v, mode)
            return number_cast_dict(out)

        self._fields = {"name": (_strict_str, _MISSING), "kwargs": (kwargs_caster, _MISSING)}
        super().__init__(**data)


class TrainingConfig(_Model):
    """config.py:51-64."""
    _fields = {
        "device": (_strict_str, _MISSING), "loss_fn": (_strict_str, "MSELoss"), "ema_decay": (float, 0.0),
        "batch_size": (_int, 32), "seed": (_int, 777), "min_epochs": (_int, 1), "max_epochs": (_int, 999),
        "save_checkpoint_every_n_epochs": (_int, 10), "sample_every_n_epochs": (_int, 5),
    }


def _any_dict(v):
    if not isinstance(v, dict):
        raise ConfigValidationError("value is not a valid dict")
    return dict(v)


class InferenceConfig(_Model):
    """config.py:67-77."""
    _fields = {
        "device": (_strict_str, _MISSING), "checkpoint": (_path_or_none, _MISSING), "parameter_space": (_any_dict, _MISSING),
        "cache_file": (_path_or_none, None), "plot_output_file": (_path_or_none, None), "seed": (_int, 777),
    }


class ExperimentConfig(_Model):
    """config.py:80-110."""

    def __init__(self, bool_mode: str | None = None, **data):
        def comp(v):
            if not isinstance(v, dict):
                raise ConfigValidationError("value is not a valid dict")
            return ComponentConfig(bool_mode=bool_mode, **v)

        def sub(cls):
            def cast(v):
                if not isinstance(v, dict):
                    raise ConfigValidationError("value is not a valid dict")
                return cls(**v)
            return cast

        self._fields = {
            "experiment": (_strict_str, _MISSING), "model": (comp, _MISSING), "dataset": (comp, _MISSING),
            "optimizer": (comp, _MISSING), "lr_scheduler": (comp, _MISSING), "noise_schedule": (comp, _MISSING),
            "training": (sub(TrainingConfig), _MISSING), "inference": (sub(InferenceConfig), _MISSING),
        }
        super().__init__(**data)

    @classmethod
    def from_json(cls, json_path: Union[str, Path], bool_mode: str | None = None) -> "ExperimentConfig":
        json_path = Path(json_path)
        if not json_path.exists():
            raise FileNotFoundError(f"Specified config file not found: {json_path}")
        with open(json_path) as fh:
            return cls(bool_mode=bool_mode, **json.load(fh))
