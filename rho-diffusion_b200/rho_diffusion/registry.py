"""Name -> class registry with the reference's categories and decorator names.

API parity with the reference's ``rho_diffusion/registry.py:28-129`` (categories, ``register_*``
decorators, ``get`` semantics incl. the KeyError / AssertionError behaviour, pre-population with
torch activations, optimizers and every ``nn.Module``, ``:160-203``), implemented as a table-driven
class instead of seven hand-written decorators.
"""
from __future__ import annotations

from typing import Any, Callable

from torch import nn, optim

_CATEGORIES = ("models", "activations", "layers", "datasets", "nn", "schedules", "optimizers")
_SINGULAR = {"models": "model", "activations": "activation", "layers": "layer", "datasets": "dataset", "nn": "nn",
             "schedules": "schedule", "optimizers": "optimizer"}


def _make_register(category: str) -> Callable[[type, str], Callable]:
    def register(cls, name: str) -> Callable:
        def wrapper(target):
            cls.mapping[category][name] = target
            return target

        return wrapper

    register.__name__ = f"register_{_SINGULAR[category]}"
    register.__doc__ = f"Decorator: register ``target`` under ``{category}[name]``."
    return classmethod(register)


class Registry:
    mapping = {c: {} for c in _CATEGORIES}

    def get(self, category: str, name: str) -> Any:
        """Look up ``name`` in ``category``; AssertionError for an unknown category, KeyError for an
        unknown (or falsy) entry -- as in the reference."""
        assert category in self.mapping, (
            f"{category} is not a category within Registry - valid entries: {self.mapping.keys()}.")
        found = self.mapping[category].get(name, None)
        if not found:
            raise KeyError(f"{name} is not a member of {category} category in Registry.")
        return found

    @property
    def categories(self) -> list[str]:
        return list(self.mapping.keys())

    def __repr__(self) -> str:
        lines = ["Registry", "========"]
        lines += [f"{k}: {v}" for k, v in self.mapping.items()]
        return "\n".join(lines) + "\n"


for _c in _CATEGORIES:
    setattr(Registry, f"register_{_SINGULAR[_c]}", _make_register(_c))

registry = Registry()

# torch building blocks the reference pre-registers (registry.py:160-203)
for _n in ("ReLU", "SiLU", "Tanh", "Sigmoid", "ELU", "GELU", "PReLU", "Softmax", "LogSoftmax"):
    Registry.register_activation(_n)(getattr(nn, _n))
for _n in ("ASGD", "Adadelta", "Adagrad", "Adam", "AdamW", "Adamax", "LBFGS", "NAdam", "RAdam", "RMSprop", "Rprop", "SGD",
           "SparseAdam"):
    Registry.register_optimizer(_n)(getattr(optim, _n))
for _n in dir(nn):
    _obj = getattr(nn, _n, None)
    if _n != "Module" and isinstance(_obj, type) and issubclass(_obj, nn.Module):
        Registry.register_nn(_n)(_obj)
