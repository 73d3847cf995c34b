"""Discrete parameter spaces for conditioning (reference ``rho_diffusion/data/parameter_space.py``)."""
from __future__ import annotations

import itertools
from collections import OrderedDict
from typing import Any

from rho_diffusion.utils import sample_from_discrete_parameter_space


class AbstractParameterSpace:
    """Dict-like container ``key -> list of admissible values`` (parameter_space.py:19-66)."""

    def __init__(self, param_dict=None, sampler=None):
        self.param_dict = param_dict if param_dict is not None else OrderedDict()
        self.sampler = sampler

    def set(self, param_dict) -> None:
        self.param_dict = param_dict

    @property
    def parameters(self):
        return self.param_dict.keys()

    def sample(self, num_samples, device=None):
        raise NotImplementedError("Method sample() is not implemented.")

    def size(self):
        raise NotImplementedError("Method size() is not implemented.")

    def push_parameter(self, key: str, value: Any) -> None:
        raise NotImplementedError("Method push_parameter() is not implemented.")

    def __repr__(self) -> str:
        return repr(self.param_dict)

    def __getitem__(self, key):
        return self.param_dict[key]

    def __setitem__(self, key, value):
        self.param_dict[key] = value

    def __len__(self) -> int:
        return len(self.param_dict)

    def items(self):
        return self.param_dict.items()

    def values(self):
        return self.param_dict.values()

    def keys(self):
        return self.param_dict.keys()


class DiscreteParameterSpace(AbstractParameterSpace):
    def __init__(self, param_dict=None, sampler=None):
        super().__init__(param_dict=param_dict, sampler=sampler or sample_from_discrete_parameter_space)

    def sample(self, num_samples, device=None):
        return self.sampler(self.param_dict, batch_size=num_samples, random=True, device=device)

    def size(self):
        return sum(1 for _ in itertools.product(*self.param_dict.values()))

    def push_parameter(self, key: str, value: Any) -> None:
        # the reference's version (parameter_space.py:87-95) cannot run (isinstance() with one argument); this is
        # the evident intent: extend the value list of `key` with unseen values
        current = self.param_dict.get(key)
        if current is None:
            current = self.param_dict[key] = []
        for v in (value if isinstance(value, (list, tuple)) else [value]):
            if v not in current:
                current.append(v)
