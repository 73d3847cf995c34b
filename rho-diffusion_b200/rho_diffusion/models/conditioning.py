"""Parameter-space conditioning (reference ``rho_diffusion/models/conditioning.py:31-139``)."""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Union

import torch
from torch import nn

from rho_diffusion.registry import registry


@registry.register_layer("MultiEmbeddings")
class MultiEmbeddings(nn.Module):
    """Sum over parameter-space keys of ``nn.Embedding[index of y[:, i] within space[key]]``.

    Same constructor and state-dict keys (``embedding_layers.<key>.weight``) as the reference.  The value ->
    index match and the gather/sum run in one kernel per key (``rho_embedding_match``) instead of building a
    device tensor from a Python list on every call (conditioning.py:132).
    """

    def __init__(self, parameter_space=None, embedding_dim: int = 512, parameter_space_dim: int = 3,
                 embedding_size: Union[int, list, dict, OrderedDict] = None) -> None:
        super().__init__()
        self.embedding_layers = nn.ModuleDict()
        self.label_encoders = OrderedDict()
        self.parameter_space = parameter_space
        self.embedding_dim = embedding_dim
        if parameter_space is not None and len(parameter_space) > 0:
            for key, value in parameter_space.items():
                self.embedding_layers[key] = nn.Embedding(num_embeddings=len(value), embedding_dim=embedding_dim)
        elif embedding_size is not None:
            if isinstance(embedding_size, int):
                assert parameter_space_dim is not None, "parameter_space_dim should be an int input when embedding_size is an int"
                sizes = {str(i): embedding_size for i in range(parameter_space_dim)}
            elif isinstance(embedding_size, list):
                sizes = {str(i): n for i, n in enumerate(embedding_size)}
            else:
                sizes = dict(embedding_size)
            for key, n in sizes.items():
                self.embedding_layers[key] = nn.Embedding(num_embeddings=n, embedding_dim=embedding_dim)
        self._value_tables = {}
        self._labels_checked = False

    def _values(self, device):
        tabs = self._value_tables.get(device)
        if tabs is None:
            if self.parameter_space is None:
                raise RuntimeError("MultiEmbeddings.forward needs a parameter_space to map values to indices")
            with torch.inference_mode(False):
                tabs = [torch.tensor([float(v) for v in self.parameter_space[key]], dtype=torch.float32, device=device)
                        for key in self.embedding_layers.keys()]
            self._value_tables[device] = tabs
        return tabs

    def forward(self, y: torch.Tensor) -> torch.Tensor:
        from rho_diffusion.b200 import ops

        tables = [layer.weight for layer in self.embedding_layers.values()]
        out = ops.EmbeddingMatchFn.apply(y, self._values(y.device), *tables)
        # A label outside the parameter space makes the reference fail (conditioning.py:132-137).  The kernel poisons such rows
        # with NaN; the first eager call (and every call under RHO_STRICT_LABELS=1) also checks the matched indices on the host
        # and raises -- one device sync, skipped while a CUDA graph is being captured.
        strict = os.environ.get("RHO_STRICT_LABELS", "") == "1"
        if (strict or not self._labels_checked) and not torch.cuda.is_current_stream_capturing():
            self._labels_checked = True
            idxs = ops.EmbeddingMatchFn.last_indices
            if idxs is not None and bool((idxs < 0).any()):
                bad = (idxs < 0).nonzero()[0].tolist()
                key = list(self.embedding_layers.keys())[bad[0]]
                raise ValueError(f"MultiEmbeddings: label value of sample {bad[1]} for key {key!r} is not a point of the parameter space "
                                 f"{list(self.parameter_space[key])}")
        return out
