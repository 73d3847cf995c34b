"""Synthetic density fields (reference ``rho_diffusion/data/synthetic.py``): the spherical-harmonic volumes that
``examples/config_spherical_harmonics.json`` trains on.  Host-side data generation (numpy + scipy), not on the GPU hot
path: the trainer only sees the resulting ``(density, sha512 label embedding)`` pairs.

Semantics kept from the reference (synthetic.py:45-124, 252-300):
* grid ``linspace(-2, 2, grid_el)`` per axis, ``meshgrid(indexing="xy")``;
* ``theta = arctan(sqrt(x^2 + y^2) / z)``, ``phi = arctan(y / x)`` and the call ``sph_harm(|m|, l, theta, phi)`` -- i.e.
  scipy's (order, degree, AZIMUTH, POLAR) signature receives ``theta`` as the azimuth and ``phi`` as the polar angle
  (reference quirk, kept; ``scipy.special.sph_harm`` was removed in SciPy 1.17, ``sph_harm_y(l, m, polar, azimuth)`` is
  the same function with the angle arguments swapped);
* ``solution = Y * r`` min-max normalised on the COMPLEX values (numpy complex ordering is lexicographic), density =
  ``|solution|``; labels ``(l, m)`` with ``l = randint(0, max_l)``, ``m = randint(-l, l)`` from Python's ``random`` seeded
  with ``random_seed`` (default ``PL_GLOBAL_SEED`` or 1616); label embedding ``calculate_sha512_embedding({"l", "m"}, 256)``.
"""
from __future__ import annotations

from os import getenv
from random import Random

import numpy as np
import torch

from rho_diffusion.data.parameter_space import DiscreteParameterSpace
from rho_diffusion.registry import registry
from rho_diffusion.utils import calculate_sha512_embedding

__all__ = ["make_spherical_grid", "compute_spherical_harmonic", "SphericalHarmonicDataset"]


def _sph_harm(m, l, azimuth, polar):
    """scipy.special.sph_harm(m, n, theta=azimuth, phi=polar) on any SciPy version."""
    from scipy import special
    if hasattr(special, "sph_harm_y"):
        return special.sph_harm_y(l, m, polar, azimuth)
    return special.sph_harm(m, l, azimuth, polar)     # pragma: no cover (old SciPy)


def make_spherical_grid(x: np.ndarray, y: np.ndarray, z: np.ndarray):
    xg, yg, zg = np.meshgrid(x, y, z, indexing="xy")
    theta = np.arctan(np.sqrt(xg ** 2 + yg ** 2) / zg)
    phi = np.arctan(yg / xg)
    return np.array([xg, yg, zg]), theta, phi


def compute_spherical_harmonic(x, y, z, m: int, l: int, normalize: bool = True):
    xyz, theta, phi = make_spherical_grid(x, y, z)
    radial = np.sqrt((xyz ** 2).sum(axis=0))
    solution = _sph_harm(np.abs(m), l, theta, phi) * radial
    if normalize:
        lo, hi = solution.min(), solution.max()
        solution = (solution - lo) / (hi - lo)
    return xyz, np.abs(solution), np.real(solution)


@registry.register_dataset("SphericalHarmonicDataset")
class SphericalHarmonicDataset(torch.utils.data.Dataset):
    parameter_space = None          # class-level default of the reference's MultiVariateDataset (data/base.py:27)

    def __init__(self, max_l: int | None, h5_path=None, length: int = 1000, random_seed: int | None = None,
                 use_emb_as_labels: bool = True, **grid_kwargs):
        if h5_path is not None:
            raise NotImplementedError("the HDF5 cache of the reference needs h5py, which is not a dependency of the B200 package")
        if isinstance(max_l, int):
            assert max_l > 0, f"Invalid maximum value of l > 0: {max_l}"
        self.parameter_space = DiscreteParameterSpace(param_dict={"l": list(range(0, max_l)), "m": list(range(-max_l, max_l))})
        self.loaded_parameter_space = DiscreteParameterSpace()
        self.max_l = max_l
        self.random_seed = int(random_seed) if random_seed else int(getenv("PL_GLOBAL_SEED", 1616))
        self._rng = Random(self.random_seed)
        grid_kwargs.setdefault("grid_el", 32)
        for key in ("x", "y", "z"):
            grid_kwargs.setdefault(key, np.linspace(-2.0, 2.0, grid_kwargs["grid_el"]))
        self.grid_kwargs = grid_kwargs
        self.length = length
        self.labels_emb_map = {}

    @property
    def random_set(self) -> tuple[int, int]:
        l = self._rng.randint(0, self.max_l)  # noqa: E741
        return l, self._rng.randint(-l, l)

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, index: int):
        l, m = self.random_set  # noqa: E741
        _, density, _ = compute_spherical_harmonic(self.grid_kwargs["x"], self.grid_kwargs["y"], self.grid_kwargs["z"], m=m, l=l)
        c = {"l": l, "m": m}
        emb = calculate_sha512_embedding(c, l=256)
        self.labels_emb_map[emb] = c
        return torch.from_numpy(density.astype(np.float32)).unsqueeze(0), emb
