"""Synthetic density fields (reference ``rho_diffusion/data/synthetic.py``): the spherical-harmonic volumes that
``examples/config_spherical_harmonics.json`` trains on, plus n-D Gaussian-blob fields.

Two ways to get the data, with identical samples: ``dataset[i]`` (host: numpy + scipy, what a ``DataLoader`` calls, as in the
reference) and ``dataset.device_loader(batch_size, device)`` -- batches generated ON THE DEVICE by one kernel launch each
(``rho_sph_harm_density`` / ``rho_gaussian_blobs`` of librho_b200.so), which keeps the input pipeline off the host at 8 GPUs;
``scripts/training_ddp.py`` uses it whenever the dataset offers it.

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

__all__ = ["make_spherical_grid", "compute_spherical_harmonic", "SphericalHarmonicDataset", "GaussianBlobDataset", "DeviceLoader"]


class DeviceLoader:
    """Iterates ``len(dataset) // (batch_size * world)`` batches per epoch, each produced by ``dataset.device_batch`` on ``device``
    (no host tensors, no worker processes).  ``set_epoch`` is accepted for symmetry with ``DistributedSampler``."""

    def __init__(self, dataset, batch_size: int, device, world: int = 1):
        self.dataset, self.batch_size, self.device, self.world = dataset, int(batch_size), torch.device(device), max(int(world), 1)

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __len__(self) -> int:
        return max(len(self.dataset) // (self.batch_size * self.world), 1)

    def __iter__(self):
        for _ in range(len(self)):
            yield self.dataset.device_batch(self.batch_size, self.device)


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

    # -- device path ---------------------------------------------------------------------------------------------
    def _linspace_grid(self):
        """(lo, hi, G) when the three axes are the same ``linspace`` (the only grids the device kernel generates), else None."""
        g = self.grid_kwargs
        G = int(g["grid_el"])
        for key in ("x", "y", "z"):
            a = np.asarray(g[key], dtype=np.float64)
            if a.shape != (G,) or not np.allclose(a, np.linspace(a[0], a[-1], G), rtol=0, atol=1e-12):
                return None
        lo, hi = float(np.asarray(g["x"])[0]), float(np.asarray(g["x"])[-1])
        if any(float(np.asarray(g[k])[0]) != lo or float(np.asarray(g[k])[-1]) != hi for k in ("y", "z")):
            return None
        return lo, hi, G

    def seed_for_rank(self, rank: int) -> None:
        """Distinct (l, m) streams per data-parallel rank (the reference's ranks all draw the same stream: __getitem__ ignores
        the sampler's index)."""
        self._rng = Random(self.random_seed + 7919 * int(rank))

    def device_batch(self, batch_size: int, device):
        """The next ``batch_size`` samples -- the same (l, m) stream ``__getitem__`` consumes -- as ``(density (B, 1, G, G, G),
        label embeddings (B, 256))`` on ``device``; one kernel pair for the whole batch."""
        from rho_diffusion.b200 import ops
        grid = self._linspace_grid()
        if grid is None:
            raise NotImplementedError("device generation needs the three axes to be one common linspace (the default grid)")
        lo, hi, G = grid
        lm = [self.random_set for _ in range(batch_size)]
        embs = []
        for l, m in lm:  # noqa: E741
            key = (l, m)
            emb = self._emb_cache.get(key) if hasattr(self, "_emb_cache") else None
            if emb is None:
                self.__dict__.setdefault("_emb_cache", {})
                emb = calculate_sha512_embedding({"l": l, "m": m}, l=256)
                self._emb_cache[key] = emb
                self.labels_emb_map[emb] = {"l": l, "m": m}
            embs.append(emb)
        lt = torch.tensor([v[0] for v in lm], dtype=torch.int32).to(device, non_blocking=True)
        mt = torch.tensor([v[1] for v in lm], dtype=torch.int32).to(device, non_blocking=True)
        mm = torch.tensor([self._extremes(l, m) for l, m in lm], dtype=torch.float64).to(device, non_blocking=True)  # noqa: E741
        return ops.sph_harm_density(lt, mt, G, lo, hi, minmax=mm), torch.stack(embs).to(device, non_blocking=True)

    def _extremes(self, l: int, m: int):  # noqa: E741
        """(re, im) of the complex min and max the reference normalises with, from the host restatement (cached per (l, |m|)):
        numpy orders complex numbers lexicographically and symmetric grid points tie in the real part up to rounding noise of the
        special-function library, so these two numbers cannot be reproduced by an independent implementation -- everything else
        of the field is generated on the device."""
        cache = self.__dict__.setdefault("_mm_cache", {})
        key = (l, abs(m))
        if key not in cache:
            g = self.grid_kwargs
            xyz, theta, phi = make_spherical_grid(g["x"], g["y"], g["z"])
            sol = _sph_harm(abs(m), l, theta, phi) * np.sqrt((xyz ** 2).sum(axis=0))
            mn, mx = sol.min(), sol.max()
            cache[key] = (float(np.real(mn)), float(np.imag(mn)), float(np.real(mx)), float(np.imag(mx)))
        return cache[key]

    def device_loader(self, batch_size: int, device, rank: int = 0, world: int = 1) -> DeviceLoader:
        if world > 1:
            self.seed_for_rank(rank)
        return DeviceLoader(self, batch_size, device, world)


def _philox4x32_10(ctr_lo: int, ctr_hi: int, key: int):
    """Host restatement of the generator in csrc/datagen.cu (Salmon et al. 2011): four 32-bit words."""
    M = 0xFFFFFFFF
    c = [ctr_lo & M, (ctr_lo >> 32) & M, ctr_hi & M, (ctr_hi >> 32) & M]
    k0, k1 = key & M, (key >> 32) & M
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M, p1 & M, ((p0 >> 32) ^ c[3] ^ k1) & M, p0 & M]
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c


@registry.register_dataset("GaussianBlobDataset")
class GaussianBlobDataset(torch.utils.data.Dataset):
    """n-D (n = 1, 2, 3) Gaussian-blob density fields in [-1, 1] with a discrete parameter space (the synthetic
    stand-in for density maps such as DeepGalaxy's when the real HDF5 is absent):

        rho(x) = sum_{k < K} exp(-|x - c_k|^2 / (2 w^2)),   field = 2 min(rho, 1) - 1,   x at the cell centres of [0, 1]^n

    Labels are the float rows ``(K, w)`` -- the form ``MultiEmbeddings`` conditions on (``cond_fn="MultiEmbeddings"`` over the
    class-level ``parameter_space``, like ``DeepGalaxyDataset``).  Sample ``i`` is a pure function of ``(random_seed, i)``: the
    centres come from Philox4x32-10 (key = seed, counter = (i, k)), so the host path (``dataset[i]``) and the device path
    (``device_batch``: one ``rho_gaussian_blobs`` launch per batch) produce the same field; the first two moments of a field
    are smooth functions of ``(K, w)``, which is what the density-moments acceptance check of tools/moments_acceptance.py uses."""
    parameter_space = DiscreteParameterSpace(param_dict={"n": [1, 2, 3, 4], "w": [0.05, 0.1, 0.2]})

    def __init__(self, data_shape, length: int = 1000, random_seed: int | None = None, n_values=None, w_values=None):
        self.data_shape = [int(s) for s in ([data_shape] if isinstance(data_shape, int) else data_shape)]
        assert 1 <= len(self.data_shape) <= 3, "GaussianBlobDataset generates 1-D, 2-D or 3-D fields"
        self.length = int(length)
        self.random_seed = int(random_seed) if random_seed else int(getenv("PL_GLOBAL_SEED", 1616))
        space = type(self).parameter_space
        self.n_values = [int(v) for v in (n_values if n_values is not None else space["n"])]
        self.w_values = [float(v) for v in (w_values if w_values is not None else space["w"])]
        assert max(self.n_values) <= 8, "at most 8 blobs per field"
        self.loaded_parameter_space = DiscreteParameterSpace(param_dict={"n": self.n_values, "w": self.w_values})
        self._cursor = 0           # device path: index of the next sample
        self._rank, self._world = 0, 1

    def __len__(self) -> int:
        return self.length

    def params_of(self, index: int):
        """(K, w) of sample ``index``: words 2 and 3 of Philox(counter = (index, 255))."""
        r = _philox4x32_10(index, 255, self.random_seed)
        return self.n_values[r[2] % len(self.n_values)], self.w_values[r[3] % len(self.w_values)]

    def centres_of(self, index: int, K: int) -> np.ndarray:
        out = np.empty((K, 3), dtype=np.float32)
        for k in range(K):
            r = _philox4x32_10(index, k, self.random_seed)
            for a in range(3):
                out[k, a] = np.float32(0.2) + np.float32(0.6) * (np.float32(r[a] >> 8) + np.float32(0.5)) * np.float32(5.9604644775390625e-08)
        return out

    def __getitem__(self, index: int):
        K, w = self.params_of(index)
        c = self.centres_of(index, K)
        axes = [(np.arange(s, dtype=np.float32) + np.float32(0.5)) / np.float32(s) for s in self.data_shape]
        grids = np.meshgrid(*axes, indexing="ij")
        rho = np.zeros(self.data_shape, dtype=np.float32)
        nd = len(self.data_shape)
        for k in range(K):
            d2 = np.zeros(self.data_shape, dtype=np.float32)
            for a in range(nd):                       # coordinate 0 of a centre belongs to the LAST axis
                d2 += (grids[nd - 1 - a] - c[k, a]) ** 2
            rho += np.exp(-d2 * np.float32(0.5 / (w * w)))
        field = 2.0 * np.minimum(rho, 1.0) - 1.0
        return torch.from_numpy(field.astype(np.float32)).unsqueeze(0), torch.tensor([float(K), float(w)], dtype=torch.float32)

    def device_batch(self, batch_size: int, device):
        """Samples ``[first, first + batch_size)`` with ``first = (batches drawn so far * world + rank) * batch_size``: disjoint
        index ranges per rank, an endless stream (a field is a pure function of its index; ``length`` only sets the epoch size)."""
        from rho_diffusion.b200 import ops
        first = (self._cursor * self._world + self._rank) * batch_size
        self._cursor += 1
        params = [self.params_of(first + i) for i in range(batch_size)]
        n = torch.tensor([p[0] for p in params], dtype=torch.int32).to(device, non_blocking=True)
        w = torch.tensor([p[1] for p in params], dtype=torch.float32).to(device, non_blocking=True)
        labels = torch.tensor([[float(p[0]), float(p[1])] for p in params], dtype=torch.float32).to(device, non_blocking=True)
        return ops.gaussian_blobs(n, w, self.data_shape, self.random_seed, first), labels

    def device_loader(self, batch_size: int, device, rank: int = 0, world: int = 1) -> DeviceLoader:
        self._rank, self._world, self._cursor = int(rank), max(int(world), 1), 0
        return DeviceLoader(self, batch_size, device, world)
