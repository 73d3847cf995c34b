"""Rotational-spectrum dataset (reference ``rho_diffusion/data/spectroscopy.py:35-189``): the 1-D testbed -- every record holds the
transitions (centre frequencies, log10 intensities) of one molecule plus the parameters they were generated from; an item is the
spectrum simulated on a fixed frequency grid with Gaussian line shapes, divided by its maximum.

* Host path = the reference's arithmetic restated (``__getitem__`` :103-133, ``simulate_lineprofile`` :135-189): intensities clipped
  to [-10, -2] (log10), ``width ~ |N(1, 0.1)|`` per item from Python's ``random.gauss``, numpy promotion rules (fp32 difference and
  square, fp64 division / exp / sum), ``float32`` tensors out, a leading channel dimension on ``spectrum``.
* Device path (``device_batch`` / ``simulate_lineprofile_device``): the same sum of Gaussians as ONE kernel over a ragged batch of
  records (``rho_lineprofile``), so a 1-D training run keeps its input pipeline off the host.
* ``h5py`` is optional (absent from this image): an ``.h5`` / ``.hdf5`` path needs it; an ``.npz`` archive whose keys are
  ``"<index>/<name>"`` (``"0/transitions"``, ``"0/A"`` ...) is read with numpy alone, which is what the tests use.

Reference quirks kept: ``__len__`` is a ``cached_property`` there (so ``len(dataset)`` raises ``TypeError``); here it is a method
-- ``len()`` works -- and the record count is also available as ``num_records``; the ``linewidth`` constructor argument is stored but
``__getitem__`` always draws ``gauss(1.0, 0.1)`` (spectroscopy.py:110), which is what this class does too.
"""
from __future__ import annotations

from pathlib import Path
from random import gauss

import numpy as np
import torch
from torch import Tensor
from torch.utils.data import Dataset

from rho_diffusion.registry import registry

__all__ = ["SpectroscopyDataset"]


class _NpzRecords:
    """``{"<index>/<name>": array}`` archive with the interface of the HDF5 file the reference reads (groups keyed by str(index))."""

    def __init__(self, path):
        self._z = np.load(path)
        self._groups = {}
        for k in self._z.files:
            g, _, name = k.strip("/").partition("/")
            self._groups.setdefault(g, {})[name] = k

    def __len__(self):
        return len(self._groups)

    def __getitem__(self, group):
        return {name: self._z[k] for name, k in self._groups[group].items()}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _open_records(path):
    if str(path).endswith(".npz"):
        return _NpzRecords(path)
    try:
        from h5py import File
    except ImportError as exc:  # pragma: no cover (h5py is not part of this image)
        raise ImportError("SpectroscopyDataset: reading an HDF5 file needs h5py (optional dependency); an .npz archive with "
                          "'<index>/<name>' keys is read without it") from exc
    return File(str(path), mode="r")


@registry.register_dataset("SpectroscopyDataset")
class SpectroscopyDataset(Dataset):
    def __init__(self, h5_path, min_freq: float | None = None, max_freq: float | None = None, grid_size: int = 50_000,
                 linewidth: float | tuple | None = None) -> None:
        super().__init__()
        self.h5_path = h5_path
        self.min_freq = min_freq
        self.max_freq = max_freq
        self.grid_size = grid_size
        self.linewidth = linewidth
        self._grid = None
        self._len = None
        self._dev_cache = {}

    # -- constructor arguments with the reference's validation (spectroscopy.py:60-95) -----------------------------------------
    @property
    def linewidth(self) -> float:
        linewidth = self._linewidth
        if not isinstance(linewidth, float):
            mu, sigma = linewidth if isinstance(linewidth, tuple) else (1.0, 0.1)
            return np.abs(gauss(mu, sigma))
        return linewidth

    @linewidth.setter
    def linewidth(self, value) -> None:
        if isinstance(value, tuple):
            assert len(value) == 2, "Expected two-tuple for linewidth specification."
        self._linewidth = value

    @property
    def h5_path(self) -> Path:
        return self._h5_path

    @h5_path.setter
    def h5_path(self, value) -> None:
        if isinstance(value, str):
            value = Path(value)
        assert value.exists(), f"Target HDF5 not found; passed {value}."
        self._h5_path = value

    @property
    def h5_data(self):
        return _open_records(self.h5_path)

    @property
    def num_records(self) -> int:
        if self._len is None:
            with self.h5_data as h5_data:
                self._len = len(h5_data)
        return self._len

    def __len__(self) -> int:
        return self.num_records

    @property
    def frequency_grid(self) -> np.ndarray:
        """``np.linspace(min_freq or 1000, max_freq or 32000, grid_size, dtype=float32)`` (spectroscopy.py:91-101)."""
        if self._grid is None:
            self._grid = np.linspace(self.min_freq or 1000, self.max_freq or 32000, self.grid_size, dtype=np.float32)
        return self._grid

    # -- host path ------------------------------------------------------------------------------------------------------------
    def _record(self, index) -> dict:
        with self.h5_data as h5_data:
            group = h5_data[str(index)]
            return {key: np.array(value).astype(np.float32) for key, value in group.items()}

    def __getitem__(self, index) -> dict:
        data = self._record(index)
        width = np.abs(gauss(1.0, 0.1))
        centers, intensities = data["transitions"]
        intensities = np.clip(intensities, -10.0, -2.0)                     # log10; some may have underflowed
        lineprofile = self.simulate_lineprofile(self.frequency_grid, centers, 10 ** intensities, width)
        max_int = 10 ** intensities.max()
        lineprofile /= lineprofile.max()
        del data["transitions"]
        data["spectrum"] = Tensor(lineprofile)[None, :]                     # channel dimension
        data["max_int"] = Tensor([max_int])
        for key in data.keys():
            if not isinstance(data[key], Tensor):
                data[key] = Tensor(data[key]).float()
        return data

    @staticmethod
    def simulate_lineprofile(frequency_grid: np.ndarray, centers: np.ndarray, intensities: np.ndarray, width) -> np.ndarray:
        """Sum of Gaussian lines on ``frequency_grid`` (spectroscopy.py:135-189); transitions outside the grid are dropped."""
        if isinstance(width, float):
            width = np.array([width])
        min_freq, max_freq = frequency_grid.min(), frequency_grid.max()
        mask = np.all([centers <= max_freq, min_freq <= centers], axis=0)
        lineprofile = intensities[mask, None] * np.exp(-((frequency_grid[None, :] - centers[mask, None]) ** 2.0) / (2 * width[:, None] ** 2.0))
        return lineprofile.sum(axis=0)

    # -- device path -------------------------------------------------------------------------------------------------------------
    def simulate_lineprofile_device(self, records, widths, device) -> Tensor:
        """``records``: list of ``(centers, linear intensities)`` float32 arrays; ``widths``: one float per record.  Returns the
        max-normalised spectra ``(B, 1, grid_size)`` float32 on ``device`` from one ragged-batch kernel."""
        import ctypes as C  # noqa: F401
        from rho_diffusion.b200 import _lib as L, ops
        device = torch.device(device)
        key = (device.type, device.index)
        if key not in self._dev_cache:
            self._dev_cache[key] = torch.from_numpy(self.frequency_grid).to(device)
        grid = self._dev_cache[key]
        offs = np.zeros(len(records) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(c) for c, _ in records])
        cen = torch.from_numpy(np.concatenate([np.asarray(c, dtype=np.float32) for c, _ in records])).to(device)
        inten = torch.from_numpy(np.concatenate([np.asarray(i, dtype=np.float32) for _, i in records])).to(device)
        offsets = torch.from_numpy(offs).to(device)
        w = torch.tensor([float(v) for v in widths], dtype=torch.float64, device=device)
        B, G = len(records), int(grid.numel())
        out = torch.empty((B, 1, G), dtype=torch.float32, device=device)
        ws = ops.workspace(L.lib().rho_lineprofile_workspace_bytes(B, G), device)
        L.check(L.lib().rho_lineprofile(out.data_ptr(), B, G, grid.data_ptr(), cen.data_ptr(), inten.data_ptr(), offsets.data_ptr(), w.data_ptr(),
                                        ws.data_ptr(), ws.numel(), ops.stream()), "lineprofile")
        return out

    def device_batch(self, batch_size: int, device, indices=None) -> dict:
        """A batch of items assembled like ``__getitem__`` + default collation, the spectra simulated on ``device``."""
        if indices is None:
            indices = np.random.randint(0, len(self), size=batch_size)
        recs, widths, extras, max_int = [], [], {}, []
        for idx in indices:
            data = self._record(int(idx))
            widths.append(np.abs(gauss(1.0, 0.1)))
            centers, intensities = data.pop("transitions")
            intensities = np.clip(intensities, -10.0, -2.0)
            recs.append((centers, 10 ** intensities))
            max_int.append(10 ** intensities.max())
            for k, v in data.items():
                extras.setdefault(k, []).append(np.asarray(v, dtype=np.float32))
        batch = {k: torch.from_numpy(np.stack(v)).to(device) for k, v in extras.items()}
        batch["spectrum"] = self.simulate_lineprofile_device(recs, widths, device)
        batch["max_int"] = torch.tensor(max_int, dtype=torch.float32, device=device)[:, None]
        return batch
