"""DeepGalaxy (reference ``rho_diffusion/data/deep_galaxy.py``): galaxy-merger density maps labelled by size ratio ``s``, mass
ratio ``m``, time ``t`` (Myr) and camera position ``c``.

* The CLASS-LEVEL discrete parameter space (deep_galaxy.py:41-47) is what ``scripts/inference.py:122-143`` hands to
  ``MultiEmbeddings``: it fixes the row counts of the four embedding tables of every DeepGalaxy checkpoint (pinned to the
  reference by ``tests/golden/deep_galaxy.pt``).
* The reader restates ``_load_all`` / ``_load_dataset`` / ``_get_labels`` / ``load`` / ``__getitem__`` (deep_galaxy.py:84-93,
  118-132, 165-318): groups ``s_<s>_m_<m>`` matching ``dset_name_pattern``, datasets ``images_camera_%02d`` (N, H, W, C) and
  ``t_myr_camera_%02d`` (N,), each image set divided by its own maximum, optional ``t_lim`` window, ``swapaxes(1, 3)`` to
  (N, C, W, H), labels = float rows ``(s, m, t, c)``; default transform ``CenterCrop(256) -> Resize(128, 128) -> 2x - 1``.
  Reference quirk kept: ``_get_labels`` returns ``(m, s, t)`` and ``_load_all`` unpacks them as ``labels_m, labels_s, labels_t``.
* ``h5py`` is an optional dependency (absent from this image): an ``.h5`` / ``.hdf5`` path needs it; an ``.npz`` archive with the
  same paths as keys (``"s_1_m_1/images_camera_00"`` ...) is read with numpy alone, which is what the tests use.
"""
from __future__ import annotations

import re
from collections import OrderedDict

import numpy as np
import torch

from rho_diffusion.data.parameter_space import DiscreteParameterSpace
from rho_diffusion.registry import registry


class _NpzArchive:
    """Mapping view of an ``.npz`` file with HDF5-style paths: ``keys()`` lists the top-level groups."""

    def __init__(self, path):
        self._z = np.load(path)

    def keys(self):
        return sorted({k.strip("/").split("/")[0] for k in self._z.files})

    def __getitem__(self, h5_path):
        return self._z[h5_path.strip("/")]


def _open_archive(path):
    if str(path).endswith(".npz"):
        return _NpzArchive(path)
    try:
        import h5py
    except ImportError as exc:  # pragma: no cover (h5py is not part of this image)
        raise ImportError("DeepGalaxyDataset: reading an HDF5 file needs h5py (optional dependency); an .npz archive with the same "
                          "paths as keys is read without it") from exc
    return h5py.File(path, "r")


def default_transform():
    """``CenterCrop(256) -> Resize((128, 128)) -> 2x - 1`` (deep_galaxy.py:84-93) on (C, H, W) float tensors."""
    import torch.nn.functional as F

    def fn(img: torch.Tensor) -> torch.Tensor:
        h, w = img.shape[-2:]
        if h < 256 or w < 256:                                   # torchvision's CenterCrop zero-pads smaller images
            ph, pw = max(256 - h, 0), max(256 - w, 0)
            img = F.pad(img, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
            h, w = img.shape[-2:]
        top, left = int(round((h - 256) / 2.0)), int(round((w - 256) / 2.0))
        img = img[..., top:top + 256, left:left + 256]
        img = F.interpolate(img[None], size=(128, 128), mode="bilinear", antialias=True, align_corners=False)[0]
        return img * 2 - 1
    return fn


@registry.register_dataset("DeepGalaxyDataset")
class DeepGalaxyDataset(torch.utils.data.Dataset):
    parameter_space = DiscreteParameterSpace(param_dict=OrderedDict([
        ("s", [0.25 * i for i in range(1, 7)]),            # 0.25 ... 1.5
        ("m", [0.25 * i for i in range(1, 7)]),
        ("t", list(range(300, 655, 5))),                   # 71 snapshots
        ("c", list(range(14))),                            # camera positions
    ]))

    def __init__(self, path: str, use_emb_as_labels: bool = True, dset_name_pattern: str = "s_*", camera_pos=(0,), t_lim=None,
                 transform=None, target_transform=None):
        self.h5fn = path
        self.use_emb_labels = use_emb_as_labels
        self.data = None
        self.labels = None
        self.num_classes = 0
        self.dset_name_pattern = (dset_name_pattern,)
        self.camera_pos = camera_pos
        self.t_lim = t_lim
        self.loaded_parameter_space = DiscreteParameterSpace(param_dict={"s": [], "m": [], "t": [], "c": []})
        self.attributes = ["s", "m", "t", "c"]
        self.transform = default_transform() if transform is None else transform
        self.target_transform = target_transform
        self.load(dset_name_pattern, camera_pos, t_lim)

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, idx):
        image, label = self.data[idx], self.labels[idx]
        if self.transform:
            image = self.transform(image)
        if self.target_transform:
            label = self.target_transform(label)
        return image, label

    def load(self, dset_name_pattern, camera_pos, t_lim):
        data, labels = self._load_all(dset_name_pattern=dset_name_pattern, camera_pos=camera_pos, t_lim=t_lim)
        self.data = torch.as_tensor(np.asarray(data).swapaxes(1, 3), dtype=torch.float32)     # (N, H, W, C) -> (N, C, W, H)
        self.labels = labels

    def _load_all(self, dset_name_pattern="*", camera_pos="*", t_lim=None):
        self.h5f = _open_archive(self.h5fn)
        r = re.compile(dset_name_pattern)
        matched = list(filter(r.match, self.h5f.keys()))
        if isinstance(camera_pos, int):
            camera_pos = [camera_pos]
        elif isinstance(camera_pos, str) and camera_pos == "*":
            camera_pos = range(0, 14)
        images_set, m_set, s_set, t_set, c_set = [], [], [], [], []
        for dset_name in matched:
            for cpos in camera_pos:
                images = self._load_dataset("/%s/images_camera_%02d" % (dset_name, cpos))
                labels_m, labels_s, labels_t = self._get_labels(dset_name, cpos)
                labels_c = np.ones(labels_m.shape, dtype=np.int32) * cpos
                if t_lim is not None:
                    flags = np.logical_and(labels_t >= np.min(t_lim), labels_t <= np.max(t_lim))
                    images, labels_t, labels_m, labels_s, labels_c = images[flags], labels_t[flags], labels_m[flags], labels_s[flags], labels_c[flags]
                images_set.append(images); m_set.append(labels_m); s_set.append(labels_s); t_set.append(labels_t); c_set.append(labels_c)
        if not images_set:
            raise ValueError(f"DeepGalaxyDataset: no group of {self.h5fn} matches {dset_name_pattern!r}")
        images_set = np.concatenate(images_set, axis=0)
        loaded = {"m": np.concatenate(m_set), "s": np.concatenate(s_set), "t": np.concatenate(t_set), "c": np.concatenate(c_set)}
        for key in self.attributes:
            self.loaded_parameter_space[key] = sorted(np.unique(loaded[key]).tolist())
        self.num_classes = int(np.unique(loaded["t"]).shape[0])
        labels = torch.zeros((len(loaded["m"]), len(self.attributes)), dtype=torch.float)
        for i, attr in enumerate(self.attributes):
            labels[:, i] = torch.tensor(np.asarray(loaded[attr]), dtype=torch.float)
        return images_set, labels

    def _load_dataset(self, h5_path):
        images = np.asarray(self.h5f[h5_path][()])
        return images / np.max(images)

    def _get_labels(self, dset_name, camera_pos=0):
        s = float(dset_name.split("_")[1])
        m = float(dset_name.split("_")[3])
        cat_t = np.asarray(self.h5f["%s/t_myr_camera_%02d" % (dset_name, camera_pos)][()])
        cat_s = np.array([s] * cat_t.shape[0])
        cat_m = np.array([m] * cat_t.shape[0])
        return cat_m, cat_s, cat_t
