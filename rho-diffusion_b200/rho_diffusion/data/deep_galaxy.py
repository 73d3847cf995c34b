"""DeepGalaxy (reference ``rho_diffusion/data/deep_galaxy.py``).

What the hot path needs from this dataset is its CLASS-LEVEL discrete parameter space (deep_galaxy.py:41-47): the scripts pass it
to ``MultiEmbeddings`` (``scripts/inference.py:122-143``), so it fixes the row counts of the four embedding tables
(size ratio s, mass ratio m, time t, camera position c) of every DeepGalaxy checkpoint.  The HDF5 reader
(``CenterCrop(256) -> Resize(128) -> 2x - 1`` over ``output_bw_512.h5``) is outside this backend's scope (h5py is not part of the
image; DESIGN.md, out of scope): constructing the dataset raises, the class attribute is all the sampling path touches.
Pinned to the reference by ``tests/golden/deep_galaxy.pt`` (``oracle/make_golden.py::gen_deep_galaxy``)."""
from __future__ import annotations

from collections import OrderedDict

import torch

from rho_diffusion.data.parameter_space import DiscreteParameterSpace
from rho_diffusion.registry import registry


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
        raise NotImplementedError(
            "DeepGalaxyDataset: the HDF5 reader is not part of the B200 hot-path backend (h5py is absent from this image); "
            "DeepGalaxyDataset.parameter_space (class attribute) is available for checkpoints / sampling")
