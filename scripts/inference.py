#!/usr/bin/env python
"""Sampling from a trained checkpoint -- the B200 counterpart of the reference's ``scripts/inference.py``: builds the
``GaussianDiffusionPipeline`` exactly as the reference does (lines 122-143), loads ``model.pth`` strictly, and calls
``generate(parameter_space=config.inference.parameter_space)`` (DDIM eta = 0 with dynamic thresholding, one CUDA graph
per step).  The result is written with ``torch.save`` (``--out``) and, when h5py is importable, to
``config.inference.cache_file`` under the key ``data`` like the reference (``<cache_file>.npz`` without h5py); 2-D results are also
rendered to ``config.inference.plot_output_file`` (``rho_diffusion.utils.plot_image_grid``).  As in the reference, an existing cache is
plotted instead of sampling again unless ``-f`` is given.

    python scripts/inference.py examples/config_spherical_harmonics.json -p model.pth -d cuda -n 10
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rho-diffusion_b200"))

import torch  # noqa: E402

from rho_diffusion import diffusion  # noqa: E402
from rho_diffusion.config import ExperimentConfig  # noqa: E402
from rho_diffusion.registry import registry  # noqa: E402


def build_pipeline(config, device, n_samples: int, steps: int = 0):
    """The GaussianDiffusionPipeline of ``config`` as the reference's inference script assembles it (scripts/inference.py:122-143).
    Returns (pipeline, cond_fn_kwargs or None)."""
    schedule = registry.get("schedules", config.noise_schedule.name)(device=device, **config.noise_schedule.kwargs)
    parameter_space = registry.get("datasets", config.dataset.name).parameter_space
    cond_fn = config.model.kwargs.get("cond_fn")
    # the shipped spherical-harmonics labels are same-shape sha512 embeddings (added to the time embedding directly,
    # unet_v2.py:705-712); MultiEmbeddings is only built when the dataset class exposes a discrete parameter space
    cond_kwargs = {"parameter_space": parameter_space, "embedding_dim": 128} if (cond_fn and parameter_space is not None) else None
    model = diffusion.GaussianDiffusionPipeline(
        backbone=config.model.name, backbone_kwargs=config.model.kwargs, schedule=schedule, loss_func=config.training.loss_fn,
        timesteps=steps or config.noise_schedule.kwargs["num_steps"], cond_fn=cond_fn if cond_kwargs else None,
        cond_fn_kwargs=cond_kwargs, optimizer=config.optimizer.name, opt_kwargs=config.optimizer.kwargs,
        sample_every_n_epochs=config.training.sample_every_n_epochs, sampling_batch_size=n_samples,
        sample_parameter_space=config.inference.parameter_space)
    return model, cond_kwargs


def read_cache(path):
    """The cached samples of an earlier run (scripts/inference.py:94-107 of the reference: HDF5, key ``data``); ``<path>.npz`` is the
    h5py-free fallback this script writes when h5py is not importable.  None when there is no cache."""
    if os.path.isfile(path):
        try:
            import h5py
            with h5py.File(path, "r") as h5f:
                return torch.tensor(h5f["data"][()])
        except ImportError:
            pass
    if os.path.isfile(str(path) + ".npz"):
        import numpy as np
        with np.load(str(path) + ".npz") as z:
            return torch.from_numpy(z["data"])
    return None


def write_cache(path, pred) -> str:
    try:
        import h5py
        with h5py.File(path, "w") as h5f:
            h5f["data"] = pred.cpu().numpy()
        return str(path)
    except ImportError:
        import numpy as np
        np.savez(str(path) + ".npz", data=pred.cpu().numpy())
        return str(path) + ".npz"


def plot(pred, plot_file) -> None:
    # scripts/inference.py:103-107, 172-176 of the reference (2-D fields; volumes and profiles have no grid rendering there either)
    if plot_file is not None and pred.ndim in (4, 5) and pred.shape[-3] in (1, 3):
        from rho_diffusion.utils import plot_image_grid
        plot_image_grid(pred, transpose=False, filename=plot_file)
        print("grid ->", plot_file)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("json_config", type=Path)
    parser.add_argument("-p", dest="model_checkpoint_path", type=Path, required=False)
    parser.add_argument("-d", "--device", dest="device", type=str, default="cuda")
    parser.add_argument("-n", dest="n_samples", type=int, default=10)
    parser.add_argument("-f", dest="forced_overwrite", action="store_true", default=False)
    parser.add_argument("--out", default="inference_output.pt")
    parser.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--steps", type=int, default=0, help="override the number of diffusion steps (0 = config)")
    parser.add_argument("--plot", default=None, help="PNG grid of a 2-D result (default: config.inference.plot_output_file)")
    parser.add_argument("--cache", default=None, help="sample cache file (default: config.inference.cache_file)")
    args = parser.parse_args()

    config = ExperimentConfig.from_json(args.json_config)
    plot_file = args.plot or config.inference.plot_output_file
    cache = args.cache or config.inference.cache_file
    if cache and not args.forced_overwrite:
        cached = read_cache(cache)
        if cached is not None:
            print("Found a cache copy of the generated data: %s. Will plot this file. If you wish to plot newly generated data, "
                  "please use the `-f` argument to overwrite." % cache)
            plot(cached, plot_file)
            return

    from rho_diffusion import b200 as accel

    torch.manual_seed(config.inference.seed)
    device = torch.device(args.device)
    model, cond_kwargs = build_pipeline(config, device, args.n_samples, args.steps)
    ckpt = args.model_checkpoint_path or config.inference.checkpoint
    model.backbone.load_state_dict(torch.load(ckpt), strict=True)
    model.eval()
    model.to(device)
    model = accel.optimize(model, dtype=torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    if cond_kwargs is not None:
        pred = model.generate(parameter_space=config.inference.parameter_space, random=False)
    else:
        # same-shape label embeddings (the convention of SphericalHarmonicDataset.__getitem__, synthetic.py:299-300): one
        # sha512 embedding per point of the inference parameter space, batch-sharded across ranks like generate()
        import itertools
        from rho_diffusion.utils import calculate_sha512_embedding
        keys = list(config.inference.parameter_space.keys())
        combos = list(itertools.product(*config.inference.parameter_space.values()))[: args.n_samples]
        emb_dim = 4 * int(config.model.kwargs["model_channels"])
        cond = torch.stack([calculate_sha512_embedding({k: int(v) for k, v in zip(keys, c)}, l=emb_dim) for c in combos]).to(device)
        shape = [len(combos), int(config.model.kwargs["out_channels"])] + [int(s) for s in config.model.kwargs["data_shape"]]
        pred = model.reverse_process(torch.zeros(shape, device=device), conditions=cond)["denoised"]
    torch.save(pred.cpu(), args.out)
    print("generated", tuple(pred.shape), "mean %.4f std %.4f ->" % (float(pred.mean()), float(pred.std())), args.out)
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0:     # (sharded runs: rank 0's slice)
        plot(pred, plot_file)
        if cache:
            print("cache ->", write_cache(cache, pred))


if __name__ == "__main__":
    main()
