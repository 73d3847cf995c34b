"""The reference's entry points on the B200 backend: ``scripts/training_ddp.py`` trains the shipped spherical-harmonics
config (examples/config_spherical_harmonics.json, identical to the reference's file up to the device strings) for a few
steps and writes a reference-layout checkpoint; ``scripts/inference.py`` loads it strictly into the
GaussianDiffusionPipeline and generates."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import ROOT  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def run(cmd, **kw):
    res = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, **kw)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return res.stdout


def test_training_script_then_inference_script(tmp_path):
    cfg = os.path.join(ROOT, "examples", "config_spherical_harmonics.json")
    ckpt = str(tmp_path / "model.pth")
    out = run(["scripts/training_ddp.py", cfg, "-d", "cuda", "-e", "1", "--max-steps", "4", "--num-workers", "2", "--save", ckpt])
    losses = [float(line.split("loss: ")[1].split(",")[0]) for line in out.splitlines() if "loss: " in line]
    assert losses and all(0.0 < v < 5.0 for v in losses), out[-1500:]
    sd = torch.load(ckpt)
    assert "input_blocks.0.0.weight" in sd and sd["input_blocks.0.0.weight"].shape == (64, 1, 3, 3, 3)
    assert all(v.dtype == torch.float32 for v in sd.values())
    gen = str(tmp_path / "gen.pt")
    out = run(["scripts/inference.py", cfg, "-p", ckpt, "-d", "cuda", "-n", "3", "--steps", "6", "--out", gen, "-f", "--cache", str(tmp_path / "c.h5")])
    x = torch.load(gen)
    assert x.shape == (3, 1, 32, 32, 32) and torch.isfinite(x).all(), out[-1500:]


def test_inference_script_on_the_deep_galaxy_config(tmp_path):
    """examples/config_deep_galaxy.json (2-D 128^2, MultiEmbeddings over DeepGalaxyDataset's class-level parameter space): a checkpoint in
    the reference's layout (tests/golden/deep_galaxy.pt manifest, synthetic values) loads strictly and generate() samples the first
    rows of the inference parameter space."""
    from conftest import load_golden
    from oracle.synth import synth_state_dict
    g = load_golden("deep_galaxy.pt")
    ckpt = str(tmp_path / "galaxy.pth")
    torch.save(synth_state_dict(g["manifest"], seed=777), ckpt)
    gen = str(tmp_path / "gen.pt")
    cfg = os.path.join(ROOT, "examples", "config_deep_galaxy.json")
    png = tmp_path / "grid.png"
    out = run(["scripts/inference.py", cfg, "-p", ckpt, "-d", "cuda", "-n", "4", "--steps", "4", "--out", gen, "--plot", str(png), "-f", "--cache", str(tmp_path / "c.h5")])
    x = torch.load(gen)
    assert x.shape == (4, 1, 128, 128) and torch.isfinite(x).all(), out[-1500:]
    assert png.read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"        # the grid of scripts/inference.py:172-176 (utils.plot_image_grid)
    assert float(x.abs().max()) <= 1.0 + 1e-6           # dynamic thresholding keeps samples in [-1, 1]


def test_training_script_builds_cond_fn_and_saves_the_ema_checkpoint(tmp_path):
    """examples/config_gaussian_blobs.json: `model.kwargs.cond_fn = "MultiEmbeddings"` (the mechanism of the DeepGalaxy config) over the
    dataset class's parameter space, batches generated on the device, EMA on: the script trains, writes model.pth and model_ema.pth in the
    reference layout (embedding tables included), and the inference script samples from the result."""
    cfg = os.path.join(ROOT, "examples", "config_gaussian_blobs.json")
    ckpt = str(tmp_path / "blobs.pth")
    out = run(["scripts/training_ddp.py", cfg, "-d", "cuda", "-e", "1", "--max-steps", "6", "--save", ckpt])
    losses = [float(line.split("loss: ")[1].split(",")[0]) for line in out.splitlines() if "loss: " in line]
    assert losses and all(0.0 < v < 5.0 for v in losses), out[-1500:]
    sd, ema = torch.load(ckpt), torch.load(str(tmp_path / "blobs_ema.pth"))
    assert sd.keys() == ema.keys() and "cond_fn.embedding_layers.n.weight" in sd and sd["cond_fn.embedding_layers.w.weight"].shape == (3, 128)
    assert any(not torch.equal(sd[k], ema[k]) for k in sd)          # the shadow lags the weights
    gen = str(tmp_path / "gen.pt")
    out = run(["scripts/inference.py", cfg, "-p", ckpt, "-d", "cuda", "-n", "5", "--steps", "5", "--out", gen, "--plot", str(tmp_path / "blobs.png"),
               "-f", "--cache", str(tmp_path / "c.h5")])
    assert (tmp_path / "blobs.png").stat().st_size > 0
    # a second call without -f finds the cache (HDF5 `data`, or <file>.npz without h5py), plots it and does not sample again
    (tmp_path / "blobs.png").unlink()
    out = run(["scripts/inference.py", cfg, "-p", ckpt, "-d", "cuda", "--out", str(tmp_path / "never.pt"), "--plot", str(tmp_path / "blobs.png"),
               "--cache", str(tmp_path / "c.h5")])
    assert "Found a cache copy" in out and (tmp_path / "blobs.png").stat().st_size > 0 and not (tmp_path / "never.pt").exists()
    x = torch.load(gen)
    assert x.shape == (5, 1, 64, 64) and torch.isfinite(x).all(), out[-1500:]
    # the same script on the host data path (DataLoader workers, pinned copies)
    run(["scripts/training_ddp.py", cfg, "-d", "cuda", "-e", "1", "--max-steps", "2", "--num-workers", "2", "--host-data", "--save", ""])
