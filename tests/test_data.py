"""Datasets of the hot path's callers: host-side logic on CPU, device generators against the host
restatements on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401


def _fake_deep_galaxy(path, seed=0):
    rng = np.random.default_rng(seed)
    arc = {}
    for s_, m_ in ((1, 1), (0.5, 1), (1.5, 0.25)):
        g = f"s_{s_}_m_{m_}"
        for c in (0, 3):
            arc[f"{g}/images_camera_{c:02d}"] = (rng.random((6, 288, 300, 1)) * 200).astype(np.float32)
            arc[f"{g}/t_myr_camera_{c:02d}"] = np.arange(390, 420, 5)
    np.savez(path, **arc)
    return arc


def test_deep_galaxy_reader_restates_the_reference_pipeline(tmp_path):
    """Groups / cameras / t_lim selection, per-set max normalisation, swapaxes(1, 3), float label rows (s, m, t, c), loaded parameter
    space, CenterCrop(256) -> Resize(128) -> 2x - 1 (identical to the torchvision transform the reference composes)."""
    from rho_diffusion.data import DeepGalaxyDataset
    path = str(tmp_path / "dg.npz")
    arc = _fake_deep_galaxy(path)
    d = DeepGalaxyDataset(path, use_emb_as_labels=False, dset_name_pattern="s_*", camera_pos=[0, 3], t_lim=[400, 520])
    assert len(d) == 3 * 2 * 4                                     # t in {400, 405, 410, 415} of 6 snapshots
    img, lab = d[0]
    assert img.shape == (1, 128, 128) and lab.shape == (4,) and -1.0 <= float(img.min()) and float(img.max()) <= 1.0
    assert d.loaded_parameter_space["t"] == [400, 405, 410, 415] and d.loaded_parameter_space["c"] == [0, 3]
    assert d.loaded_parameter_space["s"] == [0.5, 1.0, 1.5] and d.loaded_parameter_space["m"] == [0.25, 1.0]
    rows = {tuple(r.tolist()) for r in d.labels}
    assert (1.5, 0.25, 400.0, 3.0) in rows and (0.5, 1.0, 415.0, 0.0) in rows
    # every label value is a point of the class-level parameter space the MultiEmbeddings tables are built over
    for i, key in enumerate(("s", "m", "t", "c")):
        assert set(d.labels[:, i].tolist()) <= {float(v) for v in DeepGalaxyDataset.parameter_space[key]}
    raw = arc["s_0.5_m_1/images_camera_00"]
    first_of_group = [i for i, r in enumerate(d.labels) if tuple(r.tolist()) == (0.5, 1.0, 400.0, 0.0)][0]
    want = torch.from_numpy((raw / raw.max())[2]).permute(2, 1, 0)    # snapshot t = 400 is index 2; (H, W, C) -> (C, W, H)
    assert torch.equal(d.data[first_of_group], want)
    import torchvision.transforms as T
    tv = T.Compose([T.CenterCrop(256), T.Resize((128, 128)), T.Lambda(lambda t: (t * 2) - 1)])
    assert float((tv(d.data[5]) - d.transform(d.data[5])).abs().max()) < 1e-6
    with pytest.raises(ValueError):
        DeepGalaxyDataset(path, dset_name_pattern="nothing_*")


def test_gaussian_blob_dataset_host_path():
    from rho_diffusion.data import GaussianBlobDataset
    for shape in (64, [32, 32], [8, 16, 16]):
        d = GaussianBlobDataset(shape, length=16, random_seed=11)
        x, lab = d[3]
        x2, lab2 = d[3]
        dims = [shape] if isinstance(shape, int) else shape
        assert x.shape == (1, *dims) and torch.equal(x, x2) and torch.equal(lab, lab2)       # pure function of (seed, index)
        assert float(x.min()) >= -1.0 and float(x.max()) <= 1.0
        assert int(lab[0]) in GaussianBlobDataset.parameter_space["n"] and float(lab[1]) in [float(np.float32(w)) for w in (0.05, 0.1, 0.2)]
        assert not torch.equal(d[3][0], d[4][0])
    # wider blobs, more blobs -> larger mean density: the moments are a function of the label
    d = GaussianBlobDataset([32, 32], length=400, random_seed=5)
    means = {}
    for i in range(400):
        x, lab = d[i]
        means.setdefault((int(lab[0]), round(float(lab[1]), 3)), []).append(float(x.mean()))
    avg = {k: sum(v) / len(v) for k, v in means.items()}
    assert avg[(4, 0.2)] > avg[(1, 0.2)] > avg[(1, 0.05)]


def test_host_philox_matches_the_known_answer():
    """Random123 known-answer vector of Philox4x32-10 (the generator the device kernels and their host restatements share)."""
    from rho_diffusion.data.synthetic import _philox4x32_10
    assert _philox4x32_10(0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    ctr = 0xffffffffffffffff
    assert _philox4x32_10(ctr, ctr, ctr) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


# ---- device generators ----------------------------------------------------------------------------------------------
def _spectroscopy_dataset(tmp_path, **kw):
    from oracle.cases import spectroscopy_records
    from rho_diffusion.data import SpectroscopyDataset
    path = os.path.join(str(tmp_path), "records.npz")
    np.savez(path, **spectroscopy_records())
    return SpectroscopyDataset(path, **kw)


def test_spectroscopy_dataset_matches_the_reference_items(tmp_path):
    """SpectroscopyDataset.__getitem__ / simulate_lineprofile (spectroscopy.py:103-189) against items of the UNMODIFIED reference
    on the same records and the same `random.gauss` line widths (tests/golden/spectroscopy.pt): bit-exact."""
    import random
    from conftest import load_golden
    from rho_diffusion.registry import registry
    g = load_golden("spectroscopy.pt")
    ds = _spectroscopy_dataset(tmp_path, grid_size=4096)
    assert registry.get("datasets", "SpectroscopyDataset") is type(ds) and len(ds) == 3
    assert torch.equal(torch.from_numpy(ds.frequency_grid), g["frequency_grid"])
    for i in range(3):
        random.seed(1000 + i)
        item = ds[i]
        assert set(item) == set(g[i])
        for k, v in g[i].items():
            assert item[k].dtype == torch.float32 and torch.equal(item[k], v), (i, k)
        assert item["spectrum"].shape == (1, 4096) and float(item["spectrum"].max()) == 1.0
    with pytest.raises(AssertionError):
        type(ds)(os.path.join(str(tmp_path), "missing.npz"))


gpu = pytest.mark.gpu


@gpu
def test_device_line_profiles_match_the_host_restatement(tmp_path):
    """rho_lineprofile (one ragged-batch kernel) against the host path on the same records and widths, at the fixture's grid and at
    the reference's default 50 000-channel grid."""
    import random
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    for grid_size in (4096, 50_000):
        ds = _spectroscopy_dataset(tmp_path, grid_size=grid_size)
        random.seed(7)
        host = [ds[i] for i in (0, 1, 2, 1)]
        random.seed(7)
        batch = ds.device_batch(4, "cuda", indices=[0, 1, 2, 1])
        assert batch["spectrum"].shape == (4, 1, grid_size) and batch["spectrum"].dtype == torch.float32
        for j, item in enumerate(host):
            ref, got = item["spectrum"], batch["spectrum"][j].cpu()
            assert float((got - ref).abs().max()) < 2e-6, (grid_size, j, float((got - ref).abs().max()))
            assert torch.equal(batch["max_int"][j].cpu(), item["max_int"])
            for k in ("A", "B", "C"):
                assert torch.equal(batch[k][j].cpu(), item[k])


@gpu
def test_device_spherical_harmonics_match_the_host_restatement():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rho_diffusion.b200 import ops
    from rho_diffusion.data import SphericalHarmonicDataset, compute_spherical_harmonic
    dev = torch.device("cuda", 0)
    x = np.linspace(-2.0, 2.0, 32)
    lm = [(l, m) for l in range(0, 7) for m in range(-l, l + 1)]
    helper = SphericalHarmonicDataset(max_l=6, length=1, random_seed=1)
    mm = torch.tensor([helper._extremes(l, m) for l, m in lm], dtype=torch.float64, device=dev)
    lt, mt = torch.tensor([v[0] for v in lm], device=dev), torch.tensor([v[1] for v in lm], device=dev)
    got = ops.sph_harm_density(lt, mt, 32, minmax=mm).cpu()
    own = ops.sph_harm_density(lt, mt, 32).cpu()                 # the kernel's own lexicographic extremes
    agree = 0
    for i, (l, m) in enumerate(lm):
        _, density, _ = compute_spherical_harmonic(x, x, x, m=m, l=l)
        want = torch.from_numpy(density.astype(np.float32))
        assert float((got[i, 0] - want).abs().max()) < 2e-6, (l, m)
        agree += int(float((own[i, 0] - want).abs().max()) < 2e-6)
    assert agree >= len(lm) // 2          # where no near-tie decides the extremes the self-contained kernel agrees too (m = 0, ...)
    # the device loader consumes the same (l, m) stream as __getitem__
    host, devd = SphericalHarmonicDataset(max_l=5, length=8, random_seed=21), SphericalHarmonicDataset(max_l=5, length=8, random_seed=21)
    xs, embs = devd.device_batch(4, dev)
    for i in range(4):
        hx, he = host[i]
        both_finite = torch.isfinite(hx) & torch.isfinite(xs[i].cpu())
        assert float((hx - xs[i].cpu())[both_finite].abs().max()) < 2e-6 and torch.equal(he, embs[i].cpu())


@gpu
def test_device_gaussian_blobs_match_the_host_restatement():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rho_diffusion.data import GaussianBlobDataset
    dev = torch.device("cuda", 0)
    for shape in (128, [48, 40], [12, 20, 16]):
        d = GaussianBlobDataset(shape, length=64, random_seed=9)
        loader = d.device_loader(8, dev)
        xs, labels = next(iter(loader))
        assert xs.shape[0] == 8 and labels.shape == (8, 2)
        for i in range(8):
            hx, hl = d[i]
            assert torch.equal(hl, labels[i].cpu())
            assert float((hx - xs[i].cpu()).abs().max()) < 2e-5, (shape, i)
    # two ranks draw disjoint index ranges
    a = GaussianBlobDataset([16, 16], length=64, random_seed=9).device_loader(4, dev, rank=0, world=2)
    b = GaussianBlobDataset([16, 16], length=64, random_seed=9).device_loader(4, dev, rank=1, world=2)
    xa, xb = next(iter(a))[0], next(iter(b))[0]
    ref = GaussianBlobDataset([16, 16], length=64, random_seed=9)
    assert float((xa[0].cpu() - ref[0][0]).abs().max()) < 2e-5 and float((xb[0].cpu() - ref[4][0]).abs().max()) < 2e-5
