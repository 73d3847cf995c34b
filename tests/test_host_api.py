"""CPU tests of the host-side mirror of the reference interface: registry, config, schedules, state-dict layout,
C-ABI symbol export.  No kernel is launched here."""
import ctypes
import json
import os
import re
from collections import OrderedDict

import pytest
import torch

from conftest import ROOT, load_golden
from oracle.cases import UNET_CASES

import rho_diffusion  # noqa: F401
from rho_diffusion.config import ConfigValidationError, ExperimentConfig
from rho_diffusion.diffusion import schedule as S
from rho_diffusion.models.conditioning import MultiEmbeddings
from rho_diffusion.models.unet_v2 import UNet
from rho_diffusion.registry import registry

EXAMPLES = {
    "sph": dict(experiment="ShpericalHarmonics",
                model={"name": "UNetv2", "kwargs": {"dims": 3, "in_channels": 1, "out_channels": 1, "model_channels": 64,
                                                    "num_res_blocks": 2, "data_shape": [32, 32, 32], "attention_resolutions": [16, 8],
                                                    "use_scale_shift_norm": True, "num_heads": 4, "num_classes": 20,
                                                    "activation": "SiLU", "use_new_attention_order": False,
                                                    "cond_fn": "MultiEmbeddings"}},
                dataset={"name": "SphericalHarmonicDataset", "kwargs": {"max_l": 5}},
                optimizer={"name": "AdamW", "kwargs": {"lr": 0.0001}},
                lr_scheduler={"name": "CosineAnnealingLR", "kwargs": {"T_max": 10, "eta_min": 0.00001}},
                noise_schedule={"name": "LinearSchedule", "kwargs": {"num_steps": 1000, "beta_1": 0.001, "beta_T": 0.02}},
                training={"device": "xpu", "np": 1, "loss_fn": "MSELoss", "batch_size": 32, "seed": 777, "benchmark_mode": True},
                inference={"device": "xpu", "checkpoint": "model.pth", "parameter_space": {"l": [-1, 1], "m": [1, 2]}}),
}


def test_registry_names_and_errors():
    # reference tests/models/test_unet.py:28-30 and tests/pipeline/test_schedule.py:49-53
    assert registry.get("models", "UNetv2") is UNet
    for name in S.__all__:
        assert registry.get("schedules", name)
    for cat, name in [("layers", "GroupNorm32"), ("layers", "conv_nd"), ("layers", "avg_pool_nd"), ("layers", "mean_flat"),
                      ("layers", "RMSNorm"), ("layers", "MultiEmbeddings"), ("layers", "SinusoidalPositionEmbedding"),
                      ("activations", "SiLU"), ("optimizers", "AdamW"), ("nn", "MSELoss")]:
        assert registry.get(cat, name)
    assert registry.categories == ["models", "activations", "layers", "datasets", "nn", "schedules", "optimizers"]
    with pytest.raises(KeyError):
        registry.get("models", "nope")
    with pytest.raises(AssertionError):
        registry.get("nope", "UNetv2")

    @registry.register_dataset("toy")
    class Toy:
        pass

    assert registry.get("datasets", "toy") is Toy


def test_config_schema_and_pydantic_v1_bool_quirk(tmp_path):
    p = tmp_path / "cfg.json"
    p.write_text(json.dumps(EXAMPLES["sph"]))
    cfg = ExperimentConfig.from_json(str(p))
    kw = cfg.model.kwargs
    # numbers survive as numbers, JSON booleans become *strings* under the pinned pydantic v1
    assert kw["model_channels"] == 64 and isinstance(kw["model_channels"], int)
    assert kw["attention_resolutions"] == [16, 8] and kw["data_shape"] == [32, 32, 32]
    assert kw["use_new_attention_order"] == "False" and kw["use_scale_shift_norm"] == "True"
    assert cfg.optimizer.kwargs["lr"] == pytest.approx(1e-4)
    assert cfg.noise_schedule.kwargs == {"num_steps": 1000, "beta_1": 0.001, "beta_T": 0.02}
    assert cfg.training.batch_size == 32 and cfg.training.ema_decay == 0.0 and cfg.training.max_epochs == 999
    assert not hasattr(cfg.training, "benchmark_mode")          # unknown keys are ignored
    assert cfg.inference.seed == 777 and cfg.inference.cache_file is None
    native = ExperimentConfig.from_json(str(p), bool_mode="native")
    assert native.model.kwargs["use_new_attention_order"] == 0 and not isinstance(native.model.kwargs["use_new_attention_order"], str)
    # the string 'False' is truthy => the new attention order is selected, as in the reference
    m = UNet(**{**kw, "model_channels": 32, "data_shape": [8, 8, 8]})
    assert type(m.middle_block[1].attention).__name__ == "QKVAttention"
    with pytest.raises(FileNotFoundError):
        ExperimentConfig.from_json(str(tmp_path / "missing.json"))


def test_config_invalid_is_rejected():
    # shape of the reference's tests/config_invalid.json: list where str expected, kwargs null, missing section
    bad = dict(EXAMPLES["sph"])
    bad["experiment"] = ["a", "b"]
    with pytest.raises(ConfigValidationError):
        ExperimentConfig(**bad)
    bad = dict(EXAMPLES["sph"])
    bad["model"] = {"name": "UNetv2", "kwargs": None}
    with pytest.raises(ConfigValidationError):
        ExperimentConfig(**bad)
    bad = dict(EXAMPLES["sph"])
    del bad["dataset"]
    with pytest.raises(ConfigValidationError):
        ExperimentConfig(**bad)


@pytest.mark.parametrize("name,make", [
    ("linear_100_kat", lambda: S.LinearSchedule(100, 1e-4, 0.02)),
    ("linear_1000", lambda: S.LinearSchedule(1000, 1e-3, 0.02)),
    ("linear_500_default", lambda: S.LinearSchedule(500)),
    ("cosine_50", lambda: S.CosineBetaSchedule(50)),
    ("cosine_1000", lambda: S.CosineBetaSchedule(1000)),
])
def test_schedule_tables_match_reference_bit_exact(name, make):
    g = load_golden("schedule.pt")[name]
    s = make()
    for k in ("beta_t", "alpha_t", "alpha_bar_t", "sigma_t", "offset_alpha_bar_t"):
        got = s[k] if k != "offset_alpha_bar_t" else s.offset_alpha_bar_t
        assert got.dtype == torch.float32
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(g[k], nan=-7.0)), k


def test_schedule_known_answers_and_dtype_cast():
    s = S.LinearSchedule(100, 1e-4, 0.02)  # reference tests/pipeline/test_schedule.py:28-46
    assert len(s.beta_t) == 100 and s.beta_t[0] == 0.001 and s.beta_t[-1] == 0.2
    assert s.alpha_t[0] == 0.999 and s.alpha_t[-1] == 0.8 and s.sigma_t[0] == 0.0
    assert float(s.sigma_t[-1]) == pytest.approx(0.4472, rel=1e-4)
    s.dtype = torch.bfloat16
    assert s["alpha_bar_t"].dtype == torch.bfloat16
    with s:
        assert s["alpha_bar_t"].dtype == torch.float32
    assert len(S.CosineBetaSchedule(50)) == 51           # T+1 points, sigma_t[0] NaN
    assert torch.isnan(S.CosineBetaSchedule(50).sigma_t[0])
    with pytest.raises(NotImplementedError):
        S.SigmoidSchedule(10)


@pytest.mark.parametrize("case", list(UNET_CASES))
def test_state_dict_layout_matches_reference(case):
    """Checkpoint contract: same keys, same shapes as the reference model => reference model.pth loads strict."""
    spec = UNET_CASES[case]
    g = load_golden(f"unet_{case}.pt")
    kwargs = dict(spec["kwargs"])
    if spec["cond"] == "multi":
        kwargs["cond_fn"] = MultiEmbeddings(parameter_space=OrderedDict(spec["parameter_space"]),
                                            embedding_dim=4 * kwargs["model_channels"])
    model = UNet(**kwargs)
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert list(got.keys()) == list(g["manifest"].keys())
    assert got == g["manifest"]
    # zero-initialised modules as in the reference (layers.py:113-119)
    assert float(model.out[2].weight.abs().sum()) == 0.0
    assert float(model.middle_block[1].proj_out.weight.abs().sum()) == 0.0
    assert float(model.middle_block[0].out_layers[3].weight.abs().sum()) == 0.0


def test_v1_unet_and_vit_have_the_reference_checkpoint_layout_and_registry_names():
    """`UNet`, `UNetBlock2d/3d`, `VisionTransformer` are registered under the reference's names and build exactly the
    reference's state-dict keys and shapes (tests/golden/backbones.pt holds the manifests of the unmodified reference modules)."""
    from conftest import load_golden
    from oracle.cases import BACKBONE_CASES
    from rho_diffusion.registry import registry
    g = load_golden("backbones.pt")
    assert registry.get("layers", "UNetBlock2d") and registry.get("layers", "UNetBlock3d")
    for case, spec in BACKBONE_CASES.items():
        model = registry.get("models", spec["model"])(**spec["kwargs"])
        assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == g[case]["manifest"], case
    with pytest.raises(Exception):            # no CPU path: running a layer without a CUDA device raises
        model(torch.zeros(BACKBONE_CASES[case]["x_shape"]), torch.zeros(BACKBONE_CASES[case]["x_shape"][0], dtype=torch.long))


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rho_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(rho_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    from rho_diffusion.b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first (__graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ set(declared)
    assert handle.rho_abi_version() == 1


def test_no_cpu_fallback():
    """The product must fail loudly off-GPU, never compute on the CPU."""
    from rho_diffusion.b200 import _lib
    m = UNet(data_shape=8, in_channels=1, model_channels=32, out_channels=1, num_res_blocks=1)
    with pytest.raises(_lib.RhoB200Error):
        m(torch.zeros(1, 1, 8, 8), torch.zeros(1, dtype=torch.long))


def test_spherical_harmonic_dataset_matches_reference():
    """SphericalHarmonicDataset (reference data/synthetic.py) -- fields, (l, m) draw sequence and label embeddings."""
    import numpy as np
    from conftest import load_golden
    from rho_diffusion.data.synthetic import SphericalHarmonicDataset, compute_spherical_harmonic
    from rho_diffusion.registry import registry
    from rho_diffusion.utils import calculate_sha512_embedding
    g = load_golden("sph_harm.pt")
    grid = np.linspace(-2.0, 2.0, 32)
    for (l, m), want in g["fields"].items():
        _, density, _ = compute_spherical_harmonic(grid, grid, grid, m=m, l=l)
        got = torch.from_numpy(density.astype(np.float32))
        assert torch.allclose(got[::3, ::3, ::3], want["density"], rtol=0, atol=1e-6), (l, m)
        stats = torch.tensor([density.min(), density.max(), density.mean(), density.std()], dtype=torch.float64)
        assert torch.allclose(stats, want["stats"], rtol=1e-9, atol=1e-12), (l, m)
    ds = registry.get("datasets", "SphericalHarmonicDataset")(max_l=5, random_seed=1616)
    assert [ds.random_set for _ in range(8)] == [tuple(t) for t in g["random_sets"]]
    ds = SphericalHarmonicDataset(max_l=5, random_seed=1616)
    x, emb = ds[0]
    l, m = g["random_sets"][0]
    assert x.shape == (1, 32, 32, 32) and x.dtype == torch.float32 and len(ds) == 1000
    assert torch.equal(emb, calculate_sha512_embedding({"l": l, "m": m}, l=256))


def test_bench_reference_arm_runs_on_cpu_and_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the CUDA path) needs no GPU: one bounded config-D eval of the oracle
    port on the host cores, one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RHO_BENCH_CPU_BUDGET_S="2")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], env=env, cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "UNet denoise evals/s" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_deep_galaxy_config_builds_the_reference_checkpoint_layout():
    """examples/config_deep_galaxy.json (the reference's file up to the device strings) through scripts/inference.py::build_pipeline on
    CPU: DeepGalaxyDataset's class-level parameter space equals the reference's (deep_galaxy.py:41-47), and the backbone -- UNetv2 +
    MultiEmbeddings over that space -- has exactly the reference's state-dict keys and shapes, so its model.pth loads strictly."""
    import importlib.util
    from conftest import load_golden
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = load_golden("deep_galaxy.pt")
    from rho_diffusion.config import ExperimentConfig
    from rho_diffusion.data import DeepGalaxyDataset
    space = DeepGalaxyDataset.parameter_space
    assert list(space.keys()) == list(g["parameter_space"].keys())
    for k, vals in g["parameter_space"].items():
        assert [float(v) for v in space[k]] == vals, k
    with pytest.raises((ImportError, OSError)):      # the reader is real now: no h5py in this image / no such file with it
        DeepGalaxyDataset(path="missing.h5")
    spec = importlib.util.spec_from_file_location("rho_inference_script", os.path.join(root, "scripts", "inference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    config = ExperimentConfig.from_json(os.path.join(root, "examples", "config_deep_galaxy.json"))
    assert {k: [float(v) for v in vals] for k, vals in config.inference.parameter_space.items()} == \
        {k: [float(v) for v in vals] for k, vals in g["inference_parameter_space"].items()}
    pipe, cond_kwargs = mod.build_pipeline(config, torch.device("cpu"), n_samples=4, steps=10)
    assert cond_kwargs is not None and cond_kwargs["embedding_dim"] == 128
    got = {k: tuple(v.shape) for k, v in pipe.backbone.state_dict().items()}
    assert got == g["manifest"]
    assert sum(int(torch.tensor(s).prod()) for s in got.values()) == 15863457


def test_plot_image_grid_matches_torchvision_and_writes_a_png(tmp_path):
    """utils.py:84-117 of the reference: clamp to [-1, 1], flatten [rows, cols, C, H, W], make_grid(nrow = cols, padding 2, white,
    normalize) -> PNG.  The dependency-free restatement (hosts without torchvision) equals torchvision's grid and file pixel for pixel."""
    import zlib
    from rho_diffusion import utils
    g = torch.Generator().manual_seed(3)
    imgs = torch.randn(2, 3, 1, 9, 7, generator=g) * 0.8            # [rows, cols, C, H, W], single channel, values beyond [-1, 1]
    flat = imgs.clamp(-1, 1).reshape(6, 1, 9, 7)
    mine = utils._make_grid(flat, nrow=3)
    assert mine.shape == (3, 2 * 11 + 2, 3 * 9 + 2)
    tv = pytest.importorskip("torchvision.utils")
    ref = tv.make_grid(flat, nrow=3, padding=2, pad_value=1, normalize=True)
    assert torch.allclose(mine, ref, atol=1e-6)
    assert torch.allclose(utils.plot_image_grid(imgs), ref, atol=1e-6)
    assert torch.allclose(utils.plot_image_grid(imgs, transpose=True),
                          tv.make_grid(imgs.clamp(-1, 1).swapaxes(0, 1).reshape(6, 1, 9, 7), nrow=2, padding=2, pad_value=1, normalize=True), atol=1e-6)
    # [N, C, H, W] (what generate() returns) is one row
    assert utils.plot_image_grid(flat).shape == (3, 11 + 2, 6 * 9 + 2)
    # files: torchvision's writer and the zlib writer decode to the same pixels
    f1, f2 = tmp_path / "tv.png", tmp_path / "own.png"
    assert utils.plot_image_grid(imgs, filename=f1) is None and f1.stat().st_size > 0
    utils._write_png(mine, f2)
    from PIL import Image
    import numpy as np
    a, b = np.asarray(Image.open(f1).convert("RGB")), np.asarray(Image.open(f2).convert("RGB"))
    assert a.shape == b.shape == (24, 29, 3) and int(np.abs(a.astype(int) - b.astype(int)).max()) <= 1   # (x*255+.5 on 1e-7-different floats)
    with pytest.raises(ValueError):
        utils.plot_image_grid(torch.zeros(2, 3, 4))


def test_inference_script_plots_an_existing_cache_without_sampling(tmp_path):
    """scripts/inference.py:94-107 of the reference: with a cache file present and no ``-f`` the script only plots it (no model, no
    device).  Without h5py the cache is ``<cache_file>.npz`` with the same ``data`` key."""
    import subprocess
    import sys
    import numpy as np
    cache = tmp_path / "cache.h5"
    data = torch.randn(5, 1, 16, 16, generator=torch.Generator().manual_seed(1)).numpy()
    try:
        import h5py
        with h5py.File(cache, "w") as f:
            f["data"] = data
    except ImportError:
        np.savez(str(cache) + ".npz", data=data)
    png = tmp_path / "grid.png"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "inference.py"), os.path.join(ROOT, "examples", "config_gaussian_blobs.json"),
                        "--cache", str(cache), "--plot", str(png), "--out", str(tmp_path / "never.pt")],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert "Found a cache copy" in r.stdout and png.read_bytes()[:8] == b"\x89PNG\r\n\x1a\n" and not (tmp_path / "never.pt").exists()
