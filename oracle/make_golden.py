"""TEST INFRASTRUCTURE -- generates tests/golden/*.pt by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Every fixture is small (inputs are re-derived from seeds via oracle/synth.py; only outputs and a
key->shape manifest are stored).  The committed fixtures pin the CPU oracle (oracle/torch_oracle.py)
and are also compared directly against the CUDA path in the ``-m gpu`` tests.
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.cases import TIMESTEPS, UNET_CASES, cond_values, loss_curve_inputs  # noqa: E402
from oracle.synth import synth_input, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_ref_unet(ref, case):
    spec = UNET_CASES[case]
    kwargs = dict(spec["kwargs"])
    if spec["cond"] == "multi":
        space = OrderedDict(spec["parameter_space"])
        kwargs["cond_fn"] = ref.MultiEmbeddings(parameter_space=space, embedding_dim=4 * kwargs["model_channels"])
    model = ref.UNet(**kwargs).eval()
    manifest = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth_state_dict(manifest, seed=777), strict=True)
    return model, manifest


def case_inputs(case):
    spec = UNET_CASES[case]
    x = synth_input(case + ".x", spec["x_shape"])
    t = torch.tensor(TIMESTEPS[case], dtype=torch.long)
    y = None
    if spec["cond"] == "same_shape":
        y = synth_input(case + ".y", (spec["x_shape"][0], 4 * spec["kwargs"]["model_channels"]), kind="rand")
    elif spec["cond"] == "multi":
        y = torch.tensor(cond_values(case), dtype=torch.float32)
    return x, t, y


def gen_unet(ref, only=None):
    for case in UNET_CASES:
        if only is not None and case not in only:
            continue
        torch.manual_seed(0)
        model, manifest = build_ref_unet(ref, case)
        x, t, y = case_inputs(case)
        taps = {}
        hooks = []

        def mk(name):
            def hook(_m, _i, o):
                taps[name] = o.detach().clone()
            return hook

        for i, blk in enumerate(model.input_blocks):
            hooks.append(blk.register_forward_hook(mk(f"input_blocks.{i}")))
        hooks.append(model.middle_block.register_forward_hook(mk("middle_block")))
        for i, blk in enumerate(model.output_blocks):
            hooks.append(blk.register_forward_hook(mk(f"output_blocks.{i}")))
        hooks.append(model.time_embed.register_forward_hook(mk("time_embed")))
        with torch.inference_mode():
            out = model(x, t, y) if y is not None else model(x, t)
        for h in hooks:
            h.remove()
        tap_stats = {k: torch.stack([v.double().mean(), v.double().std(), v.double().abs().max()]).float()
                     for k, v in taps.items()}
        small_taps = {k: v.clone() for k, v in taps.items() if v.numel() <= 9000}
        torch.save({"case": case, "manifest": manifest, "out": out.clone(), "tap_stats": tap_stats,
                    "taps": small_taps}, os.path.join(OUT, f"unet_{case}.pt"))
        print(case, "params", sum(int(torch.tensor(s).prod()) for s in manifest.values()),
              "out", tuple(out.shape), float(out.abs().mean()))


def gen_schedule(ref):
    S = ref.schedule
    d = {}
    for name, sch in {
        "linear_100_kat": S.LinearSchedule(100, 1e-4, 0.02),          # tests/pipeline/test_schedule.py:30
        "linear_1000": S.LinearSchedule(1000, 1e-3, 0.02),           # examples/config_spherical_harmonics.json
        "linear_500_default": S.LinearSchedule(500),                 # examples/config_deep_galaxy.json
        "cosine_50": S.CosineBetaSchedule(50),
        "cosine_1000": S.CosineBetaSchedule(1000),
    }.items():
        d[name] = {k: sch[k].clone() for k in ("beta_t", "alpha_t", "alpha_bar_t", "sigma_t")}
        d[name]["offset_alpha_bar_t"] = sch.offset_alpha_bar_t.clone()
    torch.save(d, os.path.join(OUT, "schedule.pt"))


def gen_embeddings(ref):
    d = {}
    for dim in (32, 64, 128):
        t = torch.tensor([0, 1, 2, 17, 499, 999], dtype=torch.long)
        d[f"timestep_{dim}"] = {"t": t, "pe": ref.common.sinosoidal_position_embedding(t, dim)}
    d["sha512_l2_m-1_256"] = ref.utils.calculate_sha512_embedding({"l": 2, "m": -1}, 256)
    d["sha512_default_128"] = ref.utils.calculate_sha512_embedding({"a": 1})
    space = {"l": [-2, -1, 1, 2], "m": [-1, 1, 3]}
    d["param_space"] = space
    d["param_space_rows_7"] = ref.parameter_space.sample_from_discrete_parameter_space(space, 7, random=False)
    d["param_space_embeddings"] = ref.utils.parameter_space_to_embeddings({"m": [1, 2], "n": [3, 4, 5]})
    # MultiEmbeddings forward
    pspace = OrderedDict({"s": [0.25, 0.5, 1.0], "t": [400.0, 405.0]})
    me = ref.MultiEmbeddings(parameter_space=pspace, embedding_dim=16)
    manifest = {k: tuple(v.shape) for k, v in me.state_dict().items()}
    me.load_state_dict(synth_state_dict(manifest, seed=5))
    y = torch.tensor([[0.5, 405.0], [1.0, 400.0], [0.25, 400.0]])
    d["multi_emb"] = {"space": dict(pspace), "manifest": manifest, "y": y, "out": me(y).detach().clone()}
    torch.save(d, os.path.join(OUT, "embeddings.pt"))


def gen_ddpm(ref):
    """DDPM.forward_process / training_step / reverse_process on the film3d-family UNet, with the RNG
    draws recorded so the fixture does not depend on torch's generator implementation."""
    case = "film3d"
    spec = UNET_CASES[case]
    kwargs = dict(spec["kwargs"])
    T = 50
    sched = ref.schedule.LinearSchedule(T, 1e-3, 0.02)
    ddpm = ref.DDPM(ref.UNet, kwargs, sched, "MSELoss", timesteps=T, optimizer="AdamW", opt_kwargs={"lr": 1e-4})
    manifest = {k: tuple(v.shape) for k, v in ddpm.backbone.state_dict().items()}
    ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
    ddpm.eval()
    x0 = synth_input("ddpm.x0", spec["x_shape"], kind="rand")
    y = synth_input(case + ".y", (spec["x_shape"][0], 4 * kwargs["model_channels"]), kind="rand")
    d = {"T": T, "manifest": manifest}

    # ---- record every randn_like / randint the pipeline makes
    rec = {"randn": [], "randint": []}
    real_randn_like, real_randint = torch.randn_like, torch.randint

    def randn_like(x, *a, **k):
        r = real_randn_like(x, *a, **k)
        rec["randn"].append(r.clone())
        return r

    def randint(*a, **k):
        r = real_randint(*a, **k)
        rec["randint"].append(r.clone())
        return r

    torch.randn_like, torch.randint = randn_like, randint
    try:
        torch.manual_seed(21516)
        # forward_process with explicit t
        t = torch.tensor([3, 10])
        x_t, noise = ddpm.forward_process(x0, t)
        d["q_sample"] = {"t": t, "noise": noise.clone(), "x_t": x_t.clone()}
        # training_step (list batch => conditional)
        rec["randn"].clear(); rec["randint"].clear()
        loss = ddpm.training_step([x0, y])
        d["train"] = {"t_idx": rec["randint"][0].clone(), "noise": rec["randn"][0].clone(), "loss": loss.detach().clone()}
        # gradients of a few parameters (for the backward parity tests)
        ddpm.zero_grad()
        loss.backward()
        grads = {}
        for name, p in ddpm.backbone.named_parameters():
            if p.grad is not None:
                g = p.grad.detach()
                grads[name] = torch.stack([g.double().norm(), g.double().sum()]).float()
        d["train"]["grad_stats"] = grads
        names = [n for n, _ in ddpm.backbone.named_parameters()]
        keep = ["out.2.weight", "out.2.bias", "input_blocks.0.0.weight", "time_embed.0.weight",
                "middle_block.1.qkv.weight", "middle_block.0.emb_layers.1.weight", "out.0.weight",
                "middle_block.1.norm.bias", "middle_block.1.proj_out.weight",
                "input_blocks.1.0.in_layers.2.weight", "input_blocks.1.0.out_layers.0.weight"]
        keep += [n for n in names if n.endswith("op.weight")][:1]
        keep += [n for n in names if n.endswith("conv.weight")][:1]
        keep += [n for n in names if n.endswith("skip_connection.weight")][:1]
        d["train"]["grads"] = {k: dict(ddpm.backbone.named_parameters())[k].grad.detach().flatten()[:4096].clone() for k in keep}
        # one AdamW step through the reference's optimizer factory (world = 4 => lr * 2)
        opt = ddpm.configure_optimizers(4)["optimizer"]
        d["opt"] = {"lr": opt.param_groups[0]["lr"], "weight_decay": opt.param_groups[0]["weight_decay"],
                    "betas": tuple(opt.param_groups[0]["betas"]), "eps": opt.param_groups[0]["eps"]}
        opt.step()
        d["opt"]["after_step"] = {k: dict(ddpm.backbone.named_parameters())[k].detach().flatten()[:4096].clone() for k in keep}
        # EMA (ema.py) two updates
        ema = ref.ExponentialMovingAverage(ddpm.backbone, decay=0.9999)
        ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=778))
        ema.update()
        ema.update()
        d["ema"] = {"frac": ema.current_ema_frac,
                    "shadow": {k: dict(ema.ema_model.named_parameters())[k].detach().flatten()[:4096].clone() for k in keep}}
        # reverse process (weights back to seed 777)
        ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
        rec["randn"].clear()
        torch.manual_seed(4242)
        res = ddpm.reverse_process(torch.zeros(spec["x_shape"]), conditions=y, t_checkpoints=None)
        draws = rec["randn"]
        d["reverse"] = {"x_init": draws[0].clone(), "z": [z.clone() for z in draws[1:]],
                        "denoised": res["denoised"].clone()}
    finally:
        torch.randn_like, torch.randint = real_randn_like, real_randint
    torch.save(d, os.path.join(OUT, "ddpm_film3d.pt"))
    print("ddpm loss", float(d["train"]["loss"]), "reverse draws", len(d["reverse"]["z"]),
          "denoised mean", float(d["reverse"]["denoised"].mean()))


def gen_gaussian(ref):
    """GaussianDiffusionPipeline (x0-prediction, cosine tables, dynamic thresholding, DDIM eta = 0) on the film3d UNet:
    tables, q_sample, p_mean_variance / p_sample / ddim_sample on fixed inputs, the 20-step generate loop and the
    double-noising training loss, with every RNG draw recorded."""
    case = "film3d"
    spec = UNET_CASES[case]
    kwargs = dict(spec["kwargs"])
    T = 20
    sched = ref.schedule.LinearSchedule(T, 1e-3, 0.02)   # ignored by the pipeline (it builds its own cosine tables)
    pipe = ref.GaussianDiffusionPipeline(ref.UNet, kwargs, sched, "MSELoss", timesteps=T, optimizer="AdamW", opt_kwargs={"lr": 1e-4})
    manifest = {k: tuple(v.shape) for k, v in pipe.backbone.state_dict().items()}
    pipe.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
    pipe.eval()
    x0 = synth_input("gauss.x0", spec["x_shape"], kind="rand") * 2 - 1
    y = synth_input(case + ".y", (spec["x_shape"][0], 4 * kwargs["model_channels"]), kind="rand")
    d = {"T": T, "manifest": manifest}
    d["tables"] = {k: torch.from_numpy(getattr(pipe, k).copy()) for k in
                   ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                    "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")}
    rec = {"randn": [], "randint": []}
    real_randn_like, real_randint = torch.randn_like, torch.randint

    def randn_like(x, *a, **k):
        r = real_randn_like(x, *a, **k)
        rec["randn"].append(r.clone())
        return r

    def randint(*a, **k):
        r = real_randint(*a, **k)
        rec["randint"].append(r.clone())
        return r

    torch.randn_like, torch.randint = randn_like, randint
    try:
        torch.manual_seed(21516)
        t = torch.tensor([3, 17])
        x_t, noise = pipe.forward_process(x0, t)
        d["q_sample"] = {"t": t, "noise": noise.clone(), "x_t": x_t.clone()}
        # fixed "model": a scaled copy of the input so that the dynamic threshold is active (|x| quantile > 1)
        fake = synth_input("gauss.model_out", spec["x_shape"]) * 1.7
        with torch.no_grad():
            pmv = pipe.p_mean_variance(lambda x, t, **kw: fake, x_t, t, clip_denoised=True)
            d["p_mean_variance"] = {k: v.clone() for k, v in pmv.items()}
            rec["randn"].clear()
            ps = pipe.p_sample(lambda x, t, **kw: fake, x_t, t)
            d["p_sample"] = {"noise": rec["randn"][0].clone(), "sample": ps["sample"].clone()}
            t0 = torch.tensor([0, 5])
            rec["randn"].clear()
            ps0 = pipe.p_sample(lambda x, t, **kw: fake, x_t, t0)
            d["p_sample_t0"] = {"t": t0, "noise": rec["randn"][0].clone(), "sample": ps0["sample"].clone()}
            dd = pipe.ddim_sample(lambda x, t, **kw: fake, x_t, t)
            d["ddim_sample"] = {"sample": dd["sample"].clone(), "pred_xstart": dd["pred_xstart"].clone()}
            d["fake_scale"] = 1.7
        # training step (double noising quirk), conditional
        rec["randn"].clear(); rec["randint"].clear()
        pipe.train()
        loss = pipe.training_step([x0, y])
        d["train"] = {"t_idx": rec["randint"][0].clone(), "noise": rec["randn"][0].clone(), "loss": loss.detach().clone()}
        pipe.zero_grad()
        loss.backward()
        d["train"]["grad_stats"] = {n: torch.stack([p.grad.double().norm(), p.grad.double().sum()]).float()
                                    for n, p in pipe.backbone.named_parameters() if p.grad is not None}
        pipe.eval()
        # generate loop (DDIM eta = 0): only the initial draw matters
        rec["randn"].clear()
        torch.manual_seed(4242)
        res = pipe.reverse_process(torch.zeros(spec["x_shape"]), conditions=y, t_checkpoints=None)
        d["reverse"] = {"x_init": rec["randn"][0].clone(), "denoised": res["denoised"].clone()}
    finally:
        torch.randn_like, torch.randint = real_randn_like, real_randint
    torch.save(d, os.path.join(OUT, "gaussian_film3d.pt"))
    print("gaussian loss", float(d["train"]["loss"]), "denoised mean/std", float(d["reverse"]["denoised"].mean()),
          float(d["reverse"]["denoised"].std()))


LOSS_CURVE_STEPS = 1000


def gen_loss_curve(ref):
    """1000 AdamW steps of the UNMODIFIED reference ``DDPM.training_step`` (lr 1e-4 as in examples/*.json, T = 1000 linear
    schedule, conditional film3d UNet) on 8 cycling synthetic batches with recorded (t, eps): the loss curve the CUDA
    trainer has to reproduce (north-star acceptance: "training loss curve matches the reference over 1k steps")."""
    case = "film3d"
    spec = UNET_CASES[case]
    kwargs = dict(spec["kwargs"])
    T = 1000
    sched = ref.schedule.LinearSchedule(T, 1e-3, 0.02)
    ddpm = ref.DDPM(ref.UNet, kwargs, sched, "MSELoss", timesteps=T, optimizer="AdamW", opt_kwargs={"lr": 1e-4})
    manifest = {k: tuple(v.shape) for k, v in ddpm.backbone.state_dict().items()}
    ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=777))
    ddpm.train()
    opt = ddpm.configure_optimizers(1)["optimizer"]
    shape = spec["x_shape"]
    data = [synth_input(f"curve.x0.{i}", shape, kind="rand") for i in range(8)]
    labels = [synth_input(f"curve.y.{i}", (shape[0], 4 * kwargs["model_channels"]), kind="rand") for i in range(8)]
    cur = {}
    real_randn_like, real_randint = torch.randn_like, torch.randint
    torch.randn_like = lambda x, *a, **k: cur["noise"].clone()
    torch.randint = lambda *a, **k: cur["t"].clone()
    losses = []
    try:
        torch.set_num_threads(max(1, (os.cpu_count() or 2)))
        for step in range(LOSS_CURVE_STEPS):
            i, cur["t"], cur["noise"] = loss_curve_inputs(step, shape, T)
            opt.zero_grad()
            loss = ddpm.training_step([data[i], labels[i]])
            loss.backward()
            opt.step()
            losses.append(float(loss))
            if step % 100 == 0:
                print("step", step, losses[-1], flush=True)
    finally:
        torch.randn_like, torch.randint = real_randn_like, real_randint
    final = {k: v.detach().flatten()[:256].clone() for k, v in ddpm.backbone.state_dict().items()
             if k in ("out.2.weight", "input_blocks.0.0.weight", "time_embed.0.weight", "middle_block.1.qkv.weight")}
    torch.save({"T": T, "manifest": manifest, "losses": torch.tensor(losses, dtype=torch.float64), "final_params": final},
               os.path.join(OUT, "loss_curve_film3d.pt"))
    print("loss curve: first", losses[:3], "last", losses[-3:])


def gen_sph_harm(ref):
    """compute_spherical_harmonic of the UNMODIFIED reference (rho_diffusion/data/synthetic.py:83-124) for a few (l, m) on
    the default 32^3 grid.  scipy.special.sph_harm (removed in SciPy 1.17) and h5py (absent) are shimmed for the import."""
    import importlib
    import types

    import numpy as np
    import scipy.special as sp
    if not hasattr(sp, "sph_harm"):
        sp.sph_harm = lambda m, n, theta, phi: sp.sph_harm_y(n, m, phi, theta)
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.SimpleNamespace(File=object)
    syn = importlib.import_module("rho_diffusion.data.synthetic")
    grid = np.linspace(-2.0, 2.0, 32)
    out = {}
    for l, m in [(0, 0), (1, -1), (2, 1), (3, 3), (4, -2), (5, 4)]:
        _, density, real = syn.compute_spherical_harmonic(grid, grid, grid, m=m, l=l)
        out[(l, m)] = {"density": torch.from_numpy(density.astype(np.float32))[::3, ::3, ::3].clone(),
                       "stats": torch.tensor([density.min(), density.max(), density.mean(), density.std()], dtype=torch.float64)}
    ds = syn.SphericalHarmonicDataset(max_l=5, random_seed=1616)
    first = [ds.random_set for _ in range(8)]
    torch.save({"fields": out, "random_sets": first}, os.path.join(OUT, "sph_harm.pt"))
    print("sph_harm fixture:", first)


def gen_deep_galaxy(ref):
    """Class-level parameter space of the UNMODIFIED reference's DeepGalaxyDataset (rho_diffusion/data/deep_galaxy.py:41-47) and the
    state-dict manifest of the backbone examples/config_deep_galaxy.json builds with it (UNetv2 kwargs :5-26, MultiEmbeddings over
    that space with embedding_dim 128 as scripts/inference.py:122-143 passes it).  h5py (absent) is shimmed for the import."""
    import importlib
    import json
    import types
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.SimpleNamespace(File=object)
    dg = importlib.import_module("rho_diffusion.data.deep_galaxy")
    space = OrderedDict((k, [float(v) for v in vals]) for k, vals in dg.DeepGalaxyDataset.parameter_space.param_dict.items())
    with open("/root/reference/examples/config_deep_galaxy.json") as fh:
        cfg = json.load(fh)
    kwargs = dict(cfg["model"]["kwargs"])
    kwargs.pop("cond_fn")
    kwargs["cond_fn"] = ref.MultiEmbeddings(parameter_space=dg.DeepGalaxyDataset.parameter_space, embedding_dim=128)
    model = ref.UNet(**kwargs)
    manifest = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    torch.save({"parameter_space": space, "manifest": manifest, "inference_parameter_space": cfg["inference"]["parameter_space"]},
               os.path.join(OUT, "deep_galaxy.pt"))
    print("deep_galaxy fixture:", {k: len(v) for k, v in space.items()}, len(manifest), "tensors",
          sum(int(torch.tensor(v).prod()) for v in manifest.values()), "parameters")


def gen_backbones(ref):
    """v1 UNet and VisionTransformer: output of the unmodified reference modules on synthetic weights and inputs,
    and the gradient of mse(out, target) w.r.t. every parameter (norms of all, full tensors of the small ones)."""
    import importlib
    from oracle.cases import BACKBONE_CASES
    importlib.import_module("rho_diffusion.models.unet")
    importlib.import_module("rho_diffusion.models.vit")
    out = {}
    for case, spec in BACKBONE_CASES.items():
        cls = ref.registry.get("models", spec["model"])
        torch.manual_seed(0)
        model = cls(**spec["kwargs"])
        manifest = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(synth_state_dict(manifest, seed=777), strict=True)
        x = synth_input(case + ".x", spec["x_shape"])
        target = synth_input(case + ".target", spec["x_shape"])
        t = torch.tensor(spec["t"], dtype=torch.long)
        entry = {"manifest": manifest}
        try:
            model.train()          # dropout = 0 in every case: train == eval arithmetic
            y = model(x, t)
            loss = torch.nn.functional.mse_loss(y, target)
            loss.backward()
            entry.update(out=y.detach().clone(), loss=float(loss),
                         grad_norms={k: float(p.grad.norm()) for k, p in model.named_parameters() if p.grad is not None},
                         grads={k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None and p.numel() <= 4096})
            print(f"backbone {case}: out {tuple(y.shape)} |y| {float(y.abs().mean()):.4f} loss {float(loss):.5f} "
                  f"{len(entry['grad_norms'])} parameter gradients")
        except RuntimeError as exc:
            entry["reference_error"] = str(exc).splitlines()[0]
            print(f"backbone {case}: the reference raises: {entry['reference_error']}")
        out[case] = entry
    torch.save(out, os.path.join(OUT, "backbones.pt"))


def gen_spectroscopy(ref):
    """SpectroscopyDataset.__getitem__ of the unmodified reference (spectroscopy.py:103-133) on three synthetic records.  The
    reference opens the file with h5py (absent here): a 15-line stand-in `h5py.File` serves the same records from an .npz archive;
    `random.seed` pins the per-item line width."""
    import importlib
    import random
    import tempfile
    import types
    import numpy as np
    from oracle.cases import spectroscopy_records

    class _File:
        def __init__(self, path, mode="r"):
            z = np.load(path)
            self._g = {}
            for k in z.files:
                g, _, name = k.partition("/")
                self._g.setdefault(g, {})[name] = z[k]

        def __len__(self):
            return len(self._g)

        def __getitem__(self, g):
            return self._g[g]

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    sys.modules.setdefault("h5py", types.SimpleNamespace(File=_File))
    mod = importlib.import_module("rho_diffusion.data.spectroscopy")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "records.npz")
        np.savez(path, **spectroscopy_records())
        ds = mod.SpectroscopyDataset(path, grid_size=4096)
        for i in range(3):
            random.seed(1000 + i)
            item = ds[i]
            out[i] = {k: v.clone() for k, v in item.items()}
        out["frequency_grid"] = torch.from_numpy(ds.frequency_grid.copy())
    torch.save(out, os.path.join(OUT, "spectroscopy.pt"))
    print("spectroscopy fixture:", {k: tuple(v.shape) for k, v in out[0].items()}, float(out[0]["spectrum"].sum()))


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == "gaussian":     # regenerate one fixture only
        gen_gaussian(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1].startswith("unet:"):
        gen_unet(ref, only=sys.argv[1][5:].split(","))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sph_harm":
        gen_sph_harm(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "deep_galaxy":
        gen_deep_galaxy(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "spectroscopy":
        gen_spectroscopy(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "backbones":
        gen_backbones(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "loss_curve":
        gen_loss_curve(ref)
        return
    gen_schedule(ref)
    gen_embeddings(ref)
    gen_unet(ref)
    gen_ddpm(ref)
    gen_gaussian(ref)
    gen_sph_harm(ref)
    gen_deep_galaxy(ref)
    gen_backbones(ref)
    gen_spectroscopy(ref)
    gen_loss_curve(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
