"""End-to-end parity of the CUDA path against golden outputs of the UNMODIFIED reference (tests/golden/, made by
oracle/make_golden.py) and against the CPU oracle: full UNet forward (n = 1, 2, 3; all conditioning paths; both
attention orders), DDPM training step (loss + gradients), AdamW / EMA, and the sampling trajectory with shared
noise -- eager and through the replayed CUDA graph."""
from collections import OrderedDict

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
from oracle.cases import TIMESTEPS, UNET_CASES, cond_values  # noqa: E402
from oracle.synth import synth_input, synth_state_dict  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def build(case, dtype=torch.float32):
    from rho_diffusion.b200 import set_compute_dtype
    from rho_diffusion.models.conditioning import MultiEmbeddings
    from rho_diffusion.models.unet_v2 import UNet
    spec = UNET_CASES[case]
    g = load_golden(f"unet_{case}.pt")
    kwargs = dict(spec["kwargs"])
    if spec["cond"] == "multi":
        kwargs["cond_fn"] = MultiEmbeddings(parameter_space=OrderedDict(spec["parameter_space"]),
                                            embedding_dim=4 * kwargs["model_channels"])
    model = UNet(**kwargs)
    model.load_state_dict(synth_state_dict(g["manifest"], seed=777), strict=True)   # reference checkpoint layout
    model = set_compute_dtype(model.to(DEV).eval(), dtype)
    x = synth_input(case + ".x", spec["x_shape"])
    t = torch.tensor(TIMESTEPS[case], dtype=torch.long)
    y = None
    if spec["cond"] == "same_shape":
        y = synth_input(case + ".y", (spec["x_shape"][0], 4 * kwargs["model_channels"]), kind="rand")
    elif spec["cond"] == "multi":
        y = torch.tensor(cond_values(case), dtype=torch.float32)
    return model, g, x, t, y


def nrm_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm()), float((got - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("case", list(UNET_CASES))
def test_unet_forward_fp32_matches_reference(case):
    model, g, x, t, y = build(case)
    with torch.inference_mode():
        out = model(x.to(DEV), t.to(DEV), y.to(DEV)) if y is not None else model(x.to(DEV), t.to(DEV))
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    l2, linf = nrm_err(out, g["out"])
    print(f"{case}: fp32 rel L2 {l2:.2e} max {linf:.2e}")
    assert l2 < 1e-5 and linf < 2e-5


@pytest.mark.parametrize("case", list(UNET_CASES))
def test_unet_forward_bf16_within_tolerance(case):
    model, g, x, t, y = build(case, torch.bfloat16)
    with torch.inference_mode():
        out = model(x.to(DEV), t.to(DEV), y.to(DEV)) if y is not None else model(x.to(DEV), t.to(DEV))
    l2, linf = nrm_err(out, g["out"])
    print(f"{case}: bf16 rel L2 {l2:.2e} max {linf:.2e}")
    assert l2 < 2e-2


def test_per_block_taps_fp32():
    """Per-layer parity: every block output of the 3-D FiLM case against the reference's activations."""
    from rho_diffusion.layers import ExecCtx
    from rho_diffusion.b200 import ops
    model, g, x, t, y = build("film3d")
    ctx = ExecCtx(dtype=torch.float32)
    with torch.inference_mode():
        emb = model._embedding(t.to(DEV), y.to(DEV))
        h = ops.to_channels_last(x.to(DEV), torch.float32)
        hs, seen = [], {}
        for i, m in enumerate(model.input_blocks):
            h = m.forward_cl(h, emb, ctx)
            hs.append(h)
            seen[f"input_blocks.{i}"] = h
        h = model.middle_block.forward_cl(h, emb, ctx)
        seen["middle_block"] = h
        for i, m in enumerate(model.output_blocks):
            h = m.forward_cl((h, hs.pop()), emb, ctx)
            seen[f"output_blocks.{i}"] = h
    checked = 0
    for name, act in seen.items():
        a = ops.cl_to_nc(act, torch.float32).cpu()
        st = g["tap_stats"][name]
        assert abs(float(a.double().mean()) - float(st[0])) < 1e-5 * max(1.0, float(st[2])), name
        assert abs(float(a.double().std()) - float(st[1])) < 1e-5 * max(1.0, float(st[2])), name
        if name in g["taps"]:
            assert nrm_err(a, g["taps"][name])[1] < 2e-5, name
            checked += 1
    print('full-tensor taps checked:', checked)


def make_ddpm(dtype=torch.float32):
    from rho_diffusion.b200 import set_compute_dtype
    from rho_diffusion.diffusion import DDPM, LinearSchedule
    g = load_golden("ddpm_film3d.pt")
    spec = UNET_CASES["film3d"]
    ddpm = DDPM("UNetv2", dict(spec["kwargs"]), LinearSchedule(g["T"], 1e-3, 0.02), "MSELoss", timesteps=g["T"], optimizer="AdamW",
                opt_kwargs={"lr": 1e-4})
    ddpm.backbone.load_state_dict(synth_state_dict(g["manifest"], seed=777))
    ddpm = set_compute_dtype(ddpm.to(DEV), dtype)
    x0 = synth_input("ddpm.x0", spec["x_shape"], kind="rand")
    y = synth_input("film3d.y", (spec["x_shape"][0], 128), kind="rand")
    return ddpm, g, x0, y


def test_ddpm_forward_process_and_training_step_fp32():
    ddpm, g, x0, y = make_ddpm()
    q = g["q_sample"]
    x_t, noise = ddpm.forward_process(x0.to(DEV), q["t"].to(DEV), noise=q["noise"].to(DEV))
    assert (x_t.cpu() - q["x_t"]).abs().max() < 1e-6
    tr = g["train"]
    ddpm.train()
    loss = ddpm.training_step([x0.to(DEV), y.to(DEV)], t=tr["t_idx"], noise=tr["noise"].to(DEV))
    assert float(loss) == pytest.approx(float(tr["loss"]), rel=1e-5)
    loss.backward()
    named = dict(ddpm.backbone.named_parameters())
    errs = []
    for k, stat in tr["grad_stats"].items():
        gnorm = float(named[k].grad.double().norm())
        errs.append((abs(gnorm - float(stat[0])) / max(float(stat[0]), 1e-6), k, float(stat[0])))
    # conv biases that feed a GroupNorm have a mathematically zero gradient (norm ~1e-8 = pure rounding noise): skip them
    floor = 1e-5 * max(e[2] for e in errs)
    errs = [e for e in errs if e[2] > floor]
    errs.sort(reverse=True)
    for e in errs[:8]:
        print("grad-norm rel err %.2e  %s  (ref norm %.3e)" % e)
    assert errs[0][0] < 2e-4
    for k, gref in tr["grads"].items():
        got = named[k].grad.flatten()[:4096].cpu()
        assert float((got - gref).norm() / gref.norm().clamp_min(1e-12)) < 1e-4, k


def test_fused_adamw_matches_reference_step():
    from rho_diffusion.b200 import optimize
    ddpm, g, x0, y = make_ddpm()
    tr = g["train"]
    opt = ddpm.configure_optimizers(4)["optimizer"]          # lr * sqrt(4), as recorded in the golden
    assert opt.param_groups[0]["lr"] == pytest.approx(g["opt"]["lr"])
    ddpm, fused = optimize(ddpm, optimizer=opt, dtype=torch.float32, ema_decay=0.9999)
    assert ddpm.backbone.state_dict().keys() == g["manifest"].keys()
    ddpm.train()
    fused.zero_grad()
    loss = ddpm.training_step([x0.to(DEV), y.to(DEV)], t=tr["t_idx"], noise=tr["noise"].to(DEV))
    loss.backward()
    fused.step()
    named = dict(ddpm.backbone.named_parameters())
    for k, after in g["opt"]["after_step"].items():
        got = named[k].detach().flatten()[:4096].cpu()
        # first AdamW step is ~ lr * sign(g): elements with |g| ~ eps amplify fp32 gradient noise, so bound the max by a
        # fraction of lr and the mean tightly
        assert (got - after).abs().max() < 0.1 * g["opt"]["lr"], k
        assert (got - after).abs().mean() < 1e-7, k
    # the step really went through the flat buffer (params are views into it) and a second forward sees new weights
    assert named["out.2.weight"].data_ptr() >= fused.flat_p.data_ptr()
    loss2 = ddpm.training_step([x0.to(DEV), y.to(DEV)], t=tr["t_idx"], noise=tr["noise"].to(DEV))
    assert float(loss2) != float(loss)
    sd = fused.shadow_state_dict(ddpm.backbone)
    assert sd.keys() == ddpm.backbone.state_dict().keys()


def test_ema_module_matches_reference():
    from rho_diffusion.ema import ExponentialMovingAverage
    ddpm, g, _, _ = make_ddpm()
    # shadow := post-AdamW weights is not reproducible here; check the recurrence from known starts instead
    ema = ExponentialMovingAverage(ddpm.backbone, decay=0.9999)
    ddpm.backbone.load_state_dict(synth_state_dict(g["manifest"], seed=778))
    ema.update()
    ema.update()
    assert ema.current_ema_frac == pytest.approx(g["ema"]["frac"])
    k = "out.2.weight"
    s0 = synth_state_dict({k: g["manifest"][k]}, seed=777)[k]
    p = synth_state_dict({k: g["manifest"][k]}, seed=778)[k]
    want = O.ema_update(O.ema_update(s0, p, 1), p, 2)
    assert (dict(ema.ema_model.named_parameters())[k].cpu() - want).abs().max() < 1e-7


@pytest.mark.parametrize("graph", [False, True])
def test_reverse_process_shared_noise_matches_reference(graph):
    ddpm, g, _, y = make_ddpm()
    ddpm.eval()
    rv = g["reverse"]
    T = g["T"]
    noises = {T - 1 - i: z for i, z in enumerate(rv["z"])}
    res = ddpm.reverse_process(torch.zeros(rv["x_init"].shape, device=DEV), conditions=y.to(DEV), x_init=rv["x_init"].to(DEV),
                               noises=noises, use_cuda_graph=graph, t_checkpoints=[0, 1, 2])
    den = res["denoised"].cpu()
    err = float((den - rv["denoised"]).abs().max())
    print("trajectory max abs err", err, "graph" if graph else "eager")
    assert err < 1e-3   # 50 chained fp32 evals + clamps
    # final density-field moments (north star): mean / variance / third central moment per sample against the reference's trajectory
    ref = rv["denoised"]
    for k in range(den.shape[0]):
        a, b = den[k].double().flatten(), ref[k].double().flatten()
        for mom_a, mom_b in ((a.mean(), b.mean()), (a.var(), b.var()), (((a - a.mean()) ** 3).mean(), ((b - b.mean()) ** 3).mean())):
            assert abs(float(mom_a) - float(mom_b)) < 1e-4
    assert res["buffer"].shape[1] == 3
    assert den.min() >= -1 and den.max() <= 1


def test_reverse_process_on_the_split_bf16_engine_stays_near_the_reference_trajectory():
    """The split-bf16 fp32 engine is as close to exact arithmetic as the CUDA-core engine but rounds differently from both it and the
    reference's CPU kernels; 50 chained evals + clamps amplify that to ~1.4e-3 max-abs on this fixture (CUDA-core engine: < 1e-3),
    with the final density-field moments still equal to 1e-4."""
    from rho_diffusion.b200 import ops
    ddpm, g, _, y = make_ddpm()
    ddpm.eval()
    rv, T = g["reverse"], g["T"]
    noises = {T - 1 - i: z for i, z in enumerate(rv["z"])}
    ops.set_fp32_split(True)
    try:
        res = ddpm.reverse_process(torch.zeros(rv["x_init"].shape, device=DEV), conditions=y.to(DEV), x_init=rv["x_init"].to(DEV),
                                   noises=noises, use_cuda_graph=False)
    finally:
        ops.set_fp32_split(False)
    den, ref = res["denoised"].cpu(), rv["denoised"]
    assert float((den - ref).abs().max()) < 4e-3
    for k in range(den.shape[0]):
        a, b = den[k].double().flatten(), ref[k].double().flatten()
        assert abs(float(a.mean() - b.mean())) < 1e-4 and abs(float(a.var() - b.var())) < 1e-4


def test_reverse_process_graph_in_kernel_noise_is_deterministic_and_bounded():
    ddpm, g, _, y = make_ddpm(torch.bfloat16)
    ddpm.eval()
    shape = g["reverse"]["x_init"].shape
    from rho_diffusion.b200.sampler import AncestralSampler
    alpha, beta, abar = ddpm._tables(torch.device(DEV, 0))
    s = AncestralSampler(ddpm.backbone, shape, torch.device(DEV, 0), alpha, beta, abar, cond_example=y.to(DEV))
    from rho_diffusion.b200 import ops
    x_init = ops.randn(shape, DEV, seed=5)
    a = s.run(x_init, y.to(DEV), seed=11)["denoised"]
    b = s.run(x_init, y.to(DEV), seed=11)["denoised"]
    assert torch.equal(a, b)
    assert torch.isfinite(a).all() and a.min() >= -1 and a.max() <= 1
    e = AncestralSampler(ddpm.backbone, shape, torch.device(DEV, 0), alpha, beta, abar, cond_example=y.to(DEV), use_cuda_graph=False)
    c = e.run(x_init, y.to(DEV), seed=11)["denoised"]
    assert (a - c).abs().max() < 1e-5      # same kernels, same order: graph replay == eager launches


def test_host_staged_stepper_matches_direct_steps():
    """HostStagedStepper (pinned host buffers in and out every step, copies on a second stream through double-buffered staging)
    returns exactly what a direct device-side step on the same inputs returns, for a run of steps with different inputs."""
    from rho_diffusion.b200 import ops
    from rho_diffusion.b200.sampler import AncestralSampler, HostStagedStepper
    ddpm, g, _, y = make_ddpm(torch.bfloat16)
    ddpm.eval()
    shape = g["reverse"]["x_init"].shape
    dev = torch.device(DEV, 0)
    alpha, beta, abar = ddpm._tables(dev)
    with torch.inference_mode():
        s = AncestralSampler(ddpm.backbone, shape, dev, alpha, beta, abar, cond_example=y.to(DEV))
        s._capture(seed=3)
        xs = [ops.randn(shape, DEV, seed=20 + k).cpu().pin_memory() for k in range(5)]
        ys = [(y * (0.5 + 0.1 * k)).pin_memory() for k in range(5)]
        direct = []
        s.tt.fill_(30); s.t_scalar.fill_(30)
        for k in range(5):
            s.x.copy_(xs[k]); s.cond.copy_(ys[k])
            s.graph.replay()
            direct.append(s.x.cpu())
        outs = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(5)]
        stepper = HostStagedStepper(s)
        s.tt.fill_(30); s.t_scalar.fill_(30)
        for k in range(5):
            stepper.step(xs[k], ys[k], outs[k])
        stepper.finish()
        torch.cuda.synchronize()
    for k in range(5):
        assert torch.equal(outs[k], direct[k]), k
    assert stepper.h2d_bytes == xs[0].numel() * 4 + ys[0].numel() * 4 and stepper.d2h_bytes == xs[0].numel() * 4


def test_training_loss_curve_tracks_oracle():
    """Short shared-noise training run (fp32): the loss after each of 5 AdamW steps follows the CPU oracle driven by
    torch autograd + the restated AdamW (the full 1k-step curve is a bench-time check)."""
    from rho_diffusion.b200 import optimize
    ddpm, g, x0, y = make_ddpm()
    spec = UNET_CASES["film3d"]
    uspec = O.UNetSpec(**spec["kwargs"])
    sched = O.linear_schedule(g["T"], 1e-3, 0.02)
    sd = {k: v.clone().requires_grad_(True) for k, v in synth_state_dict(g["manifest"], seed=777).items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    opt = ddpm.configure_optimizers(1)["optimizer"]
    ddpm, fused = optimize(ddpm, optimizer=opt, dtype=torch.float32)
    ddpm.train()
    gen = torch.Generator().manual_seed(3)
    for step in range(1, 6):
        t = torch.randint(0, g["T"], (x0.shape[0],), generator=gen)
        noise = torch.randn(x0.shape, generator=gen)
        # oracle
        x_t = O.q_sample(sched, x0, t, noise)
        loss_ref = O.mse_loss(O.unet_forward(sd, uspec, x_t, t, y), noise)
        grads = torch.autograd.grad(loss_ref, list(sd.values()))
        with torch.no_grad():
            for (k, p), gr in zip(sd.items(), grads):
                p_new, m[k], v2[k] = O.adamw_step(p, gr, m[k], v2[k], step, lr=1e-4)
                p.copy_(p_new)
        # CUDA
        fused.zero_grad()
        loss = ddpm.training_step([x0.to(DEV), y.to(DEV)], t=t, noise=noise.to(DEV))
        loss.backward()
        fused.step()
        print(step, float(loss), float(loss_ref))
        assert float(loss) == pytest.approx(float(loss_ref), rel=5e-4)


def test_optional_unet_structure_trains(ops=None):
    """resblock_updown=True + dropout > 0 in training mode: forward/backward run, every parameter receives a finite
    gradient, and eval mode ignores the dropout (equals the golden of the reference, which was captured in eval mode)."""
    from rho_diffusion.b200 import set_compute_dtype
    from rho_diffusion.models.unet_v2 import UNet
    spec = UNET_CASES["updown3d"]
    g = load_golden("unet_updown3d.pt")
    model = UNet(dropout=0.1, **spec["kwargs"])
    model.load_state_dict(synth_state_dict(g["manifest"], seed=777), strict=True)
    model = set_compute_dtype(model.to(DEV), torch.float32)
    x = synth_input("updown3d.x", spec["x_shape"]).to(DEV)
    t = torch.tensor(TIMESTEPS["updown3d"], dtype=torch.long).to(DEV)
    y = synth_input("updown3d.y", (spec["x_shape"][0], 128), kind="rand").to(DEV)
    model.eval()
    with torch.inference_mode():
        out = model(x, t, y)
    assert nrm_err(out, g["out"])[0] < 1e-5
    model.train()
    out_t = model(x, t, y)
    assert nrm_err(out_t, g["out"])[0] > 1e-3                  # dropout is active in training mode
    out_t.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_graphed_train_step_matches_eager():
    """GraphedTrainStep (whole step in one replayed CUDA graph, device-resident AdamW step counter) reproduces the eager
    zero_grad / training_step / backward / step sequence on the same (t, noise): same losses, same weights after 4 steps."""
    from rho_diffusion.b200 import GraphedTrainStep, optimize

    def run(graphed):
        torch.manual_seed(5)                   # the warm-up steps draw their timesteps from the CPU generator
        ddpm, g, x0, y = make_ddpm()
        opt = ddpm.configure_optimizers(1)["optimizer"]
        ddpm, fused = optimize(ddpm, optimizer=opt, dtype=torch.float32, ema_decay=0.999)
        ddpm.train()
        xd, yd = x0.to(DEV), y.to(DEV)
        gen = torch.Generator().manual_seed(11)
        draws = [(torch.randint(0, g["T"], (x0.shape[0],), generator=gen), torch.randn(x0.shape, generator=gen)) for _ in range(6)]
        losses = []
        def eager_step(t, noise):
            fused.zero_grad()
            loss = ddpm.training_step([xd, yd], t=t, noise=noise)
            loss.backward()
            fused.step()
            return float(loss.detach())

        if graphed:
            before = fused.flat_p.clone()
            step = GraphedTrainStep(ddpm, fused, xd, yd, warmup=2)        # the warm-up steps are undone: weights, moments, EMA, counters
            assert torch.equal(before, fused.flat_p) and fused.steps_taken == 0 and float(fused.flat_m.abs().max()) == 0.0
            assert torch.equal(fused.flat_shadow, fused.flat_p)
        for t, noise in draws:
            losses.append(float(step(xd, yd, t=t.to(DEV), noise=noise.to(DEV))) if graphed else eager_step(t, noise.to(DEV)))
        return losses, fused

    eager, fe = run(False)
    graph, fg = run(True)
    assert fg.steps_taken == 6 and fe.steps_taken == 6
    graph2, _ = run(True)
    assert graph == graph2                                   # replay is bit-reproducible
    print("eager ", eager)
    print("graph ", graph)
    for a, b in zip(eager, graph):                           # same kernels, same order, same inputs
        assert a == pytest.approx(b, rel=1e-5)
    assert torch.allclose(fe.flat_p, fg.flat_p, rtol=0, atol=1e-6)
    assert torch.allclose(fe.flat_shadow, fg.flat_shadow, rtol=0, atol=1e-6)


def test_sampler_graph_is_recaptured_when_the_weights_change():
    """ The captured sampling graph reads packed-weight snapshots: after load_state_dict / training / an EMA swap a
    later reverse_process must not replay stale conv weights next to live Linear / GroupNorm parameters."""
    from oracle.synth import synth_state_dict
    ddpm, g, x0, y = make_ddpm()
    ddpm.eval()
    gen = torch.Generator().manual_seed(3)
    x_init = torch.randn(x0.shape, generator=gen).to(DEV)
    noises = {t: torch.randn(x0.shape, generator=gen) for t in range(g["T"])}
    run = lambda graph: ddpm.reverse_process(torch.zeros(x0.shape, device=DEV), conditions=y.to(DEV), x_init=x_init, noises=noises,  # noqa: E731
                                            num_steps=6, use_cuda_graph=graph)["denoised"].clone()
    a_graph = run(True)
    manifest = {k: tuple(v.shape) for k, v in ddpm.backbone.state_dict().items()}
    ddpm.backbone.load_state_dict(synth_state_dict(manifest, seed=4242))          # new weights, same module objects
    b_graph = run(True)                                                            # same cached sampler -> must re-capture
    ddpm._samplers = {}
    b_eager = run(False)
    assert (a_graph - b_graph).abs().max() > 1e-3
    assert torch.allclose(b_graph, b_eager, rtol=0, atol=2e-5)


def test_sampler_graph_serves_every_seed():
    """ The Philox key of the run lives in device memory: one captured graph is replayed for every reverse_process call
    (no re-capture per call), different calls draw different noise, the same seed reproduces the trajectory."""
    ddpm, g, x0, y = make_ddpm()
    ddpm.eval()
    shape = x0.shape
    r1 = ddpm.reverse_process(torch.zeros(shape, device=DEV), conditions=y.to(DEV), num_steps=5)["denoised"].clone()
    sampler = next(iter(ddpm._samplers.values()))
    graph = sampler.graph
    r2 = ddpm.reverse_process(torch.zeros(shape, device=DEV), conditions=y.to(DEV), num_steps=5)["denoised"].clone()
    assert next(iter(ddpm._samplers.values())).graph is graph                     # not re-captured
    assert (r1 - r2).abs().max() > 1e-3
    x_init = torch.randn(shape, generator=torch.Generator().manual_seed(1)).to(DEV)
    s1 = sampler.run(x_init, y.to(DEV), seed=99, num_steps=5)["denoised"].clone()
    s2 = sampler.run(x_init, y.to(DEV), seed=99, num_steps=5)["denoised"].clone()
    s3 = sampler.run(x_init, y.to(DEV), seed=100, num_steps=5)["denoised"].clone()
    assert torch.equal(s1, s2) and (s1 - s3).abs().max() > 1e-3


def test_dropout_mask_follows_the_device_step_counter():
    """ Under graph replay the host-side call counter is frozen; the mask must still change every optimizer step: the
    device step counter is folded into the Philox offset (same counter value -> same mask, forward and backward agree)."""
    from rho_diffusion.b200 import ops
    x = torch.ones(1 << 14, device=DEV)
    step = torch.zeros(16, dtype=torch.int32, device=DEV)
    old = ops.dropout_step_dev
    try:
        ops.dropout_step_dev = step
        g = torch.cuda.CUDAGraph()
        xin = x.clone().requires_grad_(True)
        ops.dropout(xin, 0.5, seed=7)                           # warm-up outside the capture
        with torch.cuda.graph(g):
            y = ops.dropout(xin, 0.5, seed=7)
            (gx,) = torch.autograd.grad(y, xin, torch.ones_like(y))
        g.replay(); y0, g0 = y.clone(), gx.clone()
        step[0] = 1
        g.replay(); y1 = y.clone()
        step[0] = 0
        g.replay(); y2 = y.clone()
        assert torch.equal(y0, y2) and not torch.equal(y0, y1)
        assert torch.equal(g0, y0)                              # backward of ones through the same mask
        assert 0.4 < float((y1 == 0).float().mean()) < 0.6
    finally:
        ops.dropout_step_dev = old


def test_unmatched_label_raises_like_the_reference():
    from collections import OrderedDict
    from rho_diffusion.models.conditioning import MultiEmbeddings
    me = MultiEmbeddings(parameter_space=OrderedDict([("a", [0.25, 0.5]), ("b", [1, 2, 3])]), embedding_dim=16).to(DEV)
    ok = me(torch.tensor([[0.5, 3.0], [0.25, 1.0]], device=DEV))
    assert torch.isfinite(ok).all()
    me2 = MultiEmbeddings(parameter_space=OrderedDict([("a", [0.25, 0.5]), ("b", [1, 2, 3])]), embedding_dim=16).to(DEV)
    with pytest.raises(ValueError, match="not a point of the parameter space"):
        me2(torch.tensor([[0.5, 3.0], [0.25, 7.0]], device=DEV))
    out = me2(torch.tensor([[0.5, 3.0], [0.25, 7.0]], device=DEV))      # later calls do not sync: the bad row is NaN-poisoned
    assert torch.isfinite(out[0]).all() and torch.isnan(out[1]).all()
