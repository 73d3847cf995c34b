"""Parity of the two remaining registered backbones -- the v1 ``UNet`` (models/unet.py) and the diffusion
``VisionTransformer`` (models/vit.py) -- against outputs and parameter gradients of the UNMODIFIED reference modules
(tests/golden/backbones.pt, oracle/make_golden.py) and, where the reference itself raises (3-D v1 blocks), against the pinned CPU
oracle.  Plus the kernels only these backbones use (rho_activation, rho_convt_weight, rho_patchify, rho_layernorm_*, ...) one by
one against CPU PyTorch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
from oracle.cases import BACKBONE_CASES  # noqa: E402
from oracle.synth import synth_input, synth_state_dict  # noqa: E402

DEV = "cuda"
FP32_TOL, BF16_TOL = 1e-5, 2e-2      # north_star: 1e-5 relative in fp32 (TF32 off), 2e-2 in bf16


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm())


def build(case, dtype):
    from rho_diffusion.b200 import set_compute_dtype
    from rho_diffusion.registry import registry
    spec = BACKBONE_CASES[case]
    g = load_golden("backbones.pt")[case]
    model = registry.get("models", spec["model"])(**spec["kwargs"])
    sd = synth_state_dict(g["manifest"], seed=777)
    model.load_state_dict(sd, strict=True)                        # the reference's checkpoint layout
    model = set_compute_dtype(model.to(DEV), dtype)
    x = synth_input(case + ".x", spec["x_shape"])
    target = synth_input(case + ".target", spec["x_shape"])
    t = torch.tensor(spec["t"], dtype=torch.long)
    if "out" not in g:                                            # the reference raises on this case: the pinned oracle is the checker
        fn = O.unet1_forward if spec["model"] == "UNet" else O.vit_forward
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        y = fn(params, spec["kwargs"], x, t)
        loss = F.mse_loss(y, target)
        loss.backward()
        g = dict(g, out=y.detach(), loss=float(loss), grad_norms={k: float(p.grad.norm()) for k, p in params.items() if p.grad is not None},
                 grads={k: p.grad for k, p in params.items() if p.grad is not None and p.numel() <= 4096})
    return model, g, x, t, target


@pytest.mark.parametrize("case", list(BACKBONE_CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_backbone_forward_and_gradients_match_reference(case, dtype):
    from rho_diffusion.b200 import ops
    model, g, x, t, target = build(case, dtype)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    model.eval()
    with torch.inference_mode():
        out = model(x.to(DEV), t.to(DEV))
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    e = rel(out, g["out"])
    print(f"{case} {dtype}: forward rel L2 {e:.2e}")
    assert e < tol
    model.train()
    out = model(x.to(DEV), t.to(DEV))
    loss = ops.MseLossFn.apply(out, target.to(DEV))
    loss.backward()
    assert abs(float(loss) - g["loss"]) < (1e-5 if dtype == torch.float32 else 3e-2) * g["loss"]
    named = dict(model.named_parameters())
    assert set(g["grad_norms"]) == {k for k, p in named.items() if p.grad is not None}
    # per tensor: 2e-4 in fp32; in bf16 a tensor's gradient may move by up to 15 % (ReLU gates flip under rounding, small
    # tensors are sums of cancelling terms) while the norm of the whole gradient stays within 2e-2
    gtol = 2e-4 if dtype == torch.float32 else 0.15
    total = sum(v ** 2 for v in g["grad_norms"].values()) ** 0.5
    got_total = sum(float(named[k].grad.float().norm()) ** 2 for k in g["grad_norms"]) ** 0.5
    worst, worst_key = 0.0, None
    for k, ref_norm in g["grad_norms"].items():
        got = float(named[k].grad.float().norm())
        # small gradients are judged against the global gradient norm
        err = abs(got - ref_norm) / max(ref_norm, 1e-3 * total)
        if k in g["grads"]:
            err = max(err, float((named[k].grad.float().cpu() - g["grads"][k]).norm() / max(ref_norm, 1e-3 * total)))
        if err > worst:
            worst, worst_key = err, k
    print(f"{case} {dtype}: worst relative gradient error {worst:.2e} ({worst_key}); gradient norm {got_total:.5f} vs {total:.5f}")
    assert worst < gtol, worst_key
    assert abs(got_total - total) < (1e-4 if dtype == torch.float32 else 2e-2) * total


def test_v1_unet_trains_through_the_fused_optimizer():
    """optimize() (flat buffers, gradient sinks, fused AdamW) on the v1 UNet: the loss of a fixed batch drops."""
    from rho_diffusion.b200 import ops, optimize
    model, g, x, t, target = build("unet1_2d", torch.float32)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3)
    model, fused = optimize(model, optimizer=opt, dtype=torch.bfloat16)
    model.train()
    losses = []
    for _ in range(8):
        fused.zero_grad()
        loss = ops.MseLossFn.apply(model(x.to(DEV), t.to(DEV)), target.to(DEV))
        loss.backward()
        fused.step()
        losses.append(float(loss))
    assert losses[-1] < 0.8 * losses[0], losses


def test_vit_trains_through_the_fused_optimizer_and_rejects_training_dropout():
    from rho_diffusion.b200 import ops, optimize
    from rho_diffusion.registry import registry
    model, g, x, t, target = build("vit_2d", torch.float32)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    model, fused = optimize(model, optimizer=opt, dtype=torch.bfloat16)
    model.train()
    losses = []
    for _ in range(8):
        fused.zero_grad()
        loss = ops.MseLossFn.apply(model(x.to(DEV), t.to(DEV)), target.to(DEV))
        loss.backward()
        fused.step()
        losses.append(float(loss))
    assert losses[-1] < 0.9 * losses[0], losses
    kw = dict(BACKBONE_CASES["vit_2d"]["kwargs"], dropout=0.2)
    m2 = registry.get("models", "VisionTransformer")(**kw).to(DEV)
    m2.eval()
    with torch.inference_mode():
        assert torch.isfinite(m2(x.to(DEV), t.to(DEV))).all()     # eval: dropout is the identity
    m2.train()
    with pytest.raises(NotImplementedError):
        m2(x.to(DEV), t.to(DEV))


# ---- the kernels behind them, one by one --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["SiLU", "ReLU", "Tanh", "Sigmoid", "GELU", "SymmetricLog"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_activation_fwd_bwd(name, dtype):
    from rho_diffusion.b200 import ops
    x = (synth_input("act.x", (3, 7, 11, 13)) * 2).to(dtype)      # 3003 elements: exercises the n % 4 tail
    dy = synth_input("act.dy", (3, 7, 11, 13)).to(dtype)
    xr = x.float().clone().requires_grad_(True)
    yr = O._ACTS[name](xr)
    yr.backward(dy.float())
    xg = x.to(DEV).requires_grad_(True)
    y = ops.activation(xg, ops.act_code(name))
    y.backward(dy.to(DEV))
    tol = 2e-6 if dtype == torch.float32 else 8e-3
    assert rel(y, yr) < tol and rel(xg.grad, xr.grad) < tol


def test_convt_weight_is_the_conv_form_of_a_stride1_transposed_conv():
    from rho_diffusion.b200 import ops
    w = synth_input("convt.w", (5, 7, 3, 3, 3)).to(DEV).requires_grad_(True)
    weff = ops.ConvTWeightFn.apply(w)
    x = synth_input("convt.x", (2, 5, 4, 6, 6))
    ref = F.conv_transpose3d(x, w.detach().cpu(), None, stride=1, padding=1)
    got = F.conv3d(x, weff.detach().cpu(), None, padding=1)
    assert torch.allclose(got, ref, atol=1e-5)
    gr = synth_input("convt.g", tuple(weff.shape)).to(DEV)
    weff.backward(gr)
    assert torch.equal(w.grad.cpu(), gr.cpu().transpose(0, 1).flip(2, 3, 4))


@pytest.mark.parametrize("shape,patch", [((2, 3, 16, 16), (4, 4)), ((2, 1, 4, 8, 8), (2, 2, 2)), ((3, 2, 128), (8,)), ((2, 2, 9, 10), (4, 4))])
def test_patchify_is_the_strided_conv_permutation(shape, patch):
    from rho_diffusion.b200 import ops
    nd = len(patch)
    x = synth_input("patch.x", shape)
    E = 5
    w = synth_input("patch.w", (E, shape[1], *patch))
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    ref = conv(x, w, None, stride=patch).flatten(2).transpose(1, 2)                   # [N, T, E]
    xg = x.to(DEV).requires_grad_(True)
    tok = ops.patchify(xg, patch, torch.float32)
    got = tok.cpu().detach() @ w.reshape(E, -1).t()
    assert torch.allclose(got, ref, atol=1e-5)
    # inverse (+ bias) == kernel = stride ConvTranspose of the identity-per-feature weights
    bias = synth_input("patch.b", (shape[1],)).to(DEV).requires_grad_(True)
    grid = [s // p for s, p in zip(shape[2:], patch)]
    field = (shape[0], shape[1], *[g * p for g, p in zip(grid, patch)])
    back = ops.unpatchify(tok, patch, field, torch.float32, bias)
    crop = x[(slice(None), slice(None), *[slice(0, f) for f in field[2:]])]
    assert torch.allclose(back.cpu().detach(), crop + bias.detach().cpu().reshape(1, -1, *([1] * nd)), atol=1e-6)
    gr = synth_input("patch.g", field)
    back.backward(gr.to(DEV))
    gx = torch.zeros(shape)
    gx[(slice(None), slice(None), *[slice(0, f) for f in field[2:]])] = gr
    assert torch.allclose(xg.grad.cpu(), gx, atol=1e-6)
    assert torch.allclose(bias.grad.cpu(), gr.sum(dim=[0, *range(2, 2 + nd)]), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("N,T,C,pre", [(2, 16, 64, True), (3, 50, 96, False), (4, 256, 256, True)])
def test_layernorm_fwd_bwd(N, T, C, pre, dtype):
    from rho_diffusion.b200 import ops
    x = synth_input("ln.x", (N, T, C)).to(dtype)
    dy = synth_input("ln.dy", (N, T, C)).to(dtype)
    gamma = 1 + 0.1 * synth_input("ln.g", (C,))
    beta = 0.1 * synth_input("ln.b", (C,))
    pa = synth_input("ln.pa", (N, C)) if pre else None
    xr, gr, br = x.float().clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    par = pa.clone().requires_grad_(True) if pre else None
    yr = F.layer_norm(xr + par[:, None, :] if pre else xr, (C,), gr, br, 1e-5)
    yr.backward(dy.float())
    xg = x.to(DEV).requires_grad_(True)
    gg, bg = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    pag = pa.to(DEV).requires_grad_(True) if pre else None
    y = ops.layernorm(xg, gg, bg, 1e-5, pre_add=pag)
    y.backward(dy.to(DEV))
    tol = 3e-6 if dtype == torch.float32 else 1e-2
    assert rel(y, yr) < tol and rel(xg.grad, xr.grad) < tol
    assert rel(gg.grad, gr.grad) < tol and rel(bg.grad, br.grad) < tol
    if pre:
        assert rel(pag.grad, par.grad) < (tol if dtype == torch.float32 else 3e-2)


def test_add_bcast_and_bias_expand():
    from rho_diffusion.b200 import ops
    x = synth_input("ab.x", (3, 10, 8)).to(DEV).requires_grad_(True)
    a = synth_input("ab.a", (10, 8)).to(DEV).requires_grad_(True)
    y = ops.add_bcast(x, a)
    gr = synth_input("ab.g", (3, 10, 8)).to(DEV)
    y.backward(gr)
    assert torch.allclose(y.detach(), x.detach() + a.detach()) and torch.equal(x.grad, gr)
    assert torch.allclose(a.grad, gr.sum(0), atol=1e-6)
    b = synth_input("be.b", (6,)).to(DEV).requires_grad_(True)
    e = ops.BiasExpandFn.apply(b, 4)
    assert torch.equal(e.detach(), b.detach().repeat_interleave(4))
    ge = synth_input("be.g", (24,)).to(DEV)
    e.backward(ge)
    assert torch.allclose(b.grad, ge.reshape(6, 4).sum(1), atol=1e-6)


@pytest.mark.parametrize("case", ["unet1_2d", "vit_2d"])
def test_ddpm_pipeline_trains_and_samples_with_the_other_backbones(case):
    """DDPM(<registry name>) unconditional (ddpm.py:231-288, 132-229): the training-step loss equals the oracle's on shared (t, noise),
    and a T = 50 ancestral trajectory with shared noise -- eager and through the replayed CUDA graph -- equals the oracle's, with the
    final field's mean / variance equal to 1e-4."""
    from rho_diffusion.b200 import set_compute_dtype
    from rho_diffusion.diffusion import DDPM, LinearSchedule
    spec = BACKBONE_CASES[case]
    g = load_golden("backbones.pt")[case]
    T = 50          # LinearSchedule scales beta by 1000 / T (schedule.py:141-168): T <= 20 gives beta > 1
    ddpm = DDPM(spec["model"], dict(spec["kwargs"]), LinearSchedule(T, 1e-3, 0.02), "MSELoss", timesteps=T, optimizer="AdamW", opt_kwargs={"lr": 1e-4})
    sd = synth_state_dict(g["manifest"], seed=777)
    ddpm.backbone.load_state_dict(sd, strict=True)
    ddpm = set_compute_dtype(ddpm.to(DEV), torch.float32)
    fn = O.unet1_forward if spec["model"] == "UNet" else O.vit_forward
    den = lambda x, tt, c: fn(sd, spec["kwargs"], x, tt)  # noqa: E731
    sched = O.linear_schedule(T, 1e-3, 0.02)
    x0 = synth_input(case + ".x0", spec["x_shape"], kind="rand") * 2 - 1
    noise = synth_input(case + ".noise", spec["x_shape"])
    t = torch.tensor(spec["t"]) % T
    ddpm.train()
    loss = ddpm.training_step([x0.to(DEV), None], t=t, noise=noise.to(DEV))
    loss_ref = O.mse_loss(den(O.q_sample(sched, x0, t, noise), t, None), noise)
    assert float(loss) == pytest.approx(float(loss_ref), rel=2e-5)
    ddpm.eval()
    x_init = synth_input(case + ".x_init", spec["x_shape"])
    zs = {s: synth_input(f"{case}.z{s}", spec["x_shape"]) for s in range(T)}
    with torch.no_grad():
        ref = O.reverse_process(sched, den, x_init, zs)
    for graph in (False, True):
        res = ddpm.reverse_process(torch.zeros(spec["x_shape"], device=DEV), conditions=None, x_init=x_init.to(DEV), noises=zs, use_cuda_graph=graph)
        got = res["denoised"].cpu()
        assert float((got - ref).abs().max()) < 1e-3, ("graph" if graph else "eager")     # 50 chained fp32 evals + clamps
        a, b = got.double().flatten(), ref.double().flatten()
        assert abs(float(a.mean() - b.mean())) < 1e-4 and abs(float(a.var() - b.var())) < 1e-4
