"""Per-kernel parity tests: every C-ABI op vs the CPU oracle (plain PyTorch fp32 on CPU) on seeded inputs.

fp32 path: 1e-5 relative (north star), bf16 path: 2e-2.  Integer work (Philox words, embedding indices) is
bit-exact.  All calls go through the C ABI (ctypes -> librho_b200.so)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import torch_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rho_diffusion.b200 import ops as _ops
    return _ops


DEV = "cuda"


def rel_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


def randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def cl(x):  # (N, C, *sp) cpu -> channels-last cuda
    return x.permute(0, *range(2, x.dim()), 1).contiguous().to(DEV)


def nc(x):  # channels-last cuda -> (N, C, *sp) cpu
    return x.float().cpu().permute(0, x.dim() - 1, *range(1, x.dim() - 1)).contiguous()


TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


# ---------------------------------------------------------------------------------------------------
def test_layout_roundtrip(ops):
    for shape in [(2, 3, 24, 16), (2, 1, 8, 16, 16), (3, 32, 65), (1, 64, 5, 7, 9)]:
        x = randn(*shape, seed=1).to(DEV)
        y = ops.nc_to_cl(x, torch.float32)
        assert torch.equal(nc(y), x.cpu())
        assert torch.equal(ops.cl_to_nc(y, torch.float32).cpu(), x.cpu())
        yb = ops.nc_to_cl(x, torch.bfloat16)
        assert torch.equal(nc(yb), x.cpu().bfloat16().float())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,c2,film,pre", [
    ((2, 64, 8, 8), 0, False, False),
    ((2, 32, 33), 0, True, False),          # 1 channel per group, ragged S
    ((2, 128, 4, 6, 5), 64, True, False),   # concat 128+64 = 192: groups straddle the two sources
    ((1, 96, 7, 9), 0, False, True),        # 3 channels per group, additive embedding
    ((2, 1024, 6), 512, True, False),       # 1536 channels: multi-pass over the channel dimension
])
def test_groupnorm_act_fwd_bwd(ops, dtype, shape, c2, film, pre):
    N, C1 = shape[0], shape[1]
    C = C1 + c2
    x1 = randn(*shape, seed=2) * 1.5 + 0.3
    x2 = randn(N, c2, *shape[2:], seed=3) - 0.5 if c2 else None
    gamma, beta = 1 + 0.1 * randn(C, seed=4), 0.1 * randn(C, seed=5)
    e = 0.3 * randn(N, 2 * C, seed=6)
    pa = 0.5 * randn(N, C, seed=7)
    dy = randn(N, C, *shape[2:], seed=8)

    # ---- oracle (CPU fp32; inputs rounded to the storage dtype first so both sides see the same values)
    q = (lambda t: t.to(dtype).float())
    leaves = [q(x1).requires_grad_(True)]
    if c2:
        leaves.append(q(x2).requires_grad_(True))
    gr, br, er, par = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True), e.clone().requires_grad_(True), pa.clone().requires_grad_(True)
    xin = torch.cat(leaves, 1)
    bshape = (N, C, *([1] * (len(shape) - 2)))
    if pre:
        xin = xin + par.view(bshape)
    h = F.group_norm(xin, 32, gr, br, 1e-5)
    if film:
        h = h * (1 + er[:, :C].reshape(bshape)) + er[:, C:].reshape(bshape)
    ref = F.silu(h)
    ref.backward(q(dy))

    # ---- CUDA
    xs = [cl(t.detach()).to(dtype).requires_grad_(True) for t in leaves]
    gd, bd = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    ed = e.to(DEV).requires_grad_(True) if film else None
    pd = pa.to(DEV).requires_grad_(True) if pre else None
    from rho_diffusion.b200 import _lib
    y = ops.GnActFn.apply(xs[0], xs[1] if c2 else None, gd, bd, ed, pd, _lib.ACT_SILU, dtype, 32, 1e-5)
    y.backward(cl(dy).to(dtype))
    tol = TOL[dtype]
    assert rel_err(nc(y), ref) < tol
    gtol = tol * 4
    assert rel_err(nc(xs[0].grad), leaves[0].grad) < gtol
    if c2:
        assert rel_err(nc(xs[1].grad), leaves[1].grad) < gtol
    assert rel_err(gd.grad, gr.grad) < gtol and rel_err(bd.grad, br.grad) < gtol
    if film:
        assert rel_err(ed.grad, er.grad) < gtol
    if pre:
        assert rel_err(pd.grad, par.grad) < gtol


CONV_CASES = [
    # (dims, N, Cin1, Cin2, Cout, spatial, k, stride, up, residual)
    (2, 2, 3, 0, 32, (24, 16), 3, 1, 1, False),        # reference fixture: 3 input channels (odd), unaligned
    (2, 2, 32, 0, 3, (12, 8), 3, 1, 1, False),         # 3 output channels
    (3, 1, 1, 0, 32, (4, 8, 8), 3, 1, 1, False),       # first conv of the 3-D configs
    (3, 1, 64, 0, 64, (4, 8, 8), 3, 1, 1, True),       # ResBlock out conv + identity residual
    (3, 2, 64, 32, 64, (3, 6, 6), 3, 1, 1, False),     # concat input (up path)
    (3, 1, 64, 32, 32, (3, 6, 6), 1, 1, 1, False),     # 1x1x1 skip conv on a concat input
    (3, 1, 32, 0, 32, (4, 8, 8), 3, (1, 2, 2), 1, False),   # Downsample, 3-D keeps depth
    (2, 2, 32, 0, 32, (8, 10), 3, 2, 1, False),        # Downsample 2-D
    (1, 2, 32, 0, 64, (33,), 3, 2, 1, False),          # Downsample 1-D, odd length
    (3, 1, 32, 0, 32, (3, 4, 5), 3, 1, (1, 2, 2), False),   # Upsample (nearest) fused into the conv, 3-D
    (2, 1, 64, 0, 32, (5, 6), 3, 1, 2, False),         # Upsample 2-D
    (1, 2, 128, 0, 384, (16,), 1, 1, 1, False),        # attention qkv projection (Conv1d k=1)
    (1, 2, 128, 0, 128, (16,), 1, 1, 1, True),         # attention proj_out + residual
    # tcgen05 engine shapes (bf16): several M tiles and N blocks, ragged spatial dims, 64/32-wide K blocks, concat
    (3, 2, 128, 0, 256, (5, 9, 20), 3, 1, 1, True),
    (2, 3, 64, 64, 512, (24, 16), 3, 1, 1, False),
    (1, 2, 256, 0, 64, (100,), 3, 1, 1, False),
    (3, 1, 96, 32, 96, (4, 6, 16), 3, 1, 1, False),
    (2, 1, 512, 256, 256, (8, 8), 1, 1, 1, False),
    (3, 2, 128, 0, 128, (3, 8, 16), 3, 1, (1, 2, 2), False),   # Upsample on the tensor cores (4 parity classes, 12 taps)
    (3, 2, 64, 0, 64, (3, 10, 12), 3, (1, 2, 2), 1, False),    # Downsample on the tensor cores (parity tensor maps)
    (1, 2, 64, 0, 64, (37,), 3, 2, 1, False),                  # odd length, stride 2
    # tcgen05 weight-gradient engine on strided / fused-upsample convs (parity views of x resp. dy)
    (2, 2, 64, 0, 128, (9, 12), 3, 2, 1, False),               # Downsample 2-D, odd height
    (2, 1, 128, 0, 64, (5, 6), 3, 1, 2, False),                # Upsample 2-D (4 output-parity classes)
    (1, 2, 64, 0, 64, (19,), 3, 1, 2, False),                  # Upsample 1-D (2 classes)
    (3, 1, 256, 0, 256, (2, 4, 8), 3, 1, (1, 2, 2), False),    # Upsample 3-D, two ci blocks per tap
    # run kernels of the degenerate first / last convs (runs of 8 resp. 4 positions along W)
    (3, 2, 64, 0, 1, (4, 6, 12), 3, 1, 1, False),              # last conv of the 3-D configs (Cout = 1, 8 channel octets)
    (3, 2, 2, 0, 64, (3, 5, 16), 3, 1, 1, False),              # first conv, two input channels
    (1, 2, 1, 0, 64, (40,), 3, 1, 1, False),                   # first conv 1-D
    (1, 3, 128, 0, 2, (20,), 3, 1, 1, False),                  # last conv 1-D, 16 octets, two output channels
    (2, 2, 256, 0, 1, (6, 8), 3, 1, 1, False),                 # last conv 2-D, 32 octets
    # halo-reuse tcgen05 kernel (stride-1 3x3(x3), Cout 64 / 128, planes of >= 256 padded-flat positions): 30..35
    (3, 2, 64, 0, 64, (3, 20, 24), 3, 1, 1, True),             # 4 tiles per plane (one group), identity residual
    (3, 1, 128, 64, 64, (4, 33, 37), 3, 1, 1, False),          # concat input (3 K blocks), ragged plane, several groups
    (3, 2, 64, 0, 128, (2, 17, 16), 3, 1, 1, False),           # Cout 128 (two tiles per group), narrowest plane
    (2, 3, 64, 64, 128, (40, 48), 3, 1, 1, True),              # 2-D (kd = 1), concat, dgrad splits 128 channels 64 | 64
    (3, 1, 192, 0, 64, (5, 64, 64), 3, 1, 1, False),           # the 64^2-plane geometry of the 3-D UNet's top level
    (2, 1, 128, 0, 128, (31, 255), 3, 1, 1, False),            # widest supported row (Wp = 256)
    # every (h channels + skip channels -> Cout) combination of the 3-D 64^3 UNet's up path (config D), small spatial dims:
    # odd counts of 64-channel K blocks (192, 320) pair weight-gradient units across taps
    (3, 1, 128, 64, 64, (2, 8, 8), 3, 1, 1, False),
    (3, 1, 128, 64, 128, (2, 8, 8), 3, 1, 1, False),
    (3, 1, 256, 128, 128, (2, 8, 8), 3, 1, 1, False),
    (3, 1, 256, 128, 256, (2, 4, 8), 3, 1, 1, False),
    (3, 1, 512, 256, 256, (2, 4, 8), 3, 1, 1, False),
    (3, 1, 512, 256, 512, (2, 4, 4), 3, 1, 1, False),
    (3, 1, 512, 512, 512, (2, 4, 4), 3, 1, 1, False),
    (3, 1, 256, 64, 64, (2, 8, 8), 3, 1, 1, False),            # 320 = 5 K blocks
    (3, 1, 192, 0, 64, (2, 8, 8), 1, 1, 1, False),             # 1x1x1 skip conv on 3 K blocks
    # halo-reuse kernel with 32-channel K blocks (64-byte rows, SWIZZLE_64B) and 32 output channels (the full-resolution
    # levels of the 2-D deep-galaxy config, mc = 32)
    (2, 3, 32, 0, 32, (40, 48), 3, 1, 1, True),
    (2, 2, 64, 32, 32, (30, 33), 3, 1, 1, False),              # concat 64 + 32: three 32-channel K blocks
    (2, 2, 32, 0, 64, (24, 40), 3, 1, 1, False),
    (3, 1, 32, 0, 32, (3, 20, 24), 3, 1, 1, False),
    (2, 1, 64, 0, 32, (128, 128), 3, 1, 1, False),             # 64-channel K block, 32 outputs, the real 128^2 plane (17 groups)
    # 1x1 skip convs with 32-channel blocks: weight gradient on the halo kernel's one-tap mode
    (2, 2, 32, 32, 32, (40, 48), 1, 1, 1, False),
    (2, 2, 64, 32, 64, (30, 33), 1, 1, 1, False),
    (2, 1, 32, 0, 64, (64, 64), 1, 1, 1, False),
    (3, 1, 64, 32, 32, (2, 24, 24), 1, 1, 1, False),
    (3, 2, 128, 64, 128, (2, 17, 20), 1, 1, 1, True),          # 1x1x1 on 64-channel blocks through the halo kernel's streaming mode, residual
    # wide 1x1 convs in the streaming mode (128-channel output blocks, 1-D sequences re-factored into rows of <= 255 positions):
    # the attention blocks' qkv / proj_out at real widths; their dgrad streams 768 -> 256
    (1, 2, 256, 0, 768, (1024,), 1, 1, 1, False),
    (1, 2, 256, 0, 256, (640,), 1, 1, 1, True),
    (3, 1, 128, 128, 256, (2, 16, 24), 1, 1, 1, False),        # concat input, 3-D, two output blocks
    (1, 1, 512, 0, 1536, (333,), 1, 1, 1, False),              # T = 333 = 3 x 111 rows, 12 output blocks, 8 K blocks
    # T = 32768 tokens (the 128^3 config's attention at 128 x 16 x 16): beyond the tensor maps' 32767-coordinate limit, folded to
    # (2, 16384) by make_geom -- must stay on the tcgen05 engine (forward, dgrad, wgrad)
    (1, 1, 64, 0, 128, (32768,), 1, 1, 1, True),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES, ids=[str(i) for i in range(len(CONV_CASES))])
def test_conv_fwd_bwd(ops, dtype, case):
    dims, N, C1, C2, Cout, sp, k, stride, up, use_res = case
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[dims]
    q = (lambda t: t.to(dtype).float())
    x1 = q(randn(N, C1, *sp, seed=10)).requires_grad_(True)
    x2 = q(randn(N, C2, *sp, seed=11)).requires_grad_(True) if C2 else None
    Cin = C1 + C2
    w = (randn(Cout, Cin, *([k] * dims), seed=12) / math.sqrt(Cin * k ** dims)).requires_grad_(True)
    b = (0.1 * randn(Cout, seed=13)).requires_grad_(True)
    xin = torch.cat([x1, x2], 1) if C2 else x1
    upt = up if isinstance(up, tuple) else (up,) * dims
    if any(u != 1 for u in upt):
        xin = F.interpolate(xin, tuple(s * u for s, u in zip(sp, upt)), mode="nearest")
    wq = q(w) if dtype == torch.bfloat16 else w
    ref = conv(xin, wq, b, stride=stride, padding=k // 2)
    res = q(randn(*ref.shape, seed=14)) if use_res else None
    if use_res:
        ref = ref + res
    dy = q(randn(*ref.shape, seed=15))
    ref.backward(dy)

    spec = ops.ConvSpec.make(dims, k, stride, up)
    xd = cl(x1.detach()).to(dtype).requires_grad_(True)
    x2d = cl(x2.detach()).to(dtype).requires_grad_(True) if C2 else None
    wd, bd = w.detach().to(DEV).requires_grad_(True), b.detach().to(DEV).requires_grad_(True)
    resd = cl(res).to(dtype).requires_grad_(True) if use_res else None
    holder = torch.nn.Module()
    y = ops.conv(holder, xd, x2d, wd, bd, resd, spec, dtype, 0, False)
    assert tuple(nc(y).shape) == tuple(ref.shape)
    y.backward(cl(dy).to(dtype))
    tol = TOL[dtype]
    assert rel_err(nc(y), ref) < tol
    assert rel_err(nc(xd.grad), x1.grad) < 4 * tol
    if C2:
        assert rel_err(nc(x2d.grad), x2.grad) < 4 * tol
    assert rel_err(wd.grad, w.grad) < 4 * tol   # (bf16: gradient at the rounded weight values)
    assert rel_err(bd.grad, b.grad) < 4 * tol
    if use_res:
        assert rel_err(nc(resd.grad), dy) < 1e-6


@pytest.mark.parametrize("case", [CONV_CASES[i] for i in (3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 30, 31, 32, 33, 34, 35, 45, 46, 47, 48, 49, len(CONV_CASES) - 1)])
def test_conv_tcgen05_engine_is_used_and_matches_simt(ops, case):
    """The bf16 shapes above must run on the tcgen05 engine (forced: an unsupported shape would raise), and agree with
    the CUDA-core engine on the same inputs to fp32-accumulation rounding."""
    from rho_diffusion.b200 import _lib
    dims, N, C1, C2, Cout, sp, k, stride, up, use_res = case
    x1 = cl(randn(N, C1, *sp, seed=10)).bfloat16()
    x2 = cl(randn(N, C2, *sp, seed=11)).bfloat16() if C2 else None
    w = (randn(Cout, C1 + C2, *([k] * dims), seed=12) / math.sqrt((C1 + C2) * k ** dims)).to(DEV)
    b = (0.1 * randn(Cout, seed=13)).to(DEV)
    spec = ops.ConvSpec.make(dims, k, stride, up)
    holder = torch.nn.Module()
    wp = ops.packed_weight(holder, w, torch.bfloat16)
    res = torch.randn(ops.conv_out_shape(x1.shape, spec, Cout), device=DEV).bfloat16() if use_res else None
    aux = ops.aux_weight(holder, w, spec, x1, x2, wp, torch.bfloat16)
    assert (aux is not None) == (max(spec.up) > 1)
    y_tc = ops.conv_forward(x1, x2, wp, b, res, spec, Cout, engine=_lib.CONV_TCGEN05, w_aux=aux)
    y_simt = ops.conv_forward(x1, x2, wp, b, res, spec, Cout, engine=_lib.CONV_SIMT)
    assert rel_err(y_tc, y_simt) < 1e-2          # both round the same fp32 sums to bf16 (different summation order)
    # the fused-upsample path rounds the pre-summed weights (w1 + w2) to bf16 once instead of w1 and w2 separately
    mean_tol = 5e-3 if max(spec.up) > 1 else 1e-3
    assert float((y_tc.float() - y_simt.float()).abs().mean() / y_simt.float().abs().mean()) < mean_tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,T,heads,ch,order", [
    (2, 6, 1, 256, 0),      # reference fixture: heads=1, ch=256, T=6
    (2, 100, 4, 16, 1),     # ragged T, new order
    (1, 512, 4, 32, 0),
    (1, 130, 2, 64, 1),
    (1, 96, 4, 128, 1),     # config A head dim
    (1, 1000, 4, 64, 0),    # tcgen05 engine (bf16): 8 key tiles, ragged last tile, legacy order
    (2, 300, 2, 128, 1),    # tcgen05 engine (bf16): head dim 128, 3 key tiles
    (1, 128, 1, 64, 1),     # exactly one tile
    (2, 1, 2, 64, 0),       # a single position (the softmax of one logit)
    (1, 63, 2, 64, 1),      # one short of a key tile
    (1, 65, 1, 128, 0),     # one past a key tile
    (1, 4097, 2, 64, 1),    # 33 query blocks, one-row tail, long enough for lazy-max rescales
    (3, 257, 4, 128, 0),    # three samples, legacy order, head dim 128, odd T
])
def test_attention_fwd_bwd(ops, dtype, N, T, heads, ch, order):
    q = (lambda t: t.to(dtype).float())
    qkv = q(randn(N, 3 * heads * ch, T, seed=20)).requires_grad_(True)   # reference layout (N, 3HC, T)
    ref = O._attention(qkv, heads, bool(order))
    dy = q(randn(*ref.shape, seed=21))
    ref.backward(dy)
    qd = qkv.detach().permute(0, 2, 1).contiguous().to(DEV).to(dtype).requires_grad_(True)   # (N, T, 3HC)
    out = ops.attention(qd, heads, order)
    out.backward(dy.permute(0, 2, 1).contiguous().to(DEV).to(dtype))
    tol = TOL[dtype]
    assert rel_err(out.float().cpu().permute(0, 2, 1), ref) < tol
    assert rel_err(qd.grad.float().cpu().permute(0, 2, 1), qkv.grad) < 4 * tol


def test_embedding_path(ops):
    from rho_diffusion.b200 import _lib
    from rho_diffusion.models.common import omega_table
    t = torch.tensor([0, 1, 2, 17, 499, 999])
    for dim in (32, 64, 128):
        got = ops.timestep_embedding(t.to(DEV), dim, omega_table(dim).to(DEV))
        assert (got.cpu() - O.timestep_embedding(t, dim)).abs().max() < 2e-6
    # linear with fused input activation and add, fwd + bwd
    x = randn(5, 48, seed=30).requires_grad_(True)
    w = (randn(96, 48, seed=31) / 7).requires_grad_(True)
    b = randn(96, seed=32).requires_grad_(True)
    add = randn(5, 96, seed=33).requires_grad_(True)
    ref = F.linear(F.silu(x), w, b) + add
    dy = randn(5, 96, seed=34)
    ref.backward(dy)
    xd, wd, bd, ad = (v.detach().to(DEV).requires_grad_(True) for v in (x, w, b, add))
    y = ops.linear(xd, wd, bd, ad, act_in=_lib.ACT_SILU)
    y.backward(dy.to(DEV))
    assert rel_err(y, ref) < 1e-5
    for got, want in ((xd.grad, x.grad), (wd.grad, w.grad), (bd.grad, b.grad), (ad.grad, add.grad)):
        assert rel_err(got, want) < 1e-5
    # MultiEmbeddings: indices are exact, the sum matches the oracle
    from collections import OrderedDict
    from rho_diffusion.models.conditioning import MultiEmbeddings
    from conftest import load_golden
    from oracle.synth import synth_state_dict
    g = load_golden("embeddings.pt")["multi_emb"]
    me = MultiEmbeddings(parameter_space=OrderedDict(g["space"]), embedding_dim=16)
    me.load_state_dict(synth_state_dict(g["manifest"], seed=5))
    me = me.to(DEV)
    out = me(g["y"].to(DEV))
    assert (out.cpu() - g["out"]).abs().max() < 1e-6
    out.backward(torch.ones_like(out))
    gs = me.embedding_layers["s"].weight.grad.cpu()
    assert torch.equal(gs, torch.ones(3, 16))          # rows 0.5, 1.0, 0.25 each hit exactly once
    gt = me.embedding_layers["t"].weight.grad.cpu()
    assert torch.equal(gt[:, 0], torch.tensor([2.0, 1.0]))


@pytest.mark.parametrize("B,I,O,act", [(32, 48, 96, "silu"), (100, 130, 70, "none"), (512, 256, 1024, "silu"), (33, 7, 5, "silu"), (64, 64, 64, "none")])
def test_linear_large_batch_tiled(ops, B, I, O, act):
    """Batches >= 32 take the tiled fp32 GEMM kernels (config C feeds 512 rows to every emb_layers Linear): forward with the fused
    input activation / bias / add, dx, dw, db against torch, ragged tile edges included; a 512-row batch row-for-row equals the
    same rows pushed through as a batch of 33 (the K order does not depend on the grid)."""
    from rho_diffusion.b200 import _lib
    code = _lib.ACT_SILU if act == "silu" else _lib.ACT_NONE
    f = F.silu if act == "silu" else (lambda v: v)
    x = randn(B, I, seed=40).requires_grad_(True)
    w = (randn(O, I, seed=41) / math.sqrt(I)).requires_grad_(True)
    b = randn(O, seed=42).requires_grad_(True)
    add = randn(B, O, seed=43).requires_grad_(True)
    ref = F.linear(f(x), w, b) + add
    dy = randn(B, O, seed=44)
    ref.backward(dy)
    xd, wd, bd, ad = (v.detach().to(DEV).requires_grad_(True) for v in (x, w, b, add))
    y = ops.linear(xd, wd, bd, ad, act_in=code)
    y.backward(dy.to(DEV))
    assert rel_err(y, ref) < 1e-5
    for got, want in ((xd.grad, x.grad), (wd.grad, w.grad), (bd.grad, b.grad), (ad.grad, add.grad)):
        assert rel_err(got, want) < 1e-5
    with torch.no_grad():
        sub = ops.linear(xd[:33].contiguous(), wd, bd, ad[:33].contiguous(), act_in=code)
    if B >= 33:
        assert torch.equal(sub, y[:33])


def test_philox_bit_exact_and_normals(ops):
    import ctypes as C
    from rho_diffusion.b200 import _lib
    groups, seed, offset = 4096, 0x1234567890ABCDEF, 7
    out = torch.empty(groups * 4, dtype=torch.int32, device=DEV)
    _lib.check(_lib.lib().rho_philox_raw(out.data_ptr(), groups, seed, offset, ops.stream()), "philox_raw")
    got = out.cpu().numpy().view(np.uint32).reshape(groups, 4)
    assert np.array_equal(got, O.philox4x32_10(groups, offset, seed))
    # Random123 known-answer vector: counter = 0, key = 0
    _lib.check(_lib.lib().rho_philox_raw(out.data_ptr(), 1, 0, 0, ops.stream()), "philox_raw")
    assert [hex(v) for v in out[:4].cpu().numpy().view(np.uint32)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    n = 100003
    z = ops.randn((n,), DEV, seed=42, offset=3).cpu().numpy()
    ref = O.philox_normal(n, 3, 42)
    assert np.abs(z - ref).max() < 2e-5
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02


def test_q_sample_p_sample_mse(ops):
    sched = O.linear_schedule(100, 1e-3, 0.02)
    dev = {k: v.to(DEV) for k, v in sched.items()}
    x0, eps = randn(3, 1, 5, 7, 9, seed=40), randn(3, 1, 5, 7, 9, seed=41)   # 315 elements/sample: unaligned
    t = torch.tensor([0, 57, 99])
    got = ops.q_sample(x0.to(DEV), eps.to(DEV), t.to(DEV), dev["alpha_bar_t"])
    assert (got.cpu() - O.q_sample(sched, x0, t, eps)).abs().max() < 1e-6
    # ancestral update with external noise, all step classes (t = 0: identity, t = 1: no noise, t > 1)
    z = randn(3, 1, 5, 7, 9, seed=42)
    for step in (0, 1, 2, 50, 99):
        x = x0.clone().to(DEV)
        ts = torch.tensor([step], device=DEV)
        ops.p_sample_step(x, eps.to(DEV), z.to(DEV), ts, dev["alpha_t"], dev["beta_t"], dev["alpha_bar_t"])
        zz = z if step > 1 else torch.zeros_like(z)
        assert (x.cpu() - O.p_sample_step(sched, x0, eps, zz, step)).abs().max() < 2e-6, step
    # in-kernel Philox noise == external noise generated by rho_randn with (seed, offset = t)
    x_a, x_b = x0.clone().to(DEV), x0.clone().to(DEV)
    ts = torch.tensor([50], device=DEV)
    zk = ops.randn(x0.shape, DEV, seed=99, offset=50)
    ops.p_sample_step(x_a, eps.to(DEV), None, ts, dev["alpha_t"], dev["beta_t"], dev["alpha_bar_t"], seed=99)
    ops.p_sample_step(x_b, eps.to(DEV), zk, ts, dev["alpha_t"], dev["beta_t"], dev["alpha_bar_t"])
    assert torch.equal(x_a, x_b)
    # timestep advance
    tsc, tv = torch.tensor([7], device=DEV), torch.zeros(5, dtype=torch.int64, device=DEV)
    ops.advance_timestep(tsc, tv)
    assert int(tsc) == 6 and tv.tolist() == [6] * 5
    # MSE + gradient
    for dtype in (torch.float32, torch.bfloat16):
        pred = randn(4, 1000, seed=43).to(dtype)
        tgt = randn(4, 1000, seed=44)
        pr = pred.float().clone().requires_grad_(True)
        ref = O.mse_loss(pr, tgt)
        (ref * 3.0).backward()
        pd = pred.detach().clone().to(DEV).requires_grad_(True)
        loss = ops.MseLossFn.apply(pd, tgt.to(DEV))
        (loss * 3.0).backward()
        assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, float(ref))
        assert rel_err(pd.grad, pr.grad) < TOL[dtype]


def test_adamw_ema_kernel(ops):
    n = 10007
    p, g = randn(n, seed=50), randn(n, seed=51) * 0.1
    m, v, sh = torch.zeros(n), torch.zeros(n), randn(n, seed=52)
    pd, gd, md, vd, sd = (t.clone().to(DEV) for t in (p, g, m, v, sh))
    for step in (1, 2, 3):
        p, m, v = O.adamw_step(p, g * 0.5, m, v, step, lr=2e-4, weight_decay=1e-2)
        sh = O.ema_update(sh, p, step)
        ops.adamw_ema_step(pd, gd, md, vd, sd, step, 2e-4, (0.9, 0.999), 1e-8, 1e-2, ema_decay_t=O.ema_decay(step), grad_scale=0.5)
    assert (pd.cpu() - p).abs().max() < 1e-6
    assert (sd.cpu() - sh).abs().max() < 1e-6
    assert rel_err(md, m) < 2e-6 and rel_err(vd, v) < 2e-6


@pytest.mark.parametrize("case", [
    # (dims, N, Cin, Cout, spatial, k, stride, up, residual, expect_fused)
    (3, 2, 64, 64, (4, 20, 24), 3, 1, 1, True, True),            # halo-reuse kernel (Cout 64)
    (3, 2, 64, 128, (2, 17, 16), 3, 1, 1, False, True),          # halo-reuse kernel (Cout 128)
    (3, 2, 128, 256, (5, 9, 20), 3, 1, 1, True, True),           # generic kernel, ragged tiles, TMA-store epilogue
    (3, 2, 128, 128, (3, 8, 16), 3, 1, (1, 2, 2), False, True),  # fused upsample: four parity-class launches
    (3, 2, 64, 64, (3, 10, 12), 3, (1, 2, 2), 1, False, True),   # stride 2
    (1, 2, 256, 256, (300,), 1, 1, 1, True, True),               # attention proj_out (k = 1) + residual
    (2, 3, 64, 512, (24, 16), 3, 1, 1, False, True),             # two 256-channel blocks per tile
    (3, 2, 64, 64, (2, 4, 4), 3, 1, 1, False, False),            # 32 positions per sample: tiles straddle samples -> not fused
    (3, 1, 64, 256, (4, 64, 65), 3, 1, 1, False, True),          # > 1024 (tile, channel) pairs per group: the CTA-wide merge kernel
    (2, 3, 32, 32, (40, 48), 3, 1, 1, True, True),               # 32-channel halo kernel (one channel per group)
])
def test_conv_epilogue_groupnorm_statistics(ops, case):
    """The (mean, rstd) GroupNorm32 obtains from the channel sums a conv epilogue leaves behind equal those of the
    streaming statistics kernel on the same bf16 tensor, alone and as either half of a channel concat."""
    dims, N, Cin, Cout, sp, k, stride, up, use_res, expect = case
    spec = ops.ConvSpec.make(dims, k, stride, up)
    x = cl(randn(N, Cin, *sp, seed=20) * 1.5 + 0.7).bfloat16()
    w = (randn(Cout, Cin, *([k] * dims), seed=21) / math.sqrt(Cin * k ** dims)).to(DEV)
    b = (0.5 * randn(Cout, seed=22)).to(DEV)
    holder = torch.nn.Module()
    res = torch.randn(ops.conv_out_shape(x.shape, spec, Cout), device=DEV).bfloat16() if use_res else None
    w2 = (randn(64, Cin, *([k] * dims), seed=23) / math.sqrt(Cin * k ** dims)).to(DEV)
    with torch.inference_mode():
        y = ops.conv(holder, x, None, w, b, res, spec, torch.bfloat16)
        assert hasattr(y, "_rho_part") == expect
        if not expect:
            return
        gamma, beta = torch.ones(Cout, device=DEV), torch.zeros(Cout, device=DEV)
        out_f, st_f = ops.gn_forward(y, None, gamma, beta, None, None, 1, torch.bfloat16)
        plain = y.clone()                                    # no side channel: the streaming kernel runs
        out_k, st_k = ops.gn_forward(plain, None, gamma, beta, None, None, 1, torch.bfloat16)
        assert (st_f[..., 0] - st_k[..., 0]).abs().max() < 2e-5 * (1 + st_k[..., 0].abs().max())
        assert ((st_f[..., 1] - st_k[..., 1]).abs() / st_k[..., 1]).max() < 2e-4
        assert rel_err(out_f, out_k) < 5e-3
        # concat: y as x (fused) and a second conv output as x2 (fused), groups straddling the boundary when C1 % cpg != 0
        y2 = ops.conv(torch.nn.Module(), x, None, w2, None, None, spec, torch.bfloat16)
        Ct = Cout + 64
        if Ct % 32 == 0 and hasattr(y2, "_rho_part"):
            g2, b2 = torch.ones(Ct, device=DEV), torch.zeros(Ct, device=DEV)
            _, st_cf = ops.gn_forward(y, y2, g2, b2, None, None, 1, torch.bfloat16)
            _, st_ck = ops.gn_forward(y.clone(), y2.clone(), g2, b2, None, None, 1, torch.bfloat16)
            assert (st_cf[..., 0] - st_ck[..., 0]).abs().max() < 2e-5 * (1 + st_ck[..., 0].abs().max())
            assert ((st_cf[..., 1] - st_ck[..., 1]).abs() / st_ck[..., 1]).max() < 2e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["SymmetricLog", "GELU", "Tanh"])
def test_groupnorm_registered_activations_fwd_bwd(ops, dtype, name):
    """Activations other than SiLU go through the generic epilogue of the fused GroupNorm kernel, incl. the reference's own
    SymmetricLog (rho_diffusion/activations.py:35-61)."""
    from oracle.torch_oracle import _ACTS
    N, C, sp = 2, 64, (5, 6)
    x = randn(N, C, *sp, seed=31) * 1.3 + 0.2
    gamma, beta = 1 + 0.1 * randn(C, seed=32), 0.1 * randn(C, seed=33)
    dy = randn(N, C, *sp, seed=34)
    q = (lambda t: t.to(dtype).float())
    xr, gr, br = q(x).clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = _ACTS[name](F.group_norm(xr, 32, gr, br, 1e-5))
    ref.backward(q(dy))
    xd = cl(x).to(dtype).requires_grad_(True)
    gd, bd = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y = ops.GnActFn.apply(xd, None, gd, bd, None, None, ops.act_code(name), dtype, 32, 1e-5)
    y.backward(cl(dy).to(dtype))
    tol = TOL[dtype]
    assert rel_err(nc(y), ref) < tol
    assert rel_err(nc(xd.grad), xr.grad) < 4 * tol
    assert rel_err(gd.grad, gr.grad) < 4 * tol and rel_err(bd.grad, br.grad) < 4 * tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,factors", [((2, 32, 9, 12), (2, 2)), ((2, 16, 3, 8, 10), (1, 2, 2)), ((3, 64, 33), (2,)),
                                           ((1, 8, 4, 4, 4), (2, 2, 2))])
def test_avg_pool_and_nearest_upsample_fwd_bwd(ops, dtype, shape, factors):
    """avg_pool_nd (kernel = stride, floor extents) and the conv-less nearest upsample with their adjoints vs torch."""
    dims = len(shape) - 2
    x = randn(*shape, seed=41)
    q = (lambda t: t.to(dtype).float())
    pool = {1: F.avg_pool1d, 2: F.avg_pool2d, 3: F.avg_pool3d}[dims]
    xr = q(x).clone().requires_grad_(True)
    ref = pool(xr, factors, factors)
    dy = q(randn(*ref.shape, seed=42))
    ref.backward(dy)
    xd = cl(x).to(dtype).requires_grad_(True)
    y = ops.avg_pool(xd, factors)
    y.backward(cl(dy).to(dtype))
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    assert tuple(nc(y).shape) == tuple(ref.shape)
    assert rel_err(nc(y), ref) < tol and rel_err(nc(xd.grad), xr.grad) < tol
    xr2 = q(x).clone().requires_grad_(True)
    ref2 = F.interpolate(xr2, tuple(s * f for s, f in zip(shape[2:], factors)), mode="nearest")
    dy2 = q(randn(*ref2.shape, seed=43))
    ref2.backward(dy2)
    xd2 = cl(x).to(dtype).requires_grad_(True)
    y2 = ops.upsample_nearest(xd2, factors)
    y2.backward(cl(dy2).to(dtype))
    assert torch.equal(nc(y2).float().cpu(), ref2.detach())                  # a pure gather: exact
    assert rel_err(nc(xd2.grad), xr2.grad) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dropout_mask_scaling_and_backward(ops, dtype):
    """nn.Dropout(p) semantics: elements are zeroed with probability p, survivors scaled by 1 / (1 - p); the backward applies
    the SAME mask; a fixed (seed, call index) reproduces the mask and successive calls draw fresh ones."""
    from rho_diffusion.b200 import ops as O2
    p = 0.3
    x = (torch.rand(4, 1000, 37, device=DEV) + 0.5).to(dtype).requires_grad_(True)
    O2._dropout_calls = 100
    y = ops.dropout(x, p, seed=5)
    y.backward(torch.ones_like(y))
    keep = y.detach().float() != 0
    frac = float((~keep).float().mean())
    assert abs(frac - p) < 5e-3
    assert rel_err(y.detach()[keep], x.detach()[keep].float() / (1 - p)) < (1e-6 if dtype == torch.float32 else 8e-3)
    assert torch.equal(x.grad != 0, keep)
    assert rel_err(x.grad[keep], torch.full_like(x.grad[keep], 1 / (1 - p))) < (1e-6 if dtype == torch.float32 else 8e-3)
    O2._dropout_calls = 100
    assert torch.equal(ops.dropout(x.detach(), p, seed=5), y.detach())       # same (seed, call index) -> same mask
    assert not torch.equal(ops.dropout(x.detach(), p, seed=5) != 0, keep)    # next call index -> a different mask
    assert torch.equal(ops.dropout(x.detach(), 0.0, seed=1), x.detach())


def _fuzz_conv_cases(n, seed):
    import random
    rng = random.Random(seed)
    cases = []
    while len(cases) < n:
        dims = rng.choice([1, 2, 3, 3])
        k = rng.choice([1, 3, 3, 3])
        C1 = rng.choice([32, 64, 64, 96, 128, 192, 256, 320])
        C2 = rng.choice([0, 0, 0, 32, 64, 128])
        Cout = rng.choice([32, 64, 64, 96, 128, 128, 256, 384])
        mode = rng.choice(["plain", "plain", "plain", "stride", "up"]) if k == 3 else "plain"
        if mode != "plain":
            C2 = 0
        if dims == 1:
            sp = (rng.randint(5, 300),)
        elif dims == 2:
            sp = (rng.randint(3, 40), rng.randint(3, 70))
        else:
            sp = (rng.randint(1, 5), rng.randint(3, 24), rng.randint(3, 40))
        stride = ((1, 2, 2) if dims == 3 else 2) if mode == "stride" else 1
        up = ((1, 2, 2) if dims == 3 else 2) if mode == "up" else 1
        N = rng.randint(1, 3)
        if N * math.prod(sp) * (C1 + C2) > 3_000_000:
            continue
        cases.append((dims, N, C1, C2, Cout, sp, k, stride, up, rng.random() < 0.3 and mode == "plain"))
    return cases


@pytest.mark.parametrize("case", _fuzz_conv_cases(48, seed=20261018), ids=lambda c: "-".join(str(v) for v in c[:5]) + "-" + "x".join(map(str, c[5])))
def test_conv_engines_agree_on_random_shapes(ops, case):
    """Seeded shape fuzz: forward, data gradient and weight gradient of the tensor-core engines (generic per-tap, halo-reuse,
    fused-upsample, strided, their weight-gradient counterparts) against the CUDA-core engine on the same bf16 inputs --
    odd / ragged extents, odd counts of 64-channel blocks, concat inputs, residuals."""
    from rho_diffusion.b200 import _lib
    dims, N, C1, C2, Cout, sp, k, stride, up, use_res = case
    spec = ops.ConvSpec.make(dims, k, stride, up)
    x1 = cl(randn(N, C1, *sp, seed=50)).bfloat16()
    x2 = cl(randn(N, C2, *sp, seed=51)).bfloat16() if C2 else None
    w = (randn(Cout, C1 + C2, *([k] * dims), seed=52) / math.sqrt((C1 + C2) * k ** dims)).to(DEV)
    b = (0.1 * randn(Cout, seed=53)).to(DEV)
    holder = torch.nn.Module()
    wp = ops.packed_weight(holder, w, torch.bfloat16)
    aux = ops.aux_weight(holder, w, spec, x1, x2, wp, torch.bfloat16)
    oshape = ops.conv_out_shape(x1.shape, spec, Cout)
    res = (0.5 * torch.randn(oshape, device=DEV)).bfloat16() if use_res else None
    y_a = ops.conv_forward(x1, x2, wp, b, res, spec, Cout, engine=_lib.CONV_AUTO, w_aux=aux)
    y_s = ops.conv_forward(x1, x2, wp, b, res, spec, Cout, engine=_lib.CONV_SIMT)
    assert rel_err(y_a, y_s) < 1.5e-2
    assert float((y_a.float() - y_s.float()).abs().mean() / y_s.float().abs().mean()) < (5e-3 if max(spec.up) > 1 else 1e-3)
    dy = torch.randn(oshape, device=DEV).bfloat16()
    ga = ops.conv_backward(x1, x2, wp, dy, spec, Cout, True, True, True, tuple(w.shape), engine=_lib.CONV_AUTO)
    gs = ops.conv_backward(x1, x2, wp, dy, spec, Cout, True, True, True, tuple(w.shape), engine=_lib.CONV_SIMT)
    for a, s, name in zip(ga, gs, ("dx", "dx2", "dw", "db")):
        if s is None:
            assert a is None
            continue
        assert rel_err(a, s) < 2e-2, name
        assert float((a.float() - s.float()).abs().mean() / s.float().abs().mean().clamp_min(1e-12)) < 3e-3, name


@pytest.mark.parametrize("cin", [32, 128, 512])
def test_split_bf16_fp32_conv_is_fp32_accurate_at_any_contraction_length(ops, cin):
    """The split-bf16 engine (fp32 conv as six bf16 tensor-core products, fp32 partial sums flushed to registers every 32 MMAs) against
    an fp64 reference: <= 2e-6 relative whatever the contraction length -- the un-flushed tensor-core accumulation truncates and
    drifts linearly with K (5e-6 at K = 864, 1.5e-4 at K = 27648, tools/fp32_split_error.py).  Strided and concatenated inputs too."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(cin)
    x = torch.randn(1, cin, 4, 16, 16, generator=gen)
    x2 = torch.randn(1, 32, 4, 16, 16, generator=gen)
    w = torch.randn(64, cin + 32, 3, 3, 3, generator=gen) / (27 * (cin + 32)) ** 0.5
    b = torch.randn(64, generator=gen)
    cl_ = lambda t: t.permute(0, 2, 3, 4, 1).contiguous().to(DEV)      # noqa: E731
    holder = torch.nn.Module()
    try:
        ops.set_fp32_split(True)
        with torch.no_grad():
            for stride in (1, (1, 2, 2)):
                ref = F.conv3d(torch.cat([x, x2], 1).double(), w.double(), b.double(), stride=stride, padding=1)
                y = ops.conv(holder, cl_(x), cl_(x2), w.to(DEV), b.to(DEV), None, ops.ConvSpec.make(3, 3, stride, 1), torch.float32)
                got = y.permute(0, 4, 1, 2, 3).double().cpu()
                assert float((got - ref).norm() / ref.norm()) < 2e-6, (cin, stride)
    finally:
        ops.set_fp32_split(False)
