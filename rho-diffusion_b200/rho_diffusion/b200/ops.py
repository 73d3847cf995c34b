"""Host-side wrappers of the C ABI (``include/rho_b200.h``): tensor plumbing + autograd glue.

PyTorch is used here for device memory, streams and the autograd tape only; every arithmetic
operation is a kernel of librho_b200.so.  Activations inside the model are **channels-last**
tensors of shape ``(N, *spatial, C)`` (contiguous); the reference's ``(N, C, *spatial)`` layout only
exists at the UNet boundary (:func:`nc_to_cl` / :func:`cl_to_nc`).

There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib as L

F32, BF16 = L.F32, L.BF16
_DT = {torch.float32: F32, torch.bfloat16: BF16}
_ACT_BY_NAME = {
    None: L.ACT_NONE, "Identity": L.ACT_NONE, "SiLU": L.ACT_SILU, "ReLU": L.ACT_RELU, "Tanh": L.ACT_TANH,
    "Sigmoid": L.ACT_SIGMOID, "GELU": L.ACT_GELU, "SymmetricLog": L.ACT_SYMLOG,
}


def act_code(act) -> int:
    """Maps a registry activation (``nn.SiLU()`` instance, class or name) to the kernel enum."""
    if act is None or isinstance(act, int):
        return L.ACT_NONE if act is None else act
    name = act if isinstance(act, str) else (act.__name__ if isinstance(act, type) else type(act).__name__)
    if name not in _ACT_BY_NAME:
        raise L.RhoB200Error(f"activation {name!r} has no fused B200 kernel (supported: SiLU, ReLU, Tanh, Sigmoid, GELU, SymmetricLog)")
    if name == "GELU" and getattr(act, "approximate", "none") != "none":
        raise L.RhoB200Error("GELU(approximate='tanh') is not supported")
    return _ACT_BY_NAME[name]


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise L.RhoB200Error(f"unsupported dtype {t.dtype}; the B200 path computes in float32 or bfloat16") from None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.RhoB200Error("rho_diffusion (B200) has no CPU path: tensors must be on a CUDA device")


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


_ws = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Per-device scratch buffer, grown on demand (stream-ordered reuse)."""
    key = (device.index, torch.cuda.is_current_stream_capturing())
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------------
def nc_to_cl(x: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """(N, C, *sp) -> channels-last (N, *sp, C) in ``out_dtype``."""
    _need_cuda(x)
    x = x.contiguous()
    N, Cc, *sp = x.shape
    S = 1
    for s in sp:
        S *= s
    y = torch.empty((N, *sp, Cc), dtype=out_dtype, device=x.device)
    L.check(L.lib().rho_nc_to_cl(x.data_ptr(), dt(x), y.data_ptr(), dt(y), N, Cc, S, stream()), "nc_to_cl")
    return y


def cl_to_nc(x: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    _need_cuda(x)
    x = x.contiguous()
    N, *sp, Cc = x.shape
    S = 1
    for s in sp:
        S *= s
    y = torch.empty((N, Cc, *sp), dtype=out_dtype, device=x.device)
    L.check(L.lib().rho_cl_to_nc(x.data_ptr(), dt(x), y.data_ptr(), dt(y), N, Cc, S, stream()), "cl_to_nc")
    return y


class _LayoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, to_cl, out_dtype):
        ctx.to_cl, ctx.in_dtype = to_cl, x.dtype
        return nc_to_cl(x, out_dtype) if to_cl else cl_to_nc(x, out_dtype)

    @staticmethod
    def backward(ctx, g):
        return (cl_to_nc(g, ctx.in_dtype) if ctx.to_cl else nc_to_cl(g, ctx.in_dtype)), None, None


def to_channels_last(x, out_dtype):
    return _LayoutFn.apply(x, True, out_dtype)


def to_channels_first(x, out_dtype):
    return _LayoutFn.apply(x, False, out_dtype)


def _pad3(t) -> Tuple[int, int, int]:
    t = tuple(int(v) for v in t)
    return (1,) * (3 - len(t)) + t


# ------------------------------------------------------------------------------------------------
# conv_nd
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ConvSpec:
    """Geometry of one conv_nd call site (kernel, stride and fused-upsample factors per spatial dim)."""
    dims: int
    kernel: Tuple[int, ...]
    stride: Tuple[int, ...]
    up: Tuple[int, ...]

    @staticmethod
    def make(dims, kernel=3, stride=1, up=1):
        def tup(v):
            return tuple(v) if isinstance(v, (tuple, list)) else (v,) * dims
        return ConvSpec(dims, tup(kernel), tup(stride), tup(up))


_weights_epoch = 0


def bump_weights_epoch():
    """Called by the fused optimizer after it has rewritten parameter memory behind autograd's back."""
    global _weights_epoch
    _weights_epoch += 1


class GradSink:
    """Destination of one parameter's gradient inside the fused optimizer's flat fp32 gradient buffer.

    The backward kernels (conv wgrad, GroupNorm / Linear / embedding backward) write the parameter gradient STRAIGHT into
    ``view`` -- a slice of the flat buffer shaped like the parameter -- and the autograd Function returns ``None`` for it, so
    there is neither a per-parameter accumulate launch nor a gather copy.  First touch per step overwrites (no zero pass);
    a parameter that receives a second contribution before ``zero_grad`` falls back to autograd's accumulation, which adds
    into the same view (``p.grad`` IS the view).  ``done`` notifies the optimizer, which launches the all-reduce of a
    bucket as soon as the last gradient of that bucket has been enqueued (overlap with the rest of the backward)."""
    __slots__ = ("view", "owner", "bucket", "touched")

    def __init__(self, view, owner, bucket):
        self.view, self.owner, self.bucket, self.touched = view, owner, bucket, False

    def done(self):
        self.touched = True
        self.owner._grad_ready(self.bucket)


def grad_sink(param, needed: bool = True):
    """The untouched sink of ``param`` (a fused-optimizer parameter), or None (plain autograd path)."""
    if not needed or param is None:
        return None
    sink = getattr(param, "_rho_sink", None)
    if sink is None or sink.touched:
        return None
    return sink


def packed_weight(owner, weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[Cout, Cin, *k] fp32 (state-dict layout) -> [Cout, taps, Cin] in ``dtype``.

    The packed copy is cached ON the owning module (``owner``), keyed by the parameter's version counter, its
    address and the optimizer epoch -- never in a global table, where a recycled device address of a different
    model could alias a stale entry."""
    cache = owner.__dict__.setdefault("_rho_packed", {})
    # (an inference tensor -- a weight computed by a kernel under inference_mode -- tracks no version; such weights are never cached)
    ver = (-1 if weight.is_inference() else weight._version, _weights_epoch, weight.data_ptr(), tuple(weight.shape))
    hit = cache.get(dtype)
    if hit is not None and hit[0] == ver:
        return hit[1]
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    Cout, Cin = w.shape[0], w.shape[1]
    taps = w.numel() // (Cout * Cin)
    if hit is not None and hit[1].shape == (Cout, taps, Cin) and hit[1].device == w.device:
        out = hit[1]
    else:
        with torch.inference_mode(False):   # cached across inference and training: must not be an inference tensor
            out = torch.empty((Cout, taps, Cin), dtype=dtype, device=w.device)
    L.check(L.lib().rho_conv_pack_weight(w.data_ptr(), out.data_ptr(), _DT[dtype], Cout, Cin, taps, stream()), "conv_pack_weight")
    cache[dtype] = (ver, out)
    return out


def _conv_params(x, x2, wp, bias, residual, y, spec: ConvSpec, Cout, engine, y_f32, w_aux=None) -> L.ConvParams:
    N = x.shape[0]
    D, H, W = _pad3(x.shape[1:-1])
    kd, kh, kw = _pad3(spec.kernel)
    sd, sh, sw = _pad3(spec.stride)
    ud, uh, uw = _pad3(spec.up)
    p = L.ConvParams()
    p.N, p.D, p.H, p.W = N, D, H, W
    p.C1, p.C2, p.Cout = x.shape[-1], (x2.shape[-1] if x2 is not None else 0), Cout
    p.kd, p.kh, p.kw, p.sd, p.sh, p.sw, p.ud, p.uh, p.uw = kd, kh, kw, sd, sh, sw, ud, uh, uw
    p.act_dtype, p.w_dtype, p.engine, p.y_f32 = dt(x), dt(wp), engine, int(y_f32)
    p.x, p.x2, p.w_packed, p.w_aux = ptr(x), ptr(x2), ptr(wp), ptr(w_aux)
    p.bias, p.residual, p.y = ptr(bias), ptr(residual), ptr(y)
    return p


def conv_out_shape(x_shape, spec: ConvSpec, Cout):
    sp = x_shape[1:-1]
    out = []
    for s, k, st, u in zip(sp, spec.kernel, spec.stride, spec.up):
        out.append((s * u + 2 * (k // 2) - k) // st + 1)
    return (x_shape[0], *out, Cout)


class KernelProfiler:
    """Optional per-launch timing with CUDA events on the launching stream (used by bench.py for the roofline)."""

    def __init__(self):
        self.records = []   # (kind, name, algorithmic flops, algorithmic bytes, start event, end event)

    def begin(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        return ev

    def end(self, kind, name, flops, nbytes, start):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        self.records.append((kind, name, flops, nbytes, start, ev))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, name, flops, nbytes, s, e in self.records:
            d = out.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


profiler: Optional[KernelProfiler] = None


def conv_forward(x, x2, wp, bias, residual, spec: ConvSpec, Cout, engine=L.CONV_AUTO, y_f32=False, w_aux=None, stats=None):
    """``stats``: optional list that receives the fused GroupNorm statistics partials of the output, ``(tensor, rows)``
    (rho_conv_params.stats_partial), when the launch can produce them."""
    if profiler is not None:
        ev = profiler.begin()
        y = _conv_forward(x, x2, wp, bias, residual, spec, Cout, engine, y_f32, w_aux, stats)
        taps = 1
        for k in spec.kernel:
            taps *= k
        cin = x.shape[-1] + (x2.shape[-1] if x2 is not None else 0)
        m = y.numel() // Cout
        nbytes = x.numel() * x.element_size() + (x2.numel() * x2.element_size() if x2 is not None else 0) + y.numel() * y.element_size() \
            + wp.numel() * wp.element_size() + (residual.numel() * residual.element_size() if residual is not None else 0)
        tag = ("_s2" if max(spec.stride) > 1 else "") + ("_up" if max(spec.up) > 1 else "")
        profiler.end("conv", f"conv{spec.dims}d_k{spec.kernel[-1]}{tag}_{cin}->{Cout}", 2.0 * m * Cout * cin * taps, nbytes, ev)
        return y
    return _conv_forward(x, x2, wp, bias, residual, spec, Cout, engine, y_f32, w_aux, stats)


def _conv_forward(x, x2, wp, bias, residual, spec: ConvSpec, Cout, engine=L.CONV_AUTO, y_f32=False, w_aux=None, stats=None):
    _need_cuda(x, x2, wp, bias, residual)
    assert x.is_contiguous() and (x2 is None or x2.is_contiguous()) and (residual is None or residual.is_contiguous())
    out_dtype = torch.float32 if y_f32 else x.dtype
    y = torch.empty(conv_out_shape(x.shape, spec, Cout), dtype=out_dtype, device=x.device)
    if residual is not None:
        assert residual.shape == y.shape and residual.dtype == x.dtype
    p = _conv_params(x, x2, wp, bias, residual, y, spec, Cout, engine, y_f32, w_aux)
    if stats is not None and x.dtype == torch.bfloat16 and not y_f32:
        rows = int(L.lib().rho_conv_stats_rows(C.byref(p)))
        if rows > 0:
            part = torch.empty((x.shape[0], rows, Cout, 2), dtype=torch.float32, device=x.device)
            p.stats_partial = part.data_ptr()
            stats.append((part, rows))
    L.check(L.lib().rho_conv_fwd(C.byref(p), stream()), "conv_fwd")
    return y


def conv_backward(x, x2, wp, dy, spec: ConvSpec, Cout, need_dx, need_dw, need_db, w_shape, engine=L.CONV_AUTO, dw_out=None, db_out=None):
    """``dw_out`` / ``db_out``: optional fp32 destinations (flat-gradient slices) the weight-gradient kernels write directly."""
    dy = dy.contiguous()
    bp = L.ConvBwdParams()
    bp.fwd = _conv_params(x, x2, wp, None, None, None, spec, Cout, engine, False)
    bp.dy = ptr(dy)
    nbytes = L.lib().rho_conv_bwd_workspace_bytes(C.byref(bp))
    ws = workspace(nbytes, x.device)
    bp.workspace, bp.workspace_bytes = ws.data_ptr(), ws.numel()
    dx = dx2 = dw = db = None
    if need_dx:
        dx = torch.empty_like(x)
        dx2 = torch.empty_like(x2) if x2 is not None else None
        bp.dx, bp.dx2 = ptr(dx), ptr(dx2)
        L.check(L.lib().rho_conv_dgrad(C.byref(bp), stream()), "conv_dgrad")
    if need_dw or need_db:
        if need_dw:
            dw = dw_out if dw_out is not None else torch.empty(w_shape, dtype=torch.float32, device=x.device)
        if need_db:
            db = db_out if db_out is not None else torch.empty((Cout,), dtype=torch.float32, device=x.device)
        bp.dw, bp.dbias = ptr(dw), ptr(db)
        L.check(L.lib().rho_conv_wgrad(C.byref(bp), stream()), "conv_wgrad")
    return dx, dx2, dw, db


class ConvFn(torch.autograd.Function):
    """y = conv_nd(cat(x, x2)) + bias (+ residual); see rho_conv_fwd in include/rho_b200.h.
    ``wp`` is the packed copy of ``weight`` (:func:`packed_weight`); ``weight`` itself only receives the gradient."""

    @staticmethod
    def forward(ctx, x, x2, weight, wp, bias, residual, spec, engine, y_f32, w_aux=None, stats=None):
        Cout = weight.shape[0]
        b = bias.detach().float() if bias is not None else None
        y = conv_forward(x, x2, wp, b, residual, spec, Cout, engine, y_f32, w_aux, stats)
        ctx.save_for_backward(x, x2, wp)
        ctx.spec, ctx.engine, ctx.has_bias, ctx.has_res = spec, engine, bias is not None, residual is not None
        ctx.w_shape = tuple(weight.shape)
        ctx.params = (weight, bias)          # the Parameters themselves: their gradient sinks are looked up in backward
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x2, wp = ctx.saved_tensors
        if dy.dtype != x.dtype:  # y_f32 head: gradient arrives in fp32
            dy = cast(dy, x.dtype)
        need_dx = ctx.needs_input_grad[0] or (x2 is not None and ctx.needs_input_grad[1])
        need_dw, need_db = ctx.needs_input_grad[2], ctx.has_bias and ctx.needs_input_grad[4]
        sw, sb = grad_sink(ctx.params[0], need_dw), grad_sink(ctx.params[1], need_db)
        dx, dx2, dw, db = conv_backward(x, x2, wp, dy, ctx.spec, ctx.w_shape[0], need_dx, need_dw, need_db, ctx.w_shape, ctx.engine,
                                        dw_out=sw.view if sw else None, db_out=sb.view if sb else None)
        if sw:
            sw.done(); dw = None
        if sb:
            sb.done(); db = None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[5]) else None
        return dx, dx2, dw, None, db, dres, None, None, None, None, None


def aux_weight(owner, weight, spec: ConvSpec, x, x2, wp, dtype):
    """Second, engine-specific packing (pre-summed weights of the fused-upsample decomposition); None if the geometry
    has none.  Cached on the owning module like :func:`packed_weight`."""
    if dtype != torch.bfloat16 or max(spec.up) == 1:
        return None
    p = _conv_params(x, x2, wp, None, None, None, spec, weight.shape[0], L.CONV_AUTO, False)
    nbytes = L.lib().rho_conv_aux_weight_bytes(C.byref(p))
    if nbytes == 0:
        return None
    cache = owner.__dict__.setdefault("_rho_aux", {})
    ver = (weight._version, _weights_epoch, weight.data_ptr(), spec, nbytes)
    hit = cache.get("aux")
    if hit is not None and hit[0] == ver:
        return hit[1]
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    with torch.inference_mode(False):
        out = hit[1] if (hit is not None and hit[1].numel() == nbytes // 2) else torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    L.check(L.lib().rho_conv_pack_weight_aux(C.byref(p), w.data_ptr(), out.data_ptr(), stream()), "conv_pack_weight_aux")
    cache["aux"] = (ver, out)
    return out


# ---- split-bf16 engine: fp32-accurate convolutions on the bf16 tensor cores (inference) ----------------------------------
_fp32_split = None


def fp32_split_enabled() -> bool:
    """``RHO_FP32_SPLIT`` = 0 | 1 (default 0; also ``ops.set_fp32_split``): under ``torch.no_grad`` the fp32 path runs its
    convolutions as six bf16 tensor-core products (x = x1 + x2 + x3, w = w1 + w2 + w3, all terms above 2^-24 kept) accumulated in
    fp32 with a register flush every 32 MMAs instead of the CUDA-core FFMA engine: 2.5x faster on config D and CLOSER to the fp64
    result (5-7e-7 relative per conv at any contraction length vs 5e-7 ... 3e-6), but not bit-compatible with it -- a 50-step
    shared-noise trajectory drifts 1.4e-3 from the reference's instead of < 1e-3 -- so the exact path keeps the CUDA-core engine
    by default.  Training (autograd) always uses the CUDA-core engine, whose backward shares its operands."""
    global _fp32_split
    if _fp32_split is None:
        import os
        _fp32_split = os.environ.get("RHO_FP32_SPLIT", "0") == "1" and os.environ.get("RHO_DISABLE_TCGEN05", "0") != "1"
    return _fp32_split


def set_fp32_split(on: bool) -> None:
    global _fp32_split
    _fp32_split = bool(on)


def split_packed_weight(owner, weight: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, *k] fp32 -> packed bf16 [Cout, taps, 6 Cin] of the bf16-exact terms [w1 | w2 | w1 | w3 | w2 | w1] (cached like
    :func:`packed_weight`)."""
    cache = owner.__dict__.setdefault("_rho_packed", {})
    ver = (weight._version, _weights_epoch, weight.data_ptr(), tuple(weight.shape))
    hit = cache.get("split6")
    if hit is not None and hit[0] == ver:
        return hit[1]
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    Cout, Cin = w.shape[0], w.shape[1]
    taps = w.numel() // (Cout * Cin)
    with torch.inference_mode(False):
        out = hit[1] if (hit is not None and hit[1].shape == (Cout, taps, 6 * Cin)) else \
            torch.empty((Cout, taps, 6 * Cin), dtype=torch.bfloat16, device=w.device)
        w6 = torch.empty((Cout, 6 * Cin, taps), dtype=torch.float32, device=w.device)
    L.check(L.lib().rho_split3_weight(w.data_ptr(), w6.data_ptr(), Cout, Cin, taps, stream()), "split3_weight")
    L.check(L.lib().rho_conv_pack_weight(w6.data_ptr(), out.data_ptr(), BF16, Cout, 6 * Cin, taps, stream()), "conv_pack_weight")
    cache["split6"] = (ver, out)
    return out


def _conv_fp32_split(owner, x, x2, weight, bias, residual, spec):
    """fp32 conv_nd through the tcgen05 engine (see fp32_split_enabled); returns None when the shape has to stay on the CUDA cores."""
    Ct = x.shape[-1] + (x2.shape[-1] if x2 is not None else 0)
    Cout = weight.shape[0]
    if Ct % 32 or Cout % 32 or x.shape[-1] % 4 or (x2 is not None and x2.shape[-1] % 4):
        return None
    if max(spec.up) > 1:                               # the fused-upsample pack pre-sums taps (not bf16-exact): upsample explicitly
        x = upsample_nearest(x, spec.up)
        x2 = upsample_nearest(x2, spec.up) if x2 is not None else None
        spec = ConvSpec(spec.dims, spec.kernel, spec.stride, (1,) * spec.dims)
    rows = x.numel() // x.shape[-1]
    xc = torch.empty((*x.shape[:-1], 6 * Ct), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().rho_split3_concat(x.data_ptr(), x.shape[-1], ptr(x2), x2.shape[-1] if x2 is not None else 0, xc.data_ptr(), rows, stream()),
            "split3_concat")
    wp6 = split_packed_weight(owner, weight)
    b = bias.detach().float() if bias is not None else None
    p = _conv_params(xc, None, wp6, b, None, None, spec, Cout, L.CONV_TCGEN05, True)
    if int(L.lib().rho_conv_engine_for(C.byref(p))) != L.CONV_TCGEN05:
        return None
    y = conv_forward(xc, None, wp6, b, None, spec, Cout, L.CONV_TCGEN05, y_f32=True)
    return cast(y, torch.float32, add=residual) if residual is not None else y


def conv(owner, x, x2, weight, bias, residual, spec, dtype, engine=L.CONV_AUTO, y_f32=False):
    """conv_nd through the autograd tape; ``owner`` is the module that holds ``weight`` (packed-weight cache)."""
    if dtype == torch.float32 and engine == L.CONV_AUTO and not torch.is_grad_enabled() and fp32_split_enabled():
        y = _conv_fp32_split(owner, x, x2, weight, bias, residual, spec)
        if y is not None:
            return y
    wp = packed_weight(owner, weight, dtype)
    aux = aux_weight(owner, weight, spec, x, x2, wp, dtype) if engine != L.CONV_SIMT else None
    stats = [] if (dtype == torch.bfloat16 and engine != L.CONV_SIMT and not y_f32) else None
    y = ConvFn.apply(x, x2, weight, wp, bias, residual, spec, engine, y_f32, aux, stats)
    if stats:
        y._rho_part = stats[0]          # (per-tile channel sums, rows per sample): consumed by the next GroupNorm
    return y


def carry_stats(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """Propagates the fused-statistics side channel across a pure view (reshape) of a conv output."""
    part = getattr(src, "_rho_part", None)
    if part is not None and dst is not src:
        dst._rho_part = part
    return dst


def cast(x: torch.Tensor, dtype: torch.dtype, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dtype conversion (optionally fused with an elementwise add)."""
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    if add is not None:
        add = add.contiguous()
    L.check(L.lib().rho_add_cast(x.data_ptr(), dt(x), ptr(add), dt(add) if add is not None else 0, y.data_ptr(), dt(y), x.numel(),
                                 stream()), "add_cast")
    return y


class ForkFn(torch.autograd.Function):
    """Identity with ``n`` outputs for a tensor that feeds ``n`` consumers (ResBlock input -> GroupNorm + skip, the UNet's long
    skips, the time embedding -> every ResBlock).  Autograd would sum the consumers' gradients with ATen's ``add``; here the sum is
    one ``rho_add_cast`` launch per extra consumer, so the training step runs no kernel that is not ours."""

    @staticmethod
    def forward(ctx, x, n):
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        gs = [g for g in grads if g is not None]
        if not gs:
            return None, None
        acc = gs[0]
        for g in gs[1:]:
            acc = cast(acc, acc.dtype, add=g if g.dtype == acc.dtype else cast(g, acc.dtype))
        return acc, None


def fork(x: Optional[torch.Tensor], n: int = 2):
    """``n`` aliases of ``x`` whose gradients are summed by our own kernel (see ForkFn); plain aliases outside training."""
    if x is None:
        return (None,) * n
    if not (torch.is_grad_enabled() and x.requires_grad):
        return (x,) * n
    outs = ForkFn.apply(x, n)
    for o in outs:
        carry_stats(o, x)
    return outs


# ------------------------------------------------------------------------------------------------
# GroupNorm32 (+FiLM / embedding add) + activation
# ------------------------------------------------------------------------------------------------
def _gn_params(x, x2, gamma, beta, film, pre_add, act, stats, y, groups, eps) -> L.GnParams:
    p = L.GnParams()
    N = x.shape[0]
    S = 1
    for s in x.shape[1:-1]:
        S *= s
    C1, C2 = x.shape[-1], (x2.shape[-1] if x2 is not None else 0)
    Ct = C1 + C2
    p.N, p.S, p.C1, p.C2, p.groups = N, S, C1, C2, groups
    p.in_dtype, p.out_dtype, p.act, p.eps = dt(x), (dt(y) if y is not None else dt(x)), act, eps
    p.x, p.x2 = ptr(x), ptr(x2)
    if pre_add is not None:
        p.pre_add, p.pre_add_stride = pre_add.data_ptr(), pre_add.stride(0)
    p.gamma, p.beta = ptr(gamma), ptr(beta)
    if film is not None:  # [N, 2C]: scale = [:, :C], shift = [:, C:]
        p.scale = film.data_ptr()
        p.shift = film.data_ptr() + Ct * 4
        p.film_stride = film.stride(0)
    p.stats, p.y = ptr(stats), ptr(y)
    return p


def gn_forward(x, x2, gamma, beta, film, pre_add, act, out_dtype, groups=32, eps=1e-5):
    _need_cuda(x, x2, gamma, beta, film, pre_add)
    assert x.is_contiguous() and (x2 is None or x2.is_contiguous())
    for t in (film, pre_add):
        assert t is None or (t.dtype == torch.float32 and t.stride(-1) == 1)
    N = x.shape[0]
    Ct = x.shape[-1] + (x2.shape[-1] if x2 is not None else 0)
    stats = torch.empty((N, groups, 2), dtype=torch.float32, device=x.device)
    y = torch.empty((*x.shape[:-1], Ct), dtype=out_dtype, device=x.device)
    p = _gn_params(x, x2, gamma, beta, film, pre_add, act, stats, y, groups, eps)
    part1 = getattr(x, "_rho_part", None)
    part2 = getattr(x2, "_rho_part", None) if x2 is not None else None
    if pre_add is None and part1 is not None and (x2 is None or part2 is not None) and x.dtype == torch.bfloat16:
        # statistics come from the producing convs' epilogues: no pass over x
        if profiler is not None:
            ev = profiler.begin()
        L.check(L.lib().rho_gn_stats_from_partials(C.byref(p), part1[0].data_ptr(), part1[1], None if part2 is None else part2[0].data_ptr(),
                                                   0 if part2 is None else part2[1], stream()), "gn_stats_from_partials")
        if profiler is not None:
            profiler.end("gn_stats", "gn_stats_fused", 0.0, 0.0, ev)
            ev = profiler.begin()
        L.check(L.lib().rho_gn_apply(C.byref(p), stream()), "gn_apply")
        if profiler is not None:
            in_bytes = x.numel() * x.element_size() + (x2.numel() * x2.element_size() if x2 is not None else 0)
            profiler.end("gn_apply", "gn_apply", 0.0, in_bytes + y.numel() * y.element_size(), ev)
        return y, stats
    ws = workspace(L.lib().rho_gn_workspace_bytes(C.byref(p)), x.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    if profiler is not None:
        in_bytes = x.numel() * x.element_size() + (x2.numel() * x2.element_size() if x2 is not None else 0)
        ev = profiler.begin()
        L.check(L.lib().rho_gn_stats(C.byref(p), stream()), "gn_stats")
        profiler.end("gn_stats", "gn_stats", 0.0, in_bytes, ev)
        ev = profiler.begin()
        L.check(L.lib().rho_gn_apply(C.byref(p), stream()), "gn_apply")
        profiler.end("gn_apply", "gn_apply", 0.0, in_bytes + y.numel() * y.element_size(), ev)
        return y, stats
    L.check(L.lib().rho_gn_stats(C.byref(p), stream()), "gn_stats")
    L.check(L.lib().rho_gn_apply(C.byref(p), stream()), "gn_apply")
    return y, stats


class GnActFn(torch.autograd.Function):
    """act(FiLM(GroupNorm32(cat(x, x2) + pre_add))); see rho_gn_* in include/rho_b200.h."""

    @staticmethod
    def forward(ctx, x, x2, gamma, beta, film, pre_add, act, out_dtype, groups, eps):
        g32, b32 = gamma.detach().float(), beta.detach().float()
        y, stats = gn_forward(x, x2, g32, b32, film, pre_add, act, out_dtype, groups, eps)
        ctx.save_for_backward(x, x2, g32, b32, film, pre_add, stats)
        ctx.act, ctx.groups, ctx.eps = act, groups, eps
        ctx.params = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x2, gamma, beta, film, pre_add, stats = ctx.saved_tensors
        dy = dy.contiguous()
        Ct = gamma.numel()
        bp = L.GnBwdParams()
        bp.fwd = _gn_params(x, x2, gamma, beta, film, pre_add, ctx.act, stats, None, ctx.groups, ctx.eps)
        bp.fwd.out_dtype = dt(dy)
        dx = torch.empty_like(x)
        dx2 = torch.empty_like(x2) if x2 is not None else None
        sg, sbt = grad_sink(ctx.params[0], ctx.needs_input_grad[2]), grad_sink(ctx.params[1], ctx.needs_input_grad[3])
        dgamma = sg.view if sg else torch.empty_like(gamma)
        dbeta = sbt.view if sbt else torch.empty_like(beta)
        dfilm = torch.empty_like(film) if film is not None else None
        dpre = torch.empty((x.shape[0], Ct), dtype=torch.float32, device=x.device) if pre_add is not None else None
        bp.dy, bp.dx, bp.dx2, bp.dgamma, bp.dbeta = ptr(dy), ptr(dx), ptr(dx2), ptr(dgamma), ptr(dbeta)
        if dfilm is not None:
            bp.dscale = dfilm.data_ptr()
            bp.dshift = dfilm.data_ptr() + Ct * 4
            bp.dfilm_stride = dfilm.stride(0)
        bp.dpre_add = ptr(dpre)
        ws = workspace(L.lib().rho_gn_bwd_workspace_bytes(C.byref(bp)), x.device)
        bp.workspace, bp.workspace_bytes = ws.data_ptr(), ws.numel()
        L.check(L.lib().rho_gn_bwd(C.byref(bp), stream()), "gn_bwd")
        if sg:
            sg.done(); dgamma = None
        if sbt:
            sbt.done(); dbeta = None
        return dx, dx2, dgamma, dbeta, dfilm, dpre, None, None, None, None


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
class AttentionFn(torch.autograd.Function):
    """QKVAttention[Legacy] on channels-last qkv [N, T, 3*heads*ch]; see rho_attention_fwd."""

    @staticmethod
    def forward(ctx, qkv, heads, order):
        _need_cuda(qkv)
        qkv = qkv.contiguous()
        N, T, W3 = qkv.shape
        ch = W3 // (3 * heads)
        out = torch.empty((N, T, heads * ch), dtype=qkv.dtype, device=qkv.device)
        need_lse = ctx.needs_input_grad[0]
        lse = torch.empty((N, heads, T), dtype=torch.float32, device=qkv.device) if need_lse else None
        p = L.AttnParams()
        p.N, p.T, p.heads, p.ch, p.order, p.dtype = N, T, heads, ch, order, dt(qkv)
        p.qkv, p.out, p.lse = ptr(qkv), ptr(out), ptr(lse)
        ev = profiler.begin() if profiler is not None else None
        L.check(L.lib().rho_attention_fwd(C.byref(p), stream()), "attention_fwd")
        if ev is not None:
            profiler.end("attention", f"attn_T{T}_ch{ch}", 4.0 * N * heads * T * T * ch,
                         (qkv.numel() + out.numel()) * qkv.element_size(), ev)
        ctx.save_for_backward(qkv, out, lse)
        ctx.dims = (N, T, heads, ch, order)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        N, T, heads, ch, order = ctx.dims
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        bp = L.AttnBwdParams()
        f = bp.fwd
        f.N, f.T, f.heads, f.ch, f.order, f.dtype = N, T, heads, ch, order, dt(qkv)
        f.qkv, f.out, f.lse = ptr(qkv), ptr(out), ptr(lse)
        bp.dout, bp.dqkv = ptr(dout), ptr(dqkv)
        ws = workspace(L.lib().rho_attention_bwd_workspace_bytes(C.byref(bp)), qkv.device)
        bp.workspace, bp.workspace_bytes = ws.data_ptr(), ws.numel()
        L.check(L.lib().rho_attention_bwd(C.byref(bp), stream()), "attention_bwd")
        return dqkv, None, None


def attention(qkv, heads, order):
    return AttentionFn.apply(qkv, heads, order)


# ------------------------------------------------------------------------------------------------
# embedding path (fp32)
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int, omega: torch.Tensor) -> torch.Tensor:
    _need_cuda(t, omega)
    t = t.to(torch.int64).contiguous()
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    L.check(L.lib().rho_timestep_embedding(t.data_ptr(), t.shape[0], dim, omega.data_ptr(), out.data_ptr(), stream()),
            "timestep_embedding")
    return out


def _linear_params(x, w, b, add, y, act_in) -> L.LinearParams:
    p = L.LinearParams()
    p.B, p.I, p.O, p.act_in = x.shape[0], x.shape[1], w.shape[0], act_in
    p.x, p.x_stride, p.w, p.b = x.data_ptr(), x.stride(0), w.data_ptr(), ptr(b)
    if add is not None:
        p.add, p.add_stride = add.data_ptr(), add.stride(0)
    if y is not None:
        p.y, p.y_stride = y.data_ptr(), y.stride(0)
    return p


class LinearFn(torch.autograd.Function):
    """y = W act_in(x) + b (+ add); nn.Linear of time_embed / emb_layers with the activation fused."""

    @staticmethod
    def forward(ctx, x, weight, bias, add, act_in):
        _need_cuda(x, weight)
        x = x.float().contiguous()
        w = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        a = add.float().contiguous() if add is not None else None
        y = torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
        p = _linear_params(x, w, b, a, y, act_in)
        L.check(L.lib().rho_linear_fwd(C.byref(p), stream()), "linear_fwd")
        ctx.save_for_backward(x, w)
        ctx.act_in, ctx.has_bias, ctx.has_add = act_in, bias is not None, add is not None
        ctx.params = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.float().contiguous()
        p = _linear_params(x, w, None, None, None, ctx.act_in)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        need_db = ctx.has_bias and ctx.needs_input_grad[2]
        sw, sb = grad_sink(ctx.params[0], ctx.needs_input_grad[1]), grad_sink(ctx.params[1], need_db)
        dw = (sw.view if sw else torch.empty_like(w)) if ctx.needs_input_grad[1] else None
        db = (sb.view if sb else torch.empty((w.shape[0],), dtype=torch.float32, device=x.device)) if need_db else None
        L.check(L.lib().rho_linear_bwd(C.byref(p), dy.data_ptr(), dy.stride(0), ptr(dx), dx.stride(0) if dx is not None else 0,
                                       ptr(dw), ptr(db), stream()), "linear_bwd")
        if sw:
            sw.done(); dw = None
        if sb:
            sb.done(); db = None
        dadd = dy if (ctx.has_add and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dadd, None


def linear(x, weight, bias=None, add=None, act_in=L.ACT_NONE):
    return LinearFn.apply(x, weight, bias, add, act_in)


class EmbeddingMatchFn(torch.autograd.Function):
    """sum_i table_i[index of y[:, i] in values_i] (MultiEmbeddings.forward, models/conditioning.py:115-139).  Rows with a value
    that matches no point of the space come out as NaN; ``last_indices`` ([n_keys, B] int32, -1 = no match) lets the caller check."""
    last_indices = None

    @staticmethod
    def forward(ctx, y, values_list, *tables):
        _need_cuda(y, *tables)
        y = y.float()
        if y.dim() == 1:
            y = y[:, None].expand(-1, len(tables))
        y = y.contiguous()
        B, dim = y.shape[0], tables[0].shape[1]
        out = torch.empty((B, dim), dtype=torch.float32, device=y.device)
        idxs = torch.empty((len(tables), B), dtype=torch.int32, device=y.device)
        for i, (vals, tab) in enumerate(zip(values_list, tables)):
            tab32 = tab.detach().float().contiguous()
            L.check(L.lib().rho_embedding_match(y.data_ptr() + 4 * i, y.stride(0), B, vals.data_ptr(), vals.numel(), tab32.data_ptr(), dim,
                                                out.data_ptr(), int(i > 0), idxs[i].data_ptr(), stream()), "embedding_match")
        ctx.save_for_backward(idxs)
        EmbeddingMatchFn.last_indices = idxs
        ctx.shapes = [tuple(t.shape) for t in tables]
        ctx.params = tables
        return out

    @staticmethod
    def backward(ctx, dout):
        (idxs,) = ctx.saved_tensors
        dout = dout.float().contiguous()
        grads = []
        for i, shp in enumerate(ctx.shapes):
            if not ctx.needs_input_grad[2 + i]:
                grads.append(None)
                continue
            sink = grad_sink(ctx.params[i])
            g = sink.view.zero_() if sink else torch.zeros(shp, dtype=torch.float32, device=dout.device)
            L.check(L.lib().rho_embedding_match_bwd(idxs[i].data_ptr(), dout.shape[0], dout.data_ptr(), shp[1], g.data_ptr(), shp[0],
                                                    stream()), "embedding_match_bwd")
            if sink:
                sink.done()
            grads.append(None if sink else g)
        return (None, None, *grads)


# ------------------------------------------------------------------------------------------------
# diffusion math
# ------------------------------------------------------------------------------------------------
def randn(shape, device, seed: int, offset: int = 0) -> torch.Tensor:
    out = torch.empty(shape, dtype=torch.float32, device=device)
    L.check(L.lib().rho_randn(out.data_ptr(), out.numel(), seed, offset, stream()), "randn")
    return out


def q_sample(x0, noise, t, alpha_bar):
    _need_cuda(x0, noise, t, alpha_bar)
    x0 = x0.float().contiguous()
    noise = noise.float().contiguous()
    t = t.to(torch.int64).contiguous()
    out = torch.empty_like(x0)
    N = x0.shape[0]
    L.check(L.lib().rho_q_sample(x0.data_ptr(), noise.data_ptr(), t.data_ptr(), alpha_bar.data_ptr(), alpha_bar.numel(), N,
                                 x0.numel() // N, out.data_ptr(), stream()), "q_sample")
    return out


def p_sample_step(x, eps, z, t_scalar, alpha, beta, alpha_bar, noise_scale=0.8, clamp=(-1.0, 1.0), seed=0, seed_dev=None):
    """In-place ancestral update of ``x`` (fp32); ``t_scalar`` is a 1-element int64 device tensor.  ``seed_dev``: optional
    1-element int64 device tensor that replaces ``seed`` (a captured graph then serves every run)."""
    p = L.PSampleParams()
    p.seed_dev = ptr(seed_dev)
    p.x, p.eps, p.z, p.t = x.data_ptr(), eps.data_ptr(), ptr(z), t_scalar.data_ptr()
    p.alpha, p.beta, p.alpha_bar, p.T = alpha.data_ptr(), beta.data_ptr(), alpha_bar.data_ptr(), alpha_bar.numel()
    p.n, p.noise_scale, p.clamp_lo, p.clamp_hi, p.seed = x.numel(), noise_scale, clamp[0], clamp[1], seed
    L.check(L.lib().rho_p_sample_step(C.byref(p), stream()), "p_sample_step")


def advance_timestep(t_scalar, t_vec):
    L.check(L.lib().rho_advance_timestep(t_scalar.data_ptr(), t_vec.data_ptr(), t_vec.numel(), stream()), "advance_timestep")


class MseLossFn(torch.autograd.Function):
    """nn.MSELoss(pred, target) (ddpm.py:280); the gradient is produced by the same kernel."""

    @staticmethod
    def forward(ctx, pred, target):
        _need_cuda(pred, target)
        pred = pred.contiguous()
        target = target.float().contiguous()
        n = pred.numel()
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        dpred = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        ws = workspace(L.lib().rho_mse_workspace_bytes(n), pred.device)
        L.check(L.lib().rho_mse_loss(pred.data_ptr(), dt(pred), target.data_ptr(), n, loss.data_ptr(), ptr(dpred), 1.0, ws.data_ptr(),
                                     ws.numel(), stream()), "mse_loss")
        ctx.dpred = dpred
        return loss

    @staticmethod
    def backward(ctx, dloss):
        g = ctx.dpred
        ctx.dpred = None
        if g is None:
            return None, None
        # upstream scale (1.0 for loss.backward()) is applied on the device: no host sync
        dloss = dloss.float().contiguous()
        L.check(L.lib().rho_scale(g.data_ptr(), dt(g), g.numel(), dloss.data_ptr(), 1.0, stream()), "scale")
        return g, None


def adamw_ema_step(p, g, m, v, shadow, step, lr, betas, eps, weight_decay, ema_decay_t=0.0, grad_scale=1.0, step_dev=None, ema_decay=0.0):
    """``step_dev``: optional 16-element int32 device tensor holding the step count (graph-replayable form, see rho_b200.h)."""
    a = L.AdamWParams()
    a.p, a.g, a.m, a.v, a.shadow, a.n = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), ptr(shadow), p.numel()
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = lr, betas[0], betas[1], eps, weight_decay
    a.step, a.ema_decay_t, a.grad_scale = step, ema_decay_t, grad_scale
    a.step_dev, a.ema_decay = ptr(step_dev), ema_decay
    L.check(L.lib().rho_adamw_ema_step(C.byref(a), stream()), "adamw_ema_step")
    bump_weights_epoch()


def ema_update(shadow: torch.Tensor, p: torch.Tensor, one_minus_decay: float) -> None:
    """shadow -= one_minus_decay * (shadow - p), in place (ema.py:52-57)."""
    _need_cuda(shadow, p)
    assert shadow.is_contiguous() and p.is_contiguous() and shadow.dtype == torch.float32 and p.dtype == torch.float32
    L.check(L.lib().rho_ema_update(shadow.data_ptr(), p.data_ptr(), shadow.numel(), one_minus_decay, stream()), "ema_update")


# ------------------------------------------------------------------------------------------------
# GaussianDiffusionPipeline step kernels (gaussian_diffusion.py)
# ------------------------------------------------------------------------------------------------
def lincomb2(a, b, t, ca, cb):
    """out[n] = ca[t_n] * a[n] + cb[t_n] * b[n]  (q_sample / posterior mean with fp32 coefficient tables)."""
    _need_cuda(a, b, t, ca, cb)
    a = a.float().contiguous()
    b = b.float().contiguous()
    t = t.to(torch.int64).contiguous()
    out = torch.empty_like(a)
    N = a.shape[0]
    L.check(L.lib().rho_lincomb2(a.data_ptr(), b.data_ptr(), t.data_ptr(), ca.data_ptr(), cb.data_ptr(), ca.numel(), N, a.numel() // N,
                                 out.data_ptr(), stream()), "lincomb2")
    return out


def quantile_ranks(per_sample: int, q: float):
    """torch.quantile's interpolation position, evaluated in float32 like ATen does (rank = q * (n - 1))."""
    pos = torch.tensor(q, dtype=torch.float32) * (per_sample - 1)
    lo = torch.floor(pos)
    k_lo = int(lo)
    k_hi = int(torch.ceil(pos))
    return k_lo, min(k_hi, per_sample - 1), float(pos - lo)


def abs_quantile(x, q: float, floor_value: float = float("-inf"), out=None):
    """out[n] = max(floor_value, quantile_q(|x[n]|)) with torch.quantile's linear interpolation (process_xstart)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    N = x.shape[0]
    per = x.numel() // N
    k_lo, k_hi, w = quantile_ranks(per, q)
    if out is None:
        out = torch.empty((N,), dtype=torch.float32, device=x.device)
    ws = workspace(L.lib().rho_abs_quantile_workspace_bytes(N), x.device)
    L.check(L.lib().rho_abs_quantile(x.data_ptr(), N, per, k_lo, k_hi, w, floor_value, out.data_ptr(), ws.data_ptr(), ws.numel(), stream()),
            "abs_quantile")
    return out


def x0_step(model_out, x, tables: dict, mode: int, t=None, t_scalar=None, thresh=None, noise=None, eta: float = 0.0, pred_xstart=None):
    """In-place x_t -> x_{t-1} for an x0-predicting model: mode 0 = ddim_sample, mode 1 = p_sample (see rho_b200.h)."""
    _need_cuda(model_out, x)
    assert x.dtype == torch.float32 and x.is_contiguous() and model_out.is_contiguous()
    p = L.X0StepParams()
    p.model_out, p.model_dtype, p.mode = model_out.data_ptr(), dt(model_out), mode
    p.x, p.noise, p.pred_xstart, p.thresh = x.data_ptr(), ptr(noise), ptr(pred_xstart), ptr(thresh)
    p.t, p.t_scalar = ptr(t), ptr(t_scalar)
    p.T, p.N, p.per_sample, p.eta = tables["alphas_cumprod"].numel(), x.shape[0], x.numel() // x.shape[0], eta
    for k in ("alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
              "posterior_mean_coef1", "posterior_mean_coef2", "model_log_variance"):
        setattr(p, k, tables[k].data_ptr())
    L.check(L.lib().rho_x0_step(C.byref(p), stream()), "x0_step")
    return x


# ------------------------------------------------------------------------------------------------
# optional resampling / dropout paths (conv_resample=False, resblock_updown=True, dropout > 0)
# ------------------------------------------------------------------------------------------------
def _resample(x, factors, mode, out_spatial):
    _need_cuda(x)
    assert x.is_contiguous()
    sp = _pad3(tuple(x.shape[1:-1]))
    f = _pad3(tuple(factors))
    osp = _pad3(tuple(out_spatial))
    y = torch.empty((x.shape[0], *out_spatial, x.shape[-1]), dtype=x.dtype, device=x.device)
    L.check(L.lib().rho_resample(x.data_ptr(), y.data_ptr(), dt(x), x.shape[0], sp[0], sp[1], sp[2], x.shape[-1], f[0], f[1], f[2], mode,
                                 osp[0], osp[1], osp[2], stream()), "resample")
    return y


class ResampleFn(torch.autograd.Function):
    """kind "avg": avg_pool_nd(kernel = stride = factors); kind "nearest": nearest upsample by ``factors`` (channels-last)."""

    @staticmethod
    def forward(ctx, x, kind, factors):
        factors = tuple(int(v) for v in factors)
        spatial = tuple(x.shape[1:-1])
        assert len(factors) == len(spatial)
        if kind == "avg":
            out_sp = tuple(s // f for s, f in zip(spatial, factors))
            y = _resample(x, factors, 0, out_sp)
        else:
            out_sp = tuple(s * f for s, f in zip(spatial, factors))
            y = _resample(x, factors, 1, out_sp)
        ctx.kind, ctx.factors, ctx.spatial = kind, factors, spatial
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = _resample(dy, ctx.factors, 3 if ctx.kind == "avg" else 2, ctx.spatial)
        return dx, None, None


def avg_pool(x, factors):
    return ResampleFn.apply(x, "avg", factors)


def upsample_nearest(x, factors):
    return ResampleFn.apply(x, "nearest", factors)


_dropout_calls = 0
dropout_step_dev: Optional[torch.Tensor] = None      # device step counter (FusedAdamW.enable_device_step) folded into the mask offset


def weights_signature(module) -> tuple:
    """Changes whenever parameter memory of ``module`` may have been rewritten: optimizer steps of the fused optimizer (epoch),
    in-place updates under autograd / ``load_state_dict`` / EMA swaps (version counters), re-pointed storage (addresses)."""
    ver = ptrs = 0
    for q in module.parameters():
        ver += q._version
        ptrs ^= q.data_ptr()
    return (_weights_epoch, ver, ptrs)


class DropoutFn(torch.autograd.Function):
    """nn.Dropout(p) in training mode; the mask is a pure function of (seed, call index within the step, optimizer step,
    element index).  The call index restarts with every optimizer step (``begin_step``), the step count is read ON THE DEVICE
    when the fused optimizer keeps it there, so a training step replayed from a CUDA graph draws a fresh mask every step."""

    @staticmethod
    def forward(ctx, x, p, seed):
        global _dropout_calls
        _dropout_calls += 1
        x = x.contiguous()
        y = torch.empty_like(x)
        ctx.p, ctx.seed, ctx.offset, ctx.step_dev = float(p), int(seed), _dropout_calls, dropout_step_dev
        L.check(L.lib().rho_dropout(x.data_ptr(), y.data_ptr(), dt(x), x.numel(), ctx.p, ctx.seed, ctx.offset, ptr(ctx.step_dev), stream()),
                "dropout")
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        L.check(L.lib().rho_dropout(dy.data_ptr(), dx.data_ptr(), dt(dy), dy.numel(), ctx.p, ctx.seed, ctx.offset, ptr(ctx.step_dev), stream()),
                "dropout_bwd")
        return dx, None, None


def dropout(x, p: float, seed: int = 0):
    return DropoutFn.apply(x, p, seed)


# ------------------------------------------------------------------------------------------------
# synthetic density fields generated on the device (rho_diffusion/data/synthetic.py)
# ------------------------------------------------------------------------------------------------
def sph_harm_density(l: torch.Tensor, m: torch.Tensor, grid_el: int, lo: float = -2.0, hi: float = 2.0, minmax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,) int32 device tensors of degrees / orders -> (B, 1, G, G, G) fp32 normalised |Y_l^|m| r| volumes (synthetic.py:84-124).
    ``minmax``: optional (B, 4) float64 device tensor with (re, im) of the complex minimum and maximum to normalise with."""
    _need_cuda(l, m, minmax)
    if minmax is not None:
        assert minmax.dtype == torch.float64 and minmax.shape == (l.numel(), 4) and minmax.is_contiguous()
    l, m = l.to(torch.int32).contiguous(), m.to(torch.int32).contiguous()
    B = l.numel()
    out = torch.empty((B, 1, grid_el, grid_el, grid_el), dtype=torch.float32, device=l.device)
    ws = workspace(L.lib().rho_sph_harm_workspace_bytes(B, grid_el), l.device)
    L.check(L.lib().rho_sph_harm_density(out.data_ptr(), B, grid_el, l.data_ptr(), m.data_ptr(), float(lo), float(hi), ptr(minmax),
                                         ws.data_ptr(), ws.numel(), stream()), "sph_harm_density")
    return out


def gaussian_blobs(n_blobs: torch.Tensor, width: torch.Tensor, data_shape, seed: int, first_index: int) -> torch.Tensor:
    """(B,) int32 blob counts and (B,) fp32 widths -> (B, 1, *data_shape) fp32 fields in [-1, 1] (see rho_gaussian_blobs)."""
    _need_cuda(n_blobs, width)
    n_blobs, width = n_blobs.to(torch.int32).contiguous(), width.to(torch.float32).contiguous()
    B = n_blobs.numel()
    D, H, W = _pad3(data_shape)
    out = torch.empty((B, 1, *[int(s) for s in data_shape]), dtype=torch.float32, device=n_blobs.device)
    L.check(L.lib().rho_gaussian_blobs(out.data_ptr(), B, D, H, W, n_blobs.data_ptr(), width.data_ptr(), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                       int(first_index), stream()), "gaussian_blobs")
    return out


# ------------------------------------------------------------------------------------------------
# the remaining registered backbones: v1 UNet (models/unet.py) and VisionTransformer (models/vit.py)
# ------------------------------------------------------------------------------------------------
class ActivationFn(torch.autograd.Function):
    """Stand-alone registry activation (unet.py:122-123 ``activation(conv(x))``; vit.py:155-161 Linear -> activation)."""

    @staticmethod
    def forward(ctx, x, act):
        _need_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        L.check(L.lib().rho_activation(x.data_ptr(), y.data_ptr(), dt(x), x.numel(), act, stream()), "activation")
        ctx.save_for_backward(x)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = cast(dy, x.dtype)
        dx = torch.empty_like(x)
        L.check(L.lib().rho_activation_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), dt(x), x.numel(), ctx.act, stream()), "activation_bwd")
        return dx, None


def activation(x, act: int):
    return x if act == L.ACT_NONE else ActivationFn.apply(x, act)


class ConvTWeightFn(torch.autograd.Function):
    """[A, B, *k] fp32 -> [B, A, *k] flipped along every kernel axis: the conv_nd weight of a stride-1 ConvTranspose
    (unet.py:69-75); with k = () a plain transpose.  Backward is the same map with A and B swapped."""

    @staticmethod
    def forward(ctx, w):
        _need_cuda(w)
        w = w.detach().float().contiguous()
        A, B = w.shape[0], w.shape[1]
        taps = w.numel() // (A * B)
        out = torch.empty((B, A, *w.shape[2:]), dtype=torch.float32, device=w.device)
        L.check(L.lib().rho_convt_weight(w.data_ptr(), out.data_ptr(), A, B, taps, stream()), "convt_weight")
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.float().contiguous()
        B, A = g.shape[0], g.shape[1]
        taps = g.numel() // (A * B)
        out = torch.empty((A, B, *g.shape[2:]), dtype=torch.float32, device=g.device)
        L.check(L.lib().rho_convt_weight(g.data_ptr(), out.data_ptr(), B, A, taps, stream()), "convt_weight_bwd")
        return out


class _NoCache:
    """Throw-away owner for weights that are recomputed every call (no packed-weight cache)."""


def conv_uncached(x, x2, weight, bias, residual, spec, dtype, engine=L.CONV_AUTO, y_f32=False):
    """:func:`conv` for a weight that is itself the output of a kernel (transposed / flipped ConvTranspose weights): the packed copy
    is rebuilt every call, because a recycled device address would alias a stale cache entry."""
    return conv(_NoCache(), x, x2, weight, bias, residual, spec, dtype, engine, y_f32)


def linear_tokens(owner, x, weight, bias=None, residual=None, dtype=None, engine=L.CONV_AUTO, y_f32=False, cached=True):
    """nn.Linear over channels-last tokens [..., I] -> [..., O] on the conv_nd engines (a 1x1 conv; tcgen05 in bf16), bias and
    an optional residual fused in the epilogue.  ``weight`` is the [O, I] Linear weight."""
    spec = ConvSpec.make(1, 1)
    lead = x.shape[:-1]
    x3 = x.reshape(x.shape[0], -1, x.shape[-1])
    r3 = residual.reshape(residual.shape[0], -1, residual.shape[-1]) if residual is not None else None
    dtype = dtype or x.dtype
    y = conv(owner if cached else _NoCache(), x3, None, weight, bias, r3, spec, dtype, engine, y_f32)
    return y.reshape(*lead, weight.shape[0])


class PatchifyFn(torch.autograd.Function):
    """(N, C, *spatial) field <-> (N, T, C prod(patch)) tokens (rho_patchify); the inverse direction adds a per-channel bias."""

    @staticmethod
    def forward(ctx, src, patch, inverse, field_shape, out_dtype, bias):
        _need_cuda(src, bias)
        src = src.contiguous()
        N, Cc, *sp = field_shape
        D, H, W = _pad3(sp)
        pd, ph, pw = _pad3(patch)
        T = (D // pd) * (H // ph) * (W // pw)
        F = Cc * pd * ph * pw
        b = bias.detach().float().contiguous() if bias is not None else None
        if inverse:
            assert tuple(src.shape) == (N, T, F), (tuple(src.shape), (N, T, F))
            dst = torch.empty(tuple(field_shape), dtype=out_dtype, device=src.device)
        else:
            assert tuple(src.shape) == tuple(field_shape)
            dst = torch.empty((N, T, F), dtype=out_dtype, device=src.device)
        L.check(L.lib().rho_patchify(src.data_ptr(), dt(src), dst.data_ptr(), dt(dst), ptr(b), N, Cc, D, H, W, pd, ph, pw, int(inverse),
                                     stream()), "patchify")
        ctx.args = (tuple(patch), bool(inverse), tuple(field_shape), src.dtype, bias is not None)
        return dst

    @staticmethod
    def backward(ctx, g):
        patch, inverse, field_shape, in_dtype, has_bias = ctx.args
        g = g.contiguous()
        dsrc = PatchifyFn.apply(g, patch, not inverse, field_shape, in_dtype, None) if ctx.needs_input_grad[0] else None
        dbias = None
        if has_bias and ctx.needs_input_grad[5]:
            N, Cc, *sp = field_shape
            S = 1
            for s in sp:
                S *= s
            # dbias[c] = sum over samples and positions of g[n, c, :]: row sums of [N C, S], then the sum over n
            rows = torch.empty((N * Cc,), dtype=torch.float32, device=g.device)
            L.check(L.lib().rho_bias_expand(cast(g, torch.float32).data_ptr(), rows.data_ptr(), N * Cc, S, 1, stream()), "bias_rowsum")
            dbias = torch.empty((Cc,), dtype=torch.float32, device=g.device)
            L.check(L.lib().rho_colsum(rows.data_ptr(), F32, Cc, 1, N, Cc, dbias.data_ptr(), stream()), "colsum")
        return dsrc, None, None, None, None, dbias


def patchify(x, patch, out_dtype):
    return PatchifyFn.apply(x, tuple(patch), False, tuple(x.shape), out_dtype, None)


def unpatchify(tokens, patch, field_shape, out_dtype, bias=None):
    return PatchifyFn.apply(tokens, tuple(patch), True, tuple(field_shape), out_dtype, bias)


class AddBcastFn(torch.autograd.Function):
    """x [N, ...] + add [...] broadcast over the leading dim (positional embedding, vit.py:349-352)."""

    @staticmethod
    def forward(ctx, x, add):
        _need_cuda(x, add)
        x = x.contiguous()
        a = add.float().contiguous()
        inner = a.numel()
        assert x.numel() % inner == 0 and tuple(x.shape[1:]) == tuple(a.shape)
        y = torch.empty_like(x)
        L.check(L.lib().rho_add_bcast(x.data_ptr(), dt(x), a.data_ptr(), y.data_ptr(), x.numel() // inner, inner, stream()), "add_bcast")
        ctx.add_shape = tuple(a.shape)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dadd = None
        if ctx.needs_input_grad[1]:
            inner = 1
            for s in ctx.add_shape:
                inner *= s
            dadd = torch.empty(ctx.add_shape, dtype=torch.float32, device=g.device)
            L.check(L.lib().rho_colsum(g.data_ptr(), dt(g), inner, 1, g.numel() // inner, inner, dadd.data_ptr(), stream()), "colsum")
        return g, dadd


def add_bcast(x, add):
    return AddBcastFn.apply(x, add)


def _ln_params(x, pre_add, gamma, beta, y, stats, eps) -> L.LayerNormParams:
    p = L.LayerNormParams()
    Cc = x.shape[-1]
    p.rows, p.C = x.numel() // Cc, Cc
    p.rows_per_sample = (x.numel() // Cc) // x.shape[0]
    p.in_dtype, p.out_dtype, p.eps = dt(x), (dt(y) if y is not None else dt(x)), eps
    p.x, p.pre_add, p.gamma, p.beta, p.y, p.stats = ptr(x), ptr(pre_add), ptr(gamma), ptr(beta), ptr(y), ptr(stats)
    return p


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(C)(x + pre_add[:, None, :]) over channels-last tokens [N, T, C] (vit.py:179-186)."""

    @staticmethod
    def forward(ctx, x, pre_add, gamma, beta, eps):
        _need_cuda(x, pre_add, gamma, beta)
        x = x.contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        pa = pre_add.float().contiguous() if pre_add is not None else None
        y = torch.empty_like(x)
        stats = torch.empty((x.numel() // x.shape[-1], 2), dtype=torch.float32, device=x.device)
        L.check(L.lib().rho_layernorm_fwd(C.byref(_ln_params(x, pa, g32, b32, y, stats, eps)), stream()), "layernorm_fwd")
        ctx.save_for_backward(x, pa, g32, stats)
        ctx.eps = eps
        ctx.params = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pa, g32, stats = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = cast(dy, x.dtype)
        p = _ln_params(x, pa, g32, None, None, stats, ctx.eps)
        dx = torch.empty_like(x)
        need_p = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        sg, sb = grad_sink(ctx.params[0], need_p), grad_sink(ctx.params[1], need_p)
        dgamma = (sg.view if sg else torch.empty_like(g32)) if need_p else None
        dbeta = (sb.view if sb else torch.empty_like(g32)) if need_p else None
        dpre = torch.empty_like(pa) if (pa is not None and ctx.needs_input_grad[1]) else None
        ws = workspace(L.lib().rho_layernorm_bwd_workspace_bytes(C.byref(p)), x.device)
        L.check(L.lib().rho_layernorm_bwd(C.byref(p), dy.data_ptr(), dx.data_ptr(), ptr(dgamma), ptr(dbeta), ptr(dpre), ws.data_ptr(), ws.numel(),
                                          stream()), "layernorm_bwd")
        if sg:
            sg.done(); dgamma = None
        if sb:
            sb.done(); dbeta = None
        return dx, dpre, dgamma, dbeta, None


def layernorm(x, gamma, beta, eps=1e-5, pre_add=None):
    return LayerNormFn.apply(x, pre_add, gamma, beta, eps)


class BiasExpandFn(torch.autograd.Function):
    """bias [C] -> [C reps] (each entry repeated): the bias of a kernel = stride ConvTranspose seen by its Linear (vit.py:297-303)."""

    @staticmethod
    def forward(ctx, b, reps):
        _need_cuda(b)
        b = b.detach().float().contiguous()
        out = torch.empty((b.numel() * reps,), dtype=torch.float32, device=b.device)
        L.check(L.lib().rho_bias_expand(b.data_ptr(), out.data_ptr(), b.numel(), reps, 0, stream()), "bias_expand")
        ctx.reps = reps
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.float().contiguous()
        out = torch.empty((g.numel() // ctx.reps,), dtype=torch.float32, device=g.device)
        L.check(L.lib().rho_bias_expand(g.data_ptr(), out.data_ptr(), out.numel(), ctx.reps, 1, stream()), "bias_expand_bwd")
        return out, None
