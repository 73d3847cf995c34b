"""ctypes binding of the C ABI declared in ``include/rho_b200.h`` (librho_b200.so).

There is no CPU fallback: if the shared library is missing the first call raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librho_b200.so")

RHO_OK = 0
F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_GELU, ACT_SYMLOG = range(7)
CONV_AUTO, CONV_SIMT, CONV_TCGEN05 = 0, 1, 2

c_i32, c_i64, c_u64, c_f32, c_vp, c_sz = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_void_p, C.c_size_t


class ConvParams(C.Structure):
    _fields_ = [
        ("N", c_i32), ("D", c_i32), ("H", c_i32), ("W", c_i32),
        ("C1", c_i32), ("C2", c_i32), ("Cout", c_i32),
        ("kd", c_i32), ("kh", c_i32), ("kw", c_i32),
        ("sd", c_i32), ("sh", c_i32), ("sw", c_i32),
        ("ud", c_i32), ("uh", c_i32), ("uw", c_i32),
        ("act_dtype", c_i32), ("w_dtype", c_i32), ("engine", c_i32), ("y_f32", c_i32),
        ("x", c_vp), ("x2", c_vp), ("w_packed", c_vp), ("w_aux", c_vp), ("bias", c_vp), ("residual", c_vp), ("y", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_sz), ("stats_partial", c_vp),
    ]


class ConvBwdParams(C.Structure):
    _fields_ = [
        ("fwd", ConvParams), ("dy", c_vp), ("dx", c_vp), ("dx2", c_vp), ("dw", c_vp), ("dbias", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_sz),
    ]


class GnParams(C.Structure):
    _fields_ = [
        ("N", c_i32), ("S", c_i64), ("C1", c_i32), ("C2", c_i32), ("groups", c_i32),
        ("in_dtype", c_i32), ("out_dtype", c_i32), ("act", c_i32), ("eps", c_f32),
        ("x", c_vp), ("x2", c_vp), ("pre_add", c_vp), ("pre_add_stride", c_i32),
        ("gamma", c_vp), ("beta", c_vp), ("scale", c_vp), ("shift", c_vp), ("film_stride", c_i32),
        ("stats", c_vp), ("y", c_vp), ("workspace", c_vp), ("workspace_bytes", c_sz),
    ]


class GnBwdParams(C.Structure):
    _fields_ = [
        ("fwd", GnParams), ("dy", c_vp), ("dx", c_vp), ("dx2", c_vp), ("dgamma", c_vp), ("dbeta", c_vp),
        ("dscale", c_vp), ("dshift", c_vp), ("dfilm_stride", c_i32), ("dpre_add", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_sz),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("N", c_i32), ("T", c_i32), ("heads", c_i32), ("ch", c_i32), ("order", c_i32), ("dtype", c_i32),
        ("qkv", c_vp), ("out", c_vp), ("lse", c_vp),
    ]


class AttnBwdParams(C.Structure):
    _fields_ = [("fwd", AttnParams), ("dout", c_vp), ("dqkv", c_vp), ("workspace", c_vp), ("workspace_bytes", c_sz)]


class LinearParams(C.Structure):
    _fields_ = [
        ("B", c_i32), ("I", c_i32), ("O", c_i32), ("act_in", c_i32),
        ("x", c_vp), ("x_stride", c_i32), ("w", c_vp), ("b", c_vp), ("add", c_vp), ("add_stride", c_i32),
        ("y", c_vp), ("y_stride", c_i32),
    ]


class PSampleParams(C.Structure):
    _fields_ = [
        ("x", c_vp), ("eps", c_vp), ("z", c_vp), ("t", c_vp), ("alpha", c_vp), ("beta", c_vp), ("alpha_bar", c_vp),
        ("T", c_i32), ("n", c_i64), ("noise_scale", c_f32), ("clamp_lo", c_f32), ("clamp_hi", c_f32), ("seed", c_u64),
        ("seed_dev", c_vp),
    ]


class AdamWParams(C.Structure):
    _fields_ = [
        ("p", c_vp), ("g", c_vp), ("m", c_vp), ("v", c_vp), ("shadow", c_vp), ("n", c_i64),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double),
        ("step", c_i32), ("ema_decay_t", c_f32), ("grad_scale", c_f32), ("step_dev", c_vp), ("ema_decay", c_f32),
    ]


class X0StepParams(C.Structure):
    _fields_ = [
        ("model_out", c_vp), ("model_dtype", c_i32), ("mode", c_i32), ("x", c_vp), ("noise", c_vp), ("pred_xstart", c_vp),
        ("thresh", c_vp), ("t", c_vp), ("t_scalar", c_vp), ("T", c_i32), ("N", c_i32), ("per_sample", c_i64), ("eta", c_f32),
        ("alphas_cumprod", c_vp), ("alphas_cumprod_prev", c_vp), ("sqrt_recip_alphas_cumprod", c_vp),
        ("sqrt_recipm1_alphas_cumprod", c_vp), ("posterior_mean_coef1", c_vp), ("posterior_mean_coef2", c_vp),
        ("model_log_variance", c_vp),
    ]


class LayerNormParams(C.Structure):
    _fields_ = [
        ("rows", c_i64), ("C", c_i32), ("rows_per_sample", c_i32),
        ("in_dtype", c_i32), ("out_dtype", c_i32), ("eps", c_f32), ("_pad", c_i32),
        ("x", c_vp), ("pre_add", c_vp), ("gamma", c_vp), ("beta", c_vp), ("y", c_vp), ("stats", c_vp),
    ]


P = C.POINTER
# name -> (restype, argtypes); every symbol declared in include/rho_b200.h
SIGNATURES = {
    "rho_abi_version": (C.c_int, []),
    "rho_error_string": (C.c_char_p, [C.c_int]),
    "rho_last_error_message": (C.c_char_p, []),
    "rho_launch_count": (c_u64, []),
    "rho_debug_attention_trace": (None, [c_vp]),
    "rho_nc_to_cl": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_i64, c_i64, c_i64, c_vp]),
    "rho_cl_to_nc": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_i64, c_i64, c_i64, c_vp]),
    "rho_add_cast": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_i64, c_vp]),
    "rho_scale": (C.c_int, [c_vp, C.c_int, c_i64, c_vp, c_f32, c_vp]),
    "rho_split3_concat": (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp]),
    "rho_split3_weight": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "rho_conv_pack_weight": (C.c_int, [c_vp, c_vp, C.c_int, c_i32, c_i32, c_i32, c_vp]),
    "rho_conv_aux_weight_bytes": (c_sz, [P(ConvParams)]),
    "rho_conv_pack_weight_aux": (C.c_int, [P(ConvParams), c_vp, c_vp, c_vp]),
    "rho_conv_workspace_bytes": (c_sz, [P(ConvParams)]),
    "rho_conv_stats_rows": (c_i32, [P(ConvParams)]),
    "rho_conv_fwd": (C.c_int, [P(ConvParams), c_vp]),
    "rho_conv_engine_for": (C.c_int, [P(ConvParams)]),
    "rho_conv_bwd_workspace_bytes": (c_sz, [P(ConvBwdParams)]),
    "rho_conv_dgrad": (C.c_int, [P(ConvBwdParams), c_vp]),
    "rho_conv_wgrad": (C.c_int, [P(ConvBwdParams), c_vp]),
    "rho_gn_workspace_bytes": (c_sz, [P(GnParams)]),
    "rho_gn_stats": (C.c_int, [P(GnParams), c_vp]),
    "rho_gn_stats_from_partials": (C.c_int, [P(GnParams), c_vp, c_i32, c_vp, c_i32, c_vp]),
    "rho_gn_apply": (C.c_int, [P(GnParams), c_vp]),
    "rho_gn_bwd_workspace_bytes": (c_sz, [P(GnBwdParams)]),
    "rho_gn_bwd": (C.c_int, [P(GnBwdParams), c_vp]),
    "rho_attention_fwd": (C.c_int, [P(AttnParams), c_vp]),
    "rho_attention_bwd_workspace_bytes": (c_sz, [P(AttnBwdParams)]),
    "rho_attention_bwd": (C.c_int, [P(AttnBwdParams), c_vp]),
    "rho_timestep_embedding": (C.c_int, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "rho_linear_fwd": (C.c_int, [P(LinearParams), c_vp]),
    "rho_linear_bwd": (C.c_int, [P(LinearParams), c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "rho_embedding_match": (C.c_int, [c_vp, c_i32, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp]),
    "rho_embedding_match_bwd": (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp]),
    "rho_randn": (C.c_int, [c_vp, c_i64, c_u64, c_u64, c_vp]),
    "rho_philox_raw": (C.c_int, [c_vp, c_i64, c_u64, c_u64, c_vp]),
    "rho_q_sample": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "rho_p_sample_step": (C.c_int, [P(PSampleParams), c_vp]),
    "rho_advance_timestep": (C.c_int, [c_vp, c_vp, c_i32, c_vp]),
    "rho_mse_workspace_bytes": (c_sz, [c_i64]),
    "rho_mse_loss": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_vp, c_vp, c_f32, c_vp, c_sz, c_vp]),
    "rho_adamw_ema_step": (C.c_int, [P(AdamWParams), c_vp]),
    "rho_ema_update": (C.c_int, [c_vp, c_vp, c_i64, c_f32, c_vp]),
    "rho_lincomb2": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "rho_abs_quantile_workspace_bytes": (c_sz, [c_i32]),
    "rho_abs_quantile": (C.c_int, [c_vp, c_i32, c_i64, c_i64, c_i64, c_f32, c_f32, c_vp, c_vp, c_sz, c_vp]),
    "rho_x0_step": (C.c_int, [P(X0StepParams), c_vp]),
    "rho_resample": (C.c_int, [c_vp, c_vp, C.c_int, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "rho_dropout": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_f32, c_u64, c_u64, c_vp, c_vp]),
    "rho_sph_harm_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "rho_sph_harm_density": (C.c_int, [c_vp, c_i32, c_i32, c_vp, c_vp, C.c_double, C.c_double, c_vp, c_vp, c_sz, c_vp]),
    "rho_gaussian_blobs": (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_u64, c_u64, c_vp]),
    "rho_lineprofile_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "rho_lineprofile": (C.c_int, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "rho_activation": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, C.c_int, c_vp]),
    "rho_activation_bwd": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_i64, C.c_int, c_vp]),
    "rho_convt_weight": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "rho_patchify": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "rho_bias_expand": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "rho_add_bcast": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "rho_colsum": (C.c_int, [c_vp, C.c_int, c_i64, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "rho_layernorm_fwd": (C.c_int, [P(LayerNormParams), c_vp]),
    "rho_layernorm_bwd_workspace_bytes": (c_sz, [P(LayerNormParams)]),
    "rho_layernorm_bwd": (C.c_int, [P(LayerNormParams), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
}

_lib = None


class RhoB200Error(RuntimeError):
    pass


def lib():
    """Loads librho_b200.so (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RhoB200Error(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()' or make -C rho-diffusion_b200/csrc)."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, op: str) -> None:
    if rc != RHO_OK:
        L = lib()
        raise RhoB200Error(f"{op}: {L.rho_error_string(rc).decode()} ({rc}): {L.rho_last_error_message().decode()}")


def launch_count() -> int:
    return int(lib().rho_launch_count())
