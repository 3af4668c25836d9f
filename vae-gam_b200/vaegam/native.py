"""ctypes binding of libvaegam_sm100.so (the C ABI declared in include/vaegam.h).

There is NO fallback: if the shared object is missing, or CUDA is not available when a
compute entry point is called, this module raises.  (The reference's arithmetic lives in
PyTorch library kernels — vae_reg_GP.py:236-264,307-413, gp.py:41-110 — and this library is
what replaces them on B200.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvaegam_sm100.so")

VG_NUM_PARAMS = 97
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
# VG_ARITH_* (include/vaegam.h): per-call arithmetic of the convolution kernels
ARITH_DEFAULT, ARITH_FP32, ARITH_BF16, ARITH_MIXED = 0, 1, 2, 3
BF16_X, BF16_Y, BF16_DX = 1, 2, 4            # VgConvDesc.bf16_mask
ARITH_BY_NAME = {"default": ARITH_DEFAULT, "fp32": ARITH_FP32, "bf16": ARITH_BF16, "mixed": ARITH_MIXED}
VG_BWD_PHASES = 3
VG_OK, VG_EINVAL = 0, -1
V = 41 * 49 * 35
VP = (V + 3) // 4 * 4


class NativeError(RuntimeError):
    pass


class VgConvDesc(C.Structure):
    _fields_ = [("transposed", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32 * 3),
                ("stride", C.c_int32), ("pad", C.c_int32 * 3), ("opad", C.c_int32 * 3), ("in_", C.c_int32 * 3),
                ("out", C.c_int32 * 3), ("n", C.c_int32), ("group_size", C.c_int32), ("arith", C.c_int32),
                ("bf16_mask", C.c_int32), ("reserved_", C.c_int32), ("x_img_stride", C.c_int64), ("y_img_stride", C.c_int64)]


class VgGainParams(C.Structure):
    _fields_ = [("sa", C.c_void_p * 8), ("logstd", C.c_void_p * 8), ("qu_m", C.c_void_p * 8),
                ("qu_S", C.c_void_p * 8), ("logkvar", C.c_void_p * 8), ("logls", C.c_void_p * 8),
                ("xu", C.c_void_p * 8), ("has_gp", C.c_int32 * 8), ("hrf", C.c_int32 * 8)]


class VgGainGrads(C.Structure):
    _fields_ = [("sa", C.c_void_p * 8), ("logstd", C.c_void_p * 8), ("qu_m", C.c_void_p * 8),
                ("qu_S", C.c_void_p * 8), ("logkvar", C.c_void_p * 8), ("logls", C.c_void_p * 8)]


class VgStepConfig(C.Structure):
    _fields_ = [("b", C.c_int32), ("m", C.c_int32), ("neural_covariates", C.c_int32), ("want_maps", C.c_int32),
                ("gp_kl_scale", C.c_float), ("glm_reg_scale", C.c_float), ("arith", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class VgMlpLayer(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("dw", C.c_void_p), ("db", C.c_void_p),
                ("n", C.c_int32), ("k", C.c_int32), ("in_", C.c_int32), ("out", C.c_int32), ("act", C.c_int32),
                ("pad_", C.c_int32)]


class VgMlpBuf(C.Structure):
    _fields_ = [("act", C.c_void_p), ("grad", C.c_void_p), ("width", C.c_int32), ("role", C.c_int32)]


class VgMlp(C.Structure):
    _fields_ = [("nlayers", C.c_int32), ("nbufs", C.c_int32), ("rows", C.c_int32), ("rows_per_cta", C.c_int32),
                ("layer", VgMlpLayer * 8), ("buf", VgMlpBuf * 10)]


MLP_INPUT, MLP_GRAD_IN, MLP_GRAD_OUT = 1, 2, 4


class VgStepIO(C.Structure):
    _fields_ = [("params", C.c_void_p * VG_NUM_PARAMS), ("grads", C.c_void_p * VG_NUM_PARAMS),
                ("xu", C.c_void_p * 6), ("glm_t", C.c_void_p), ("taps", C.c_void_p), ("x", C.c_void_p),
                ("covariates", C.c_void_p), ("eps_w", C.c_void_p), ("eps_d", C.c_void_p), ("eps_g", C.c_void_p),
                ("out_scalars", C.c_void_p), ("z", C.c_void_p), ("maps", C.c_void_p), ("g", C.c_void_p),
                ("cons", C.c_void_p), ("x_rec", C.c_void_p), ("beta_mean", C.c_void_p), ("beta_var", C.c_void_p),
                ("status", C.c_void_p)]


_P = C.c_void_p
_I = C.c_int
_LL = C.c_longlong
_SZ = C.c_size_t
_F = C.c_float
_D = C.c_double

# name -> (restype, argtypes).  Every symbol include/vaegam.h declares is listed here; the
# CPU test-suite checks the two stay in sync.
SIGNATURES = {
    "vg_version": (_I, []),
    "vg_last_error": (C.c_char_p, []),
    "vg_sm_count": (_I, []),
    "vg_launch_count": (_LL, []),
    "vg_profile_enable": (_I, [_I]),
    "vg_profile_collect": (_LL, [C.c_char_p, _SZ]),
    "vg_set_conv_mode": (_I, [_I]),
    "vg_get_conv_mode": (_I, []),
    "vg_set_conv_tuning": (_I, [C.c_char_p, _LL]),
    "vg_conv_describe": (_I, [C.POINTER(VgConvDesc), _I, C.c_char_p, _SZ]),
    "vg_conv_fwd": (_I, [C.POINTER(VgConvDesc), _P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "vg_conv_dgrad": (_I, [C.POINTER(VgConvDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vg_conv_wgrad": (_I, [C.POINTER(VgConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "vg_box_sums": (_I, [_P, _I, _I, C.POINTER(C.c_int32), _LL, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P]),
    "vg_conv_wgrad_grouped": (_I, [C.POINTER(VgConvDesc), _P, _P, _P, _P]),
    "vg_bn_fused_finalize": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _D, _P, _P, _P, _P, _P, _P]),
    "vg_conv_dgrad_bn_apply": (_I, [C.POINTER(VgConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "vg_bn_stats": (_I, [_P, _I, _I, _LL, _I, _P, _P]),
    "vg_bn_finalize": (_I, [_P, _P, _P, _I, _I, _D, _P, _P, _P, _P, _P]),
    "vg_bn_bwd_apply": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _LL, _I, _D, _I, _I, _P, _P, _P, _P, _P]),
    "vg_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _LL, _P]),
    "vg_nhwc_to_nchw": (_I, [_P, _P, _I, _I, _LL, _P]),
    "vg_linear_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "vg_linear_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "vg_mlp_fwd": (_I, [C.POINTER(VgMlp), _P]),
    "vg_mlp_bwd": (_I, [C.POINTER(VgMlp), _P]),
    "vg_latent_fwd": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    "vg_latent_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "vg_gain_workspace_bytes": (_SZ, [_I, _I]),
    "vg_gain_fwd": (_I, [C.POINTER(VgGainParams), _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "vg_gain_bwd": (_I, [C.POINTER(VgGainParams), C.POINTER(VgGainGrads), _P, _P, _P, _P, _D, _I, _I, _P, _SZ, _P]),
    "vg_gp_posterior": (_I, [_P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "vg_recon_workspace_bytes": (_SZ, [_I, _LL]),
    "vg_recon_tune": (None, [_I]),
    "vg_recon_plan": (_I, [_I, _LL, _I, C.POINTER(C.c_int)]),
    "vg_recon_loss_fwd": (_I, [_P, _P, _P, _P, _P, _I, _LL, _P, _P, _P, _P, _P, _SZ, _P]),
    "vg_recon_loss_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _LL, _F, _P, _P, _P, _P, _SZ, _P]),
    "vg_adam_step": (_I, [_P, _P, _P, _P, _LL, _P, _P, _P, _P, _LL, _D, _D, _D, _D, _D, _P, _P, _I, _P]),
    "vg_step_workspace_bytes": (_SZ, [C.POINTER(VgStepConfig)]),
    "vg_step_fwd": (_I, [C.POINTER(VgStepConfig), C.POINTER(VgStepIO), _P, _SZ, _P]),
    "vg_step_bwd": (_I, [C.POINTER(VgStepConfig), C.POINTER(VgStepIO), _P, _SZ, _P]),
    "vg_step_bwd_phase": (_I, [C.POINTER(VgStepConfig), C.POINTER(VgStepIO), _P, _SZ, _I, _P, _P]),
    "vg_encode_fwd": (_I, [C.POINTER(VgStepConfig), C.POINTER(VgStepIO), _P, _P, _SZ, _P]),
    "vg_decode_workspace_bytes": (_SZ, [_I]),
    "vg_decode_fwd": (_I, [C.POINTER(VgStepIO), _P, _I, _I, _P, _P, _SZ, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared object (once).  Raises NativeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `make -C vae-gam_b200/csrc` (or __graft_entry__.build()). "
            "There is no CPU or PyTorch fallback for the VAE-GAM hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise NativeError("CUDA device required: the VAE-GAM hot path runs only as sm_100a kernels "
                          "(no CPU path by design)")


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().vg_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'vaegam native call'} failed (rc={rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().vg_launch_count())


def profile(enable: bool):
    check(load().vg_profile_enable(int(enable)), "vg_profile_enable")


def profile_collect():
    """{op name: [ms, ...]} for every operation recorded since profile(True)."""
    buf = C.create_string_buffer(1 << 22)
    load().vg_profile_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms = line.split("\t")
        out.setdefault(name, []).append(float(ms))
    return out


def conv_desc(transposed, cin, cout, k, stride, in_, n, group_size, pad=(0, 0, 0), opad=(0, 0, 0),
              x_img_stride=0, y_img_stride=0, arith=ARITH_DEFAULT, bf16_mask=0) -> VgConvDesc:
    d = VgConvDesc()
    d.arith = arith
    d.bf16_mask = bf16_mask
    d.transposed, d.cin, d.cout, d.stride = int(transposed), cin, cout, stride
    for i in range(3):
        d.k[i], d.pad[i], d.opad[i], d.in_[i] = k[i], pad[i], opad[i], in_[i]
        d.out[i] = ((in_[i] - 1) * stride - 2 * pad[i] + k[i] + opad[i]) if transposed else ((in_[i] - k[i]) // stride + 1)
    d.n, d.group_size = n, group_size
    d.x_img_stride, d.y_img_stride = x_img_stride, y_img_stride
    return d
