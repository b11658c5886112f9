"""ctypes binding of libmaua_b200.so (C ABI declared in include/maua_b200.h).

There is NO fallback: if the shared library is missing this module raises on import of the
symbol table, and every compute entry point raises RuntimeError when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB_PATH

MB_OUT_F32_NCHW = 0
MB_OUT_F32_NCHW_01 = 1
MB_OUT_U8_NHWC = 2
MB_OUT_F32_NCHW_UNIT = 3
MB_RESIZE_NONE, MB_RESIZE_STRETCH, MB_RESIZE_PAD_ZERO = 0, 1, 2


class SG3Cfg(C.Structure):
    _fields_ = [
        ("w_dim", C.c_int32), ("img_resolution", C.c_int32), ("img_channels", C.c_int32),
        ("channel_base", C.c_int32), ("channel_max", C.c_int32), ("num_layers", C.c_int32),
        ("num_critical", C.c_int32), ("conv_kernel", C.c_int32), ("filter_size", C.c_int32),
        ("lrelu_upsampling", C.c_int32), ("use_radial_filters", C.c_int32), ("margin_size", C.c_int32),
        ("first_cutoff", C.c_double), ("first_stopband", C.c_double), ("last_stopband_rel", C.c_double),
        ("output_scale", C.c_double), ("conv_clamp", C.c_double),
    ]


class SG3Layer(C.Structure):
    _fields_ = [
        ("idx", C.c_int32), ("is_torgb", C.c_int32), ("is_critically_sampled", C.c_int32), ("use_fp16", C.c_int32),
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("in_size", C.c_int32), ("out_size", C.c_int32),
        ("in_sampling_rate", C.c_int32), ("out_sampling_rate", C.c_int32), ("tmp_sampling_rate", C.c_int32),
        ("conv_kernel", C.c_int32), ("up", C.c_int32), ("down", C.c_int32), ("up_taps", C.c_int32),
        ("down_taps", C.c_int32), ("down_radial", C.c_int32), ("pad_lo", C.c_int32), ("pad_hi", C.c_int32),
        ("in_cutoff", C.c_double), ("out_cutoff", C.c_double), ("in_half_width", C.c_double),
        ("out_half_width", C.c_double), ("name", C.c_char * 32),
    ]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "mb_init": (C.c_int, [C.c_int]),
    "mb_last_error": (C.c_char_p, []),
    "mb_abi_version": (C.c_int, []),
    "mb_sg3_default_cfg": (None, [C.POINTER(SG3Cfg), C.c_int]),
    "mb_sg3_geometry": (C.c_int, [C.POINTER(SG3Cfg), C.POINTER(SG3Layer), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mb_sg3_create": (C.c_int, [C.POINTER(SG3Cfg), C.POINTER(_P)]),
    "mb_net_destroy": (None, [_P]),
    "mb_sg2_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "mb_net_num_ws": (C.c_int, [_P]),
    "mb_net_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, _P]),
    "mb_net_finalize": (C.c_int, [_P, _P]),
    "mb_net_workspace_bytes": (C.c_size_t, [_P, C.c_int]),
    "mb_net_forward": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, _P, C.c_size_t, _P]),
    "mb_net_forward_xf": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, _P, C.c_size_t, _P]),
    "mb_net_set_resize": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mb_sg3_resized_output": (C.c_int, [C.POINTER(SG3Cfg), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mb_sg2_set_warps": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), _P, C.c_int]),
    "mb_rrdb_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "mb_rrdb_destroy": (None, [_P]),
    "mb_rrdb_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, _P]),
    "mb_rrdb_finalize": (C.c_int, [_P, _P]),
    "mb_rrdb_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int, C.c_int]),
    "mb_rrdb_forward": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, C.c_size_t, _P]),
    "mb_rrdb_last_launch_count": (C.c_int, [_P]),
    "mb_sg2_set_resize": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "mb_gaussian_filter_ex": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, _P]),
    "mb_salience": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    "mb_latent_merge": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mb_spline_loop_latents": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, _P]),
    "mb_audio_spectrogram": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_size_t, _P]),
    "mb_spectral_flatness": (C.c_int, [_P, C.c_int, C.c_float, C.c_float, _P, _P]),
    "mb_spectral_contrast": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, _P, _P, _P]),
    "mb_mfcc": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "mb_net_output_shape": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mb_net_read_activation": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "mb_net_last_launch_count": (C.c_int, [_P]),
    "mb_net_set_conv_impl": (C.c_int, [_P, C.c_int]),
    "mb_net_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "mb_net_profile_read": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]),
    "mb_net_activation_shape": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mb_audio_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "mb_audio_onsets_rms": (C.c_int, [_P, C.c_int64, _P, C.c_float, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "mb_audio_hpss_component": (C.c_int, [_P, C.c_int64, C.c_float, C.c_int, _P, _P, C.c_size_t, _P]),
    "mb_chroma_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "mb_chroma_cqt": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P,
                                C.c_int, C.c_float, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "mb_tuning_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "mb_estimate_tuning": (C.c_int, [_P, C.c_int64, C.c_float, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "mb_chroma_cens_post": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, _P, _P]),
    "mb_gaussian_filter": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, _P]),
    "mb_normalize": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P, _P]),
    "mb_resample_linear": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "mb_quantile_mid": (C.c_int, [_P, C.c_int64, C.c_float, _P, _P]),
    "mb_frames_to_rgb24": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P]),
    "mb_upfirdn2d": (C.c_int, [_P, _P, _P] + [C.c_int] * 12 + [C.c_float, _P]),
    "mb_bias_act": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _P]),
    "mb_sosfilt": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_double), C.c_int, _P, _P]),
    "mb_multi_weighted": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mb_single_weighted": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "mb_slerp_rows": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "mb_spline_loops": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "mb_select_modulo": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "mb_noise_mix": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "mb_noise_loop": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P, _P]),
    "mb_noise_combine": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _P, _P]),
    "mb_resize_bicubic": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mb_fir_reflect": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P]),
    "mb_std_normalize": (C.c_int, [_P, C.c_int, C.c_int64, _P]),
    "mb_perlin_noise": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "mb_debug_read": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "mb_modulated_conv2d": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_int, _P]),
    "mb_filtered_lrelu": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                    C.c_float, _P]),
}

_lib = None


def signatures():
    return dict(_SIGNATURES)


def register(extra):
    """Other host modules (audio, signal) add their entry points to the table before load()."""
    _SIGNATURES.update(extra)
    if _lib is not None:
        for name, (res, args) in extra.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args


def load():
    """Load the shared library; raise if it was not built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m maua_b200._build` "
                "(or __graft_entry__.build()); maua_b200 has no CPU / PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code: int):
    if code != 0:
        msg = load().mb_last_error()
        raise RuntimeError(f"libmaua_b200 error {code}: {msg.decode() if msg else ''}")


def debug_words(n=24):
    """Debug words kernels wrote to mapped host memory (MB_DEBUG=1); readable after a failed launch."""
    buf = (C.c_int * n)()
    load().mb_debug_read(buf, n)
    return list(buf)


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
