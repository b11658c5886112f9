"""Op-level host wrappers over libmaua_b200 (the same kernels the network pipeline launches).

Signatures follow the upstream ops the reference reaches through maua/GAN/nv
(torch_utils/ops/filtered_lrelu.py::filtered_lrelu, networks_stylegan3.py::modulated_conv2d).
CUDA tensors only; there is no CPU implementation here.
"""
from __future__ import annotations

import torch

from . import _lib


def _cuda_f32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"maua_b200.ops: '{name}' must be a CUDA tensor (no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


def modulated_conv2d(x, w, s, demodulate=True, padding=None, input_gain=None, impl=0):
    """x [B,Cin,H,W], w [Cout,Cin,k,k], s [B,Cin] -> [B,Cout,H+k-1,W+k-1] (padding is always k-1)."""
    lib = _lib.load()
    x, w, s = _cuda_f32(x, "x"), _cuda_f32(w, "w"), _cuda_f32(s, "s")
    B, Cin, H, W = x.shape
    Cout, Cin2, k, k2 = w.shape
    if Cin2 != Cin or k != k2 or tuple(s.shape) != (B, Cin):
        raise ValueError("modulated_conv2d: shape mismatch")
    if padding is not None and padding != k - 1:
        raise ValueError("modulated_conv2d: only padding = k-1 (StyleGAN3) is implemented")
    gain = 1.0 if input_gain is None else float(input_gain)
    y = torch.empty(B, Cout, H + k - 1, W + k - 1, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.mb_modulated_conv2d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(s), _lib.ptr(y), B, Cin, Cout, H, W, k,
                                           int(bool(demodulate)), gain, int(impl), _lib.stream_ptr()))
    return y


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=2 ** 0.5, slope=0.2, clamp=None):
    """bias -> upsample(fu) -> lrelu*gain -> clamp -> downsample(fd).  padding = int | [x0,x1,y0,y1]."""
    lib = _lib.load()
    x = _cuda_f32(x, "x")
    fu, fd, b = _cuda_f32(fu, "fu"), _cuda_f32(fd, "fd"), _cuda_f32(b, "b")
    B, Cc, H, W = x.shape
    if isinstance(padding, int):
        padding = [padding] * 4
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    px0, px1, py0, py1 = [int(p) for p in padding]
    up_taps = 1 if fu is None else fu.shape[-1]
    down_taps = 1 if fd is None else fd.shape[-1]
    fd_2d = int(fd is not None and fd.ndim == 2)
    if fu is not None and fu.ndim != 1:
        raise ValueError("filtered_lrelu: only separable (1-D) up filters are implemented")
    Ho = (H * up + py0 + py1 - (up_taps - 1) - (down_taps - 1) + (down - 1)) // down
    Wo = (W * up + px0 + px1 - (up_taps - 1) - (down_taps - 1) + (down - 1)) // down
    y = torch.empty(B, Cc, Ho, Wo, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.mb_filtered_lrelu(_lib.ptr(x), _lib.ptr(fu), _lib.ptr(fd), _lib.ptr(b), _lib.ptr(y), B, Cc, H, W,
                                         int(up), int(down), up_taps, down_taps, fd_2d, px0, px1, py0, py1, float(gain),
                                         float(slope), float(-1 if clamp is None else clamp), _lib.stream_ptr()))
    return y
