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


# ---- image-space helpers (csrc/image_ops.cu) ---------------------------------------------------------------------
def setup_filter(f=(1, 3, 3, 1)):
    """inference/ops.py:236-256: 1-D taps -> normalised 2-D outer-product filter (host-sized constant)."""
    f = torch.as_tensor(f, dtype=torch.float32)
    f = torch.outer(f, f)
    return f / f.sum()


def upfirdn2d(x, f, up=1, down=1, padding=(0, 0, 0, 0), gain=1):
    """inference/ops.py:87-114 on the device: x [B,C,H,W], f 2-D (or 1-D taps, applied separably = their outer product; None =
    identity), padding [x0, x1, y0, y1]."""
    lib = _lib.load()
    x = _cuda_f32(x, "x")
    gain = float(gain)
    if f is None:
        f = torch.ones(1, 1, device=x.device)
    f = _cuda_f32(f, "f")
    if f.ndim == 1:
        f, gain = torch.outer(f, f).contiguous(), gain       # ops.py:109-111 applies the taps per axis, each scaled by gain^(1/2)
    px0, px1, py0, py1 = [int(p) for p in padding]
    B, Cc, H, W = x.shape
    fh, fw = f.shape
    up, down = int(up), int(down)
    Ho = (H * up + py0 + py1 - fh) // down + 1
    Wo = (W * up + px0 + px1 - fw) // down + 1
    y = torch.empty(B, Cc, Ho, Wo, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.mb_upfirdn2d(_lib.ptr(x), _lib.ptr(f), _lib.ptr(y), B, Cc, H, W, fh, fw, up, down, px0, px1, py0, py1, gain,
                                    _lib.stream_ptr()))
    return y


def upsample2d(x, f, up=2, padding=0, gain=1):
    """inference/ops.py:117-133."""
    fw, fh = f.shape[-1], f.shape[0]
    p = (padding + (fw + up - 1) // 2, padding + (fw - up) // 2, padding + (fh + up - 1) // 2, padding + (fh - up) // 2)
    return upfirdn2d(x, f, up=up, padding=p, gain=gain * up * up)


def bias_act(x, b=None, act="linear", alpha=None, gain=None, clamp=None):
    """inference/ops.py:65-84 on the device (act: "linear" | "lrelu"; defaults alpha 0.2 and gain sqrt(2) for lrelu, :29-62)."""
    if act not in ("linear", "lrelu"):
        raise NotImplementedError(f"bias_act: activation '{act}' (built: linear, lrelu -- what the synthesis network uses)")
    lib = _lib.load()
    x4 = _cuda_f32(x, "x")
    if x4.ndim != 4:
        raise ValueError("bias_act: x must be [B, C, H, W]")
    b = _cuda_f32(b, "b")
    alpha = (0.2 if act == "lrelu" else 0.0) if alpha is None else float(alpha)
    gain = (2 ** 0.5 if act == "lrelu" else 1.0) if gain is None else float(gain)
    clamp = -1.0 if clamp is None else float(clamp)
    B, Cc, H, W = x4.shape
    y = torch.empty_like(x4)
    with torch.cuda.device(x4.device):
        _lib.check(lib.mb_bias_act(_lib.ptr(x4), _lib.ptr(b), _lib.ptr(y), B, Cc, H, W, int(act == "lrelu"), alpha, gain, clamp,
                                   _lib.stream_ptr()))
    return y


def resize_bicubic(x, size, align_corners=False):
    """F.interpolate(x, size, mode="bicubic", align_corners=...) on the device kernels: x [N,C,h,w] -> [N,C,H,W]."""
    lib = _lib.load()
    x = _cuda_f32(x, "x")
    n, c, h, w = x.shape
    oh, ow = int(size[0]), int(size[1])
    y = torch.empty(n, c, oh, ow, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.mb_resize_bicubic(_lib.ptr(x), _lib.ptr(y), n * c, h, w, oh, ow, int(bool(align_corners)), _lib.stream_ptr()))
    return y


def std_normalize_(x):
    """x /= x.std((1, 2, 3), keepdim=True) in place (maua/GAN/wrappers/stylegan2.py:212)."""
    if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError("std_normalize_: contiguous float32 CUDA tensor required")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_std_normalize(_lib.ptr(x), x.shape[0], x[0].numel(), _lib.stream_ptr()))
    return x


def _lanczos_taps(ratio, a=2):
    """lanczos(ramp(ratio, 2), 2) of maua/ops/image.py:196-211 (host design of the prefilter taps)."""
    import math

    n = math.ceil(2 / ratio + 1)
    out = torch.empty([n])
    cur = 0
    for i in range(n):
        out[i] = cur
        cur += ratio
    x = torch.cat([-out[1:].flip([0]), out])[1:-1]
    sinc = lambda v: torch.where(v != 0, torch.sin(math.pi * v) / (math.pi * v), v.new_ones([]))  # noqa: E731
    cond = torch.logical_and(-a < x, x < a)
    k = torch.where(cond, sinc(x) * sinc(x / a), x.new_zeros([]))
    return (k / k.sum()).contiguous()


def resample(input, size, align_corners=True):
    """maua/ops/image.py:214-240: Lanczos-2 prefilter along each shrinking axis (reflect padding), then bicubic."""
    lib = _lib.load()
    x = _cuda_f32(input, "input")
    n, c, h, w = x.shape
    if isinstance(size, (int, float)):
        short, long = (w, h) if w <= h else (h, w)
        new_short, new_long = round(size), round(size * long / short)
        dw, dh = (new_short, new_long) if w <= h else (new_long, new_short)
    else:
        dh, dw = size
    with torch.cuda.device(x.device):
        for axis, (dst, src) in enumerate(((dh, h), (dw, w))):
            if dst < src:
                k = _lanczos_taps(dst / src).to(x.device)
                y = torch.empty_like(x)
                _lib.check(lib.mb_fir_reflect(_lib.ptr(x), _lib.ptr(y), n * c, h, w, _lib.ptr(k), k.numel(), axis, _lib.stream_ptr()))
                x = y
    return resize_bicubic(x, (dh, dw), align_corners=align_corners)


def perlin_noise(shape, res, tileable=(True, False, False), rng=None):
    """maua/ops/noise.py:27-88: 3-D gradient noise [shape] in [-1, 1] on the device.  Gradient angles are drawn on the host
    from `rng` (np.random by default, exactly as the reference draws them); `res` must divide `shape` (the reference snaps
    it to the closest divisor with a random tie-break, ops/noise.py:14-20: pass an exact divisor)."""
    import numpy as np

    if any(s % r for s, r in zip(shape, res)):
        raise ValueError("perlin_noise: res must divide shape")
    rnd = np.random if rng is None else rng
    theta = 2 * np.pi * rnd.rand(res[0] + 1, res[1] + 1, res[2] + 1).astype(np.float32)
    phi = 2 * np.pi * rnd.rand(res[0] + 1, res[1] + 1, res[2] + 1).astype(np.float32)
    g = np.stack((np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)), axis=3)
    if tileable[0]:
        g[-1, :, :] = g[0, :, :]
    if tileable[1]:
        g[:, -1, :] = g[:, 0, :]
    if tileable[2]:
        g[:, :, -1] = g[:, :, 0]
    grad = torch.from_numpy(np.ascontiguousarray(g.astype(np.float32))).cuda()
    out = torch.empty(tuple(shape), device=grad.device)
    with torch.cuda.device(grad.device):
        _lib.check(_lib.load().mb_perlin_noise(_lib.ptr(grad), shape[0], shape[1], shape[2], res[0], res[1], res[2], _lib.ptr(out), _lib.stream_ptr()))
    return out
