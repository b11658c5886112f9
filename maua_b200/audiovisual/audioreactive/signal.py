"""Envelope post-ops on the device: mirror of maua/audiovisual/audioreactive/signal.py (same names/arguments)."""
from __future__ import annotations

import torch

from ... import _lib


def _cuda32(x, name):
    if not x.is_cuda:
        raise RuntimeError(f"maua_b200.audioreactive: '{name}' must be a CUDA tensor (no CPU fallback)")
    return x.detach().to(torch.float32).contiguous()


def gaussian_filter(x, sigma, causal=None, mode="circular"):
    """Smooth along the time (first) axis with a Gaussian kernel, circular padding (signal.py:108-157)."""
    if mode != "circular":
        raise NotImplementedError("gaussian_filter: only circular padding (the reference default)")
    x = _cuda32(x, "x")
    T = x.shape[0]
    y = torch.empty_like(x)
    causal_mode = int(causal is not None)
    factor = float(causal) if isinstance(causal, float) else 0.0
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_gaussian_filter(_lib.ptr(x), _lib.ptr(y), T, x.numel() // T, float(sigma), causal_mode,
                                                  factor, _lib.stream_ptr()))
    return y


def normalize(x, eps=0.0):
    """(x - min) / (max - min) over the whole tensor (signal.py:27-38; eps=1e-8 gives processing.py:53-56)."""
    x = _cuda32(x, "x")
    y = torch.empty_like(x)
    scratch = torch.empty(2, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_normalize(_lib.ptr(x), _lib.ptr(y), x.numel(), float(eps), _lib.ptr(scratch), _lib.stream_ptr()))
    return y


def resample(x, size):
    """Linear resampling along the time (first) axis (signal.py:5-24)."""
    x = _cuda32(x, "x")
    xs = x.squeeze()
    T = xs.shape[0]
    Cc = xs.numel() // T
    y = torch.empty((size,) + tuple(xs.shape[1:]), device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_resample_linear(_lib.ptr(xs.contiguous()), _lib.ptr(y), T, int(size), Cc, _lib.stream_ptr()))
    return y


def percentile(signal, p):
    """k-th smallest value with k = 1 + round(p/100 (n-1)) (signal.py:41-52); host scalar like the reference."""
    k = 1 + round(0.01 * float(p) * (signal.numel() - 1))
    return signal.reshape(-1).kthvalue(k).values.item()


def percentile_clip(signal, percent):
    """Clamp to the percentile of the strict local maxima, renormalise by the max (signal.py:55-81).
    Peak selection and the order statistic use torch device ops (index bookkeeping, not arithmetic)."""
    signal = _cuda32(signal, "signal")
    if signal.ndim < 2:
        signal = signal.unsqueeze(1)
    cols = []
    n = signal.shape[0]
    i = torch.arange(n, device=signal.device)
    for sig in signal.unbind(1):
        peaks = (sig > sig[(i + 1).clamp(0, n - 1)]) & (sig > sig[(i - 1).clamp(0, n - 1)])
        sig = sig.clamp(0, percentile(sig[peaks], percent))
        cols.append(sig / sig.max())
    return torch.stack(cols, dim=1)


def compress(signal, threshold, ratio, invert=False):
    """signal.py:84-100 (thresholded scaling, then normalize)."""
    s = _cuda32(signal, "signal").clone()
    m = s < threshold if invert else s > threshold
    s[m] = s[m] * ratio
    return normalize(s)


def expand(signal, threshold, ratio, invert=False):
    return compress(signal, threshold, ratio, invert)
