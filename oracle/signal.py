"""Oracle: envelope post-ops and latent sequencers of the reference, fp32 on CPU.  TEST INFRASTRUCTURE ONLY.

Restates maua/audiovisual/audioreactive/signal.py (rows a7, a10) and latent.py:12-80 (row a11).
PINNED for signal.py (pure torch, imports here): tests/golden/make_audio_golden.py checks against the reference
module itself.  latent.py imports torchcubicspline / torchtyping (absent): its pure-torch functions are
restated from the source and pinned through a stubbed import in the same script.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def resample(x, size):
    """signal.py:5-24: linear interpolation along the first (time) axis, align_corners False."""
    y = x.squeeze()
    if y.ndim == 1:
        y = y[None, None]
    elif y.ndim == 2:
        y = y.t()[None]
    else:
        y = y.permute(1, 2, 0)
    return F.interpolate(y, size=size, mode="linear", align_corners=False).permute(2, 0, 1).squeeze()


def normalize(x):
    """signal.py:27-38 (no epsilon, unlike processing.normalize)."""
    y = x - x.min()
    return y / y.max()


def percentile(x, p):
    """signal.py:41-52: k-th smallest with k = 1 + round(p/100 (n-1))."""
    k = 1 + round(0.01 * float(p) * (x.numel() - 1))
    return x.reshape(-1).kthvalue(k).values.item()


def peak_mask(sig):
    n = sig.shape[0]
    i = torch.arange(n)
    return (sig > sig[(i + 1).clamp(0, n - 1)]) & (sig > sig[(i - 1).clamp(0, n - 1)])


def percentile_clip(signal, percent):
    """signal.py:55-81: clamp each column to the percentile of its strict local maxima, renormalise by the max."""
    if signal.ndim < 2:
        signal = signal.unsqueeze(1)
    cols = []
    for sig in signal.unbind(1):
        sig = sig.clamp(0, percentile(sig[peak_mask(sig)], percent))
        cols.append(sig / sig.max())
    return torch.stack(cols, dim=1)


def compress(signal, threshold, ratio, invert=False):
    """signal.py:84-100 (without the reference's in-place aliasing of its argument)."""
    s = signal.clone()
    m = s < threshold if invert else s > threshold
    s[m] = s[m] * ratio
    return normalize(s)


def gaussian_kernel(sigma, n_frames, causal=None):
    radius = min(int(sigma * 4), 3 * n_frames)
    k = torch.exp(-0.5 / sigma ** 2 * torch.arange(-radius, radius + 1, dtype=torch.float32) ** 2)
    if causal is not None:
        k[radius + 1:] *= 0 if not isinstance(causal, float) else causal
    return k / k.sum(), radius


def gaussian_filter(x, sigma, causal=None):
    """signal.py:108-157: depthwise temporal Gaussian, circular padding (radius <= n_frames branch)."""
    shape = x.shape
    T = shape[0]
    k, radius = gaussian_kernel(sigma, T, causal)
    if radius > T:
        raise NotImplementedError("short-sequence branch (radius > n_frames) is not exercised by the render path")
    y = x.reshape(T, -1).t()[None]                      # [1, C, T]
    y = F.pad(y, (radius, radius), mode="circular")
    y = F.conv1d(y, k.view(1, 1, -1).repeat(y.shape[1], 1, 1), groups=y.shape[1])
    return y[0].t().reshape(shape)


def single_weighted(low, high, envelope):
    """latent.py:12-17."""
    e = envelope[:, None, None]
    return low[None] * (1 - e) + high[None] * e


def multi_weighted(latents, envelopes):
    """latent.py:21-31: normalised envelope mix of the key latents (einsum over the latent index)."""
    w = envelopes / envelopes.sum(dim=1, keepdim=True)
    sel = latents[torch.arange(w.shape[1]) % len(latents)]
    return torch.einsum("ta,awl->twl", w, sel)


def slerp(a, b, t):
    """latent.py:54-65."""
    a = a / a.norm(dim=-1, keepdim=True)
    b = b / b.norm(dim=-1, keepdim=True)
    d = (a * b).sum(dim=-1, keepdim=True)
    p = (t * torch.acos(d)).permute(2, 0, 1)[..., None]
    c = b - d * a
    c = c / c.norm(dim=-1, keepdim=True)
    out = a[None] * torch.cos(p) + c[None] * torch.sin(p)
    return out / out.norm(dim=-1, keepdim=True)


def slerp_loops(y, size, n_loops):
    """latent.py:68-80."""
    y = torch.cat([y] * n_loops + [y[[0]]])
    t = torch.linspace(0, 1, round(size / len(y))).to(y)
    out = slerp(y[:-1], y[1:], t)
    out = out.reshape(-1, *out.shape[2:])
    return F.interpolate(out.permute(1, 2, 0), size=size, mode="linear", align_corners=False).permute(2, 0, 1)


def select_modulo(latents, envelope, smooth=2):
    """latent.py:34-45: quartile clamp -> normalise -> round to a key index -> gather -> causal=0 Gaussian."""
    low, high = torch.quantile(envelope, 0.25), torch.quantile(envelope, 0.75)
    indices = normalize(envelope.clamp(low, high))
    indices = (indices * (len(latents) - 1)).round().long()
    return gaussian_filter(latents[indices], smooth, causal=0)


def spline_loops(y, size, n_loops):
    """latent.py:83-92.  The reference evaluates torchcubicspline.NaturalCubicSpline (third-party, un-pinned in
    setup.py:104, absent from the checkout): the published algorithm -- the C2 piecewise cubic through the knots with
    zero second derivative at both ends -- is restated here in float64 and checked against scipy's
    CubicSpline(bc_type="natural") in tests/test_oracle_audio.py.  PARITY UNPINNED against the reference's own call."""
    import numpy as np

    y = torch.cat([y] * n_loops + [y[[0]]])
    m = len(y)
    yy = y.reshape(m, -1).double().numpy()
    h = 1.0 / (m - 1)
    z = np.zeros_like(yy)
    if m > 2:
        a = np.zeros((m - 2, m - 2))
        np.fill_diagonal(a, 4.0)
        idx = np.arange(m - 3)
        a[idx, idx + 1] = 1.0
        a[idx + 1, idx] = 1.0
        rhs = 6.0 / h ** 2 * (yy[:-2] - 2 * yy[1:-1] + yy[2:])
        z[1:-1] = np.linalg.solve(a, rhs)
    t = np.linspace(0.0, 1.0, size)
    u = t * (m - 1)
    i = np.minimum(np.floor(u).astype(np.int64), m - 2)
    f = (u - i)[:, None]
    g = 1.0 - f
    out = g * yy[i] + f * yy[i + 1] + h * h / 6.0 * ((g ** 3 - g) * z[i] + (f ** 3 - f) * z[i + 1])
    return torch.from_numpy(out).to(y.dtype).reshape(size, *y.shape[1:])


def tempo_loops(latents, n_frames, fps, tempo, type="spline"):
    """latent.py:95-102."""
    n_loops = round(n_frames / fps * (tempo / 4 / 60))
    return spline_loops(latents, n_frames, n_loops) if type == "spline" else slerp_loops(latents, n_frames, n_loops)
