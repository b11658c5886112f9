"""Latent sequencers on the device: mirror of maua/audiovisual/audioreactive/latent.py:12-31."""
from __future__ import annotations

import torch

from ... import _lib
from .signal import _cuda32


def single_weighted(low_latent, high_latent, envelope):
    """low * (1 - e) + high * e -> [T, n_layers, latent_dim] (latent.py:12-17)."""
    low, high, env = _cuda32(low_latent, "low_latent"), _cuda32(high_latent, "high_latent"), _cuda32(envelope, "envelope")
    T, D = env.shape[0], low.numel()
    out = torch.empty((T,) + tuple(low.shape), device=low.device)
    with torch.cuda.device(low.device):
        _lib.check(_lib.load().mb_single_weighted(_lib.ptr(low), _lib.ptr(high), _lib.ptr(env.reshape(-1)), _lib.ptr(out), T, D,
                                                  _lib.stream_ptr()))
    return out


def multi_weighted(latents, envelopes):
    """Envelope-normalised mix of the key latents -> [T, n_layers, latent_dim] (latent.py:21-31)."""
    lat, env = _cuda32(latents, "latents"), _cuda32(envelopes, "envelopes")
    T, A = env.shape
    K = lat.shape[0]
    D = lat[0].numel()
    out = torch.empty((T,) + tuple(lat.shape[1:]), device=lat.device)
    with torch.cuda.device(lat.device):
        _lib.check(_lib.load().mb_multi_weighted(_lib.ptr(lat), _lib.ptr(env), _lib.ptr(out), T, A, K, D, _lib.stream_ptr()))
    return out


def select_modulo(latents, envelope, smooth=2):
    """Pick the key latent indexed by the quartile-clamped, normalised envelope, then smooth causally
    (latent.py:34-45) -> [T, n_layers, latent_dim]."""
    from .signal import gaussian_filter

    lat, env = _cuda32(latents, "latents"), _cuda32(envelope, "envelope").reshape(-1)
    T, K, Cc = env.shape[0], lat.shape[0], lat[0].numel()
    out = torch.empty((T,) + tuple(lat.shape[1:]), device=lat.device)
    with torch.cuda.device(lat.device):
        srt = torch.sort(env).values.contiguous()      # order statistics: index bookkeeping, the arithmetic is in the kernel
        _lib.check(_lib.load().mb_select_modulo(_lib.ptr(env), _lib.ptr(srt), T, _lib.ptr(lat), K, Cc, _lib.ptr(out), _lib.stream_ptr()))
    return gaussian_filter(out, smooth, causal=0)


def slerp_loops(y, size, n_loops):
    """Spherical interpolation through the looped key latents, resampled to `size` frames (latent.py:68-80)."""
    from .signal import resample

    lat = _cuda32(y, "y")
    K, L, D = lat.shape
    steps = round(size / (K * n_loops + 1))
    if steps < 1:
        raise ValueError("slerp_loops: size is smaller than the number of looped keys")
    rows = torch.empty(steps * K * n_loops, L, D, device=lat.device)
    with torch.cuda.device(lat.device):
        _lib.check(_lib.load().mb_slerp_rows(_lib.ptr(lat), K, L, D, int(n_loops), int(steps), _lib.ptr(rows), _lib.stream_ptr()))
    return resample(rows, size).reshape(size, L, D)


def spline_loops(y, size, n_loops):
    """Natural cubic spline through the looped key latents at `size` uniform positions (latent.py:83-92)."""
    lat = _cuda32(y, "y")
    K, Cc = lat.shape[0], lat[0].numel()
    out = torch.empty((size,) + tuple(lat.shape[1:]), device=lat.device)
    ws = torch.empty((K * n_loops + 1) * Cc, device=lat.device)
    with torch.cuda.device(lat.device):
        _lib.check(_lib.load().mb_spline_loops(_lib.ptr(lat), K, Cc, int(n_loops), int(size), _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr()))
    return out


def tempo_loops(latents, n_frames, fps, tempo, type="spline"):
    """Loop through the key latents once per bar (latent.py:95-102)."""
    n_loops = round(n_frames / fps * (tempo / 4 / 60))
    return spline_loops(latents, n_frames, n_loops) if type == "spline" else slerp_loops(latents, n_frames, n_loops)
