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
