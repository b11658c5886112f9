"""Noise sequencers on the device: mirror of maua/audiovisual/audioreactive/selfsupervised/noise.py:11-86.

Same class names, constructor arguments and the lazy ``forward(i, b) -> [b, H, W]`` protocol (one batch of per-frame
noise maps at a time).  The seeded Gaussian fields come from ``torch.randn(generator=rng)`` exactly as in the reference;
every per-pixel operation runs in the library's kernels (csrc/sequencers.cu).  CUDA tensors only.
"""
from __future__ import annotations

import math

import torch

from ... import _lib


def _dev(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"maua_b200 noise sequencers: '{name}' must live on a CUDA device (no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


class Noise(torch.nn.Module):
    def __init__(self, length, size):
        super().__init__()
        self.length = length
        self.size = size

    def _out(self, b, like):
        return torch.empty(b, self.size[0], self.size[1], device=like.device)


class _Mix(Noise):
    two_sided = 0

    def __init__(self, rng, length, size, modulator, noise=None):
        super().__init__(length, size)
        shape = ((2,) if self.two_sided else ()) + (modulator.shape[1], size[0], size[1])
        if noise is None:
            noise = torch.randn(shape, generator=rng, device=rng.device)
        self.register_buffer("noise", _dev(noise.to(modulator.device), "noise"))
        self.register_buffer("modulator", _dev(modulator, "modulator"))

    def forward(self, i, b):
        mod = self.modulator[i: i + b]
        mod = mod.reshape(len(mod), -1).contiguous()
        out = self._out(len(mod), mod)
        with torch.cuda.device(mod.device):
            _lib.check(_lib.load().mb_noise_mix(_lib.ptr(self.noise), _lib.ptr(mod), len(mod), mod.shape[1], self.size[0] * self.size[1],
                                                self.two_sided, _lib.ptr(out), _lib.stream_ptr()))
        return out


class Blend(_Mix):
    """noise.py:11-25."""
    two_sided = 1


class Multiply(_Mix):
    """noise.py:28-40."""
    two_sided = 0


class Loop(Noise):
    """noise.py:43-54."""

    def __init__(self, rng, length, size, n_loops=1, sigma=5, noise=None, device=None):
        super().__init__(length, size)
        self.sigma = sigma
        if noise is None:
            noise = torch.randn((3, size[0], size[1]), generator=rng, device=rng.device)
        device = device or (noise.device if noise.is_cuda else "cuda")
        self.register_buffer("noise", _dev(noise.to(device), "noise"))
        self.register_buffer("idx", torch.linspace(0, n_loops * 2 * math.pi, length).to(device))

    def forward(self, i, b):
        idx = self.idx[i: i + b].contiguous()
        out = self._out(len(idx), idx)
        with torch.cuda.device(idx.device):
            _lib.check(_lib.load().mb_noise_loop(_lib.ptr(idx), _lib.ptr(self.noise), len(idx), self.size[0] * self.size[1], float(self.sigma),
                                                 _lib.ptr(out), _lib.stream_ptr()))
        return out


def _combine(x, y, mx, my, one_minus, a, c, bias):
    out = torch.empty_like(x)
    B, P = x.shape[0], x[0].numel()
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_noise_combine(_lib.ptr(x), _lib.ptr(y), _lib.ptr(mx), _lib.ptr(my), int(one_minus), float(a), float(c),
                                                float(bias), B, P, _lib.ptr(out), _lib.stream_ptr()))
    return out


class Average(Noise):
    """noise.py:57-64."""

    def __init__(self, left, right):
        super().__init__(left.length, left.size)
        self.left, self.right = left, right

    def forward(self, i, b):
        return _combine(self.left(i, b), self.right(i, b), None, None, 0, 0.5, 0.5, 0.0)


class Modulate(Noise):
    """noise.py:67-76."""

    def __init__(self, left, right, modulator):
        super().__init__(left.length, left.size)
        self.left, self.right = left, right
        self.register_buffer("modulator", _dev(modulator, "modulator").mean(1).contiguous())

    def forward(self, i, b):
        mod = self.modulator[i: i + b].contiguous()
        return _combine(self.left(i, b), self.right(i, b), mod, mod, 1, 1.0, 1.0, 0.0)


class ScaleBias(Noise):
    """noise.py:79-86."""

    def __init__(self, base, scale, bias):
        super().__init__(base.length, base.size)
        self.base, self.scale, self.bias = base, scale, bias

    def forward(self, i, b):
        return _combine(self.base(i, b), None, None, None, 0, self.scale, 0.0, self.bias)
