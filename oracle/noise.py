"""Oracle: noise sequencers of the reference, fp32 on CPU.  TEST INFRASTRUCTURE ONLY.

Restates maua/audiovisual/audioreactive/selfsupervised/noise.py:11-86 as plain functions of the module buffers.
PINNED: tests/golden/make_audio_golden.py instantiates the reference's own classes (the file is pure torch and imports
unmodified) and requires torch.equal against these functions before writing the fixtures.
"""
import torch


def blend(noise, modulator, i, b):
    """Blend.forward, noise.py:20-25 (noise [2, M, H, W])."""
    mod = modulator[i: i + b]
    mod = mod.reshape(len(mod), -1)
    return torch.einsum("MHW,BM->BHW", noise[0], mod) + torch.einsum("MHW,BM->BHW", noise[1], 1 - mod)


def multiply(noise, modulator, i, b):
    """Multiply.forward, noise.py:36-40 (noise [M, H, W])."""
    mod = modulator[i: i + b]
    return torch.einsum("MHW,BM->BHW", noise, mod.reshape(len(mod), -1))


def loop(noise, idx, sigma, i, b):
    """Loop.forward, noise.py:50-54 (noise [3, H, W], idx = linspace(0, n_loops * 2 pi, length))."""
    freqs = torch.cos(idx[i: i + b, None, None] + noise[[0]]).div(sigma / 50)
    out = torch.sin(freqs + noise[[1]]) * noise[[2]]
    return out / (out.square().mean(dim=(1, 2), keepdim=True).sqrt() + torch.finfo(out.dtype).eps)


def average(left, right):
    """Average.forward, noise.py:63-64."""
    return (left + right) / 2


def modulate(left, right, modulator_mean, i, b):
    """Modulate.forward, noise.py:74-76 (modulator_mean = modulator.mean(1))."""
    mod = modulator_mean[i: i + b, None, None]
    return left * mod + right * (1 - mod)


def scale_bias(base, scale, bias):
    """ScaleBias.forward, noise.py:85-86."""
    return scale * base + bias
