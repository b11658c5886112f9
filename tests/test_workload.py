import numpy as np
import torch

from maua_b200 import workload


def test_sweep_definition():
    y, sr = workload.sine_sweep(1.0)
    assert sr == 48000 and y.shape == (48000,) and y.dtype == np.float32
    assert abs(float(np.abs(y).max()) - 0.8) < 1e-3
    t = np.arange(48000) / 48000.0
    assert np.allclose(y, 0.8 * np.sin(2 * np.pi * (20 * t + (20000 - 20) * t * t / 2)), atol=1e-4)


def test_c2_latents_shape_and_determinism():
    a, _ = workload.c2_latents(16)
    b, _ = workload.c2_latents(16)
    assert a.shape == (720, 16, 512) and a.dtype == torch.float32 and torch.equal(a, b)
    assert torch.isfinite(a).all() and 0.2 < float(a.std()) < 1.5


def test_key_latents_follow_reference_seed_convention():
    k = workload.key_latents(16, seeds=[3])
    assert np.array_equal(k[0, 5].numpy(), np.random.RandomState(3).randn(1, 512)[0].astype(np.float32))
