"""Device image helpers (bicubic resize, Lanczos-prefiltered resample, noise pyramid, perlin noise) against vectors of the
reference's own resample and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import image as OI

pytestmark = pytest.mark.gpu
G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "image.pt"))


def test_resample_against_reference_vectors(cuda):
    from maua_b200 import ops

    for name, size in {"down": (20, 36), "down_h_up_w": (30, 100), "up": (96, 80), "short_side": 24}.items():
        got = ops.resample(G["x"].to(cuda), size).cpu()
        assert got.shape == G[name].shape, name
        assert float((got - G[name]).abs().max()) < 2e-6, name


def test_noise_pyramid_and_wrapper(cuda):
    from maua_b200 import ops

    for key, size in (("pyr_8", (8, 8)), ("pyr_128", (128, 128))):
        got = ops.std_normalize_(ops.resize_bicubic(G["noise"].to(cuda), size, align_corners=False)).cpu()
        assert float((got - G[key]).abs().max()) < 1e-5, key
    x = torch.rand(1, 2, 17, 23)
    for ac in (False, True):
        ref = torch.nn.functional.interpolate(x, (40, 31), mode="bicubic", align_corners=ac)
        assert float((ops.resize_bicubic(x.to(cuda), (40, 31), align_corners=ac).cpu() - ref).abs().max()) < 2e-6


def test_perlin_noise(cuda):
    from maua_b200 import ops

    shape, res = (24, 32, 16), (3, 4, 2)
    got = ops.perlin_noise(shape, res, rng=np.random.RandomState(5)).cpu()
    rs = np.random.RandomState(5)
    theta = 2 * np.pi * rs.rand(res[0] + 1, res[1] + 1, res[2] + 1).astype(np.float32)
    phi = 2 * np.pi * rs.rand(res[0] + 1, res[1] + 1, res[2] + 1).astype(np.float32)
    ref = OI.perlin_noise(shape, res, theta, phi)
    assert got.shape == ref.shape == shape
    assert float((got - ref).abs().max()) < 1e-5
    assert float((got[0] - got[-1]).abs().max()) < 0.5      # tileable along the first axis: the ends are one step apart
    with pytest.raises(ValueError):
        ops.perlin_noise((10, 10, 10), (3, 2, 2))
