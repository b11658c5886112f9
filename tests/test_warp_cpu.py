"""Host arithmetic of the StyleGAN2 feature-map warps: the product's matrix builders (GAN/wrappers/_warp.py) against the
oracle's kornia restatement (oracle/warp.py), and the oracle's warp against first principles (integer translations,
quarter turns, identity)."""
import torch

from maua_b200.GAN.wrappers import _warp
from oracle import warp as W


def test_matrices_agree_with_oracle():
    t = torch.tensor([[1.5, -2.0], [0.0, 3.0]])
    assert torch.equal(_warp.translation_matrix(t), W.translation_matrix(t))
    a, s, c = torch.tensor([10.0, -75.0]), torch.tensor([1.2, 0.8]), torch.tensor([[3.0, 4.0], [1.0, 2.0]])
    assert torch.allclose(_warp.rotation_scale_matrix(a, s, c, 16, 16), W.rotation_scale_matrix(a, s, c, 16, 16))
    assert torch.allclose(_warp.rotation_scale_matrix(a, torch.ones(1), None, 8, 12), W.rotation_scale_matrix(a, torch.ones(1), None, 8, 12))
    inv = _warp.inverse_2x3(_warp.rotation_scale_matrix(a, s, c, 16, 16))
    assert inv.shape == (2, 2, 3) and inv.dtype == torch.float32


def test_oracle_warp_first_principles():
    x = torch.arange(2 * 3 * 6 * 6, dtype=torch.float32).reshape(2, 3, 6, 6)
    assert torch.allclose(W.translate(x, torch.zeros(2, 2)), x)
    y = W.translate(x, torch.tensor([[2.0, 0.0], [0.0, 1.0]]))
    assert torch.allclose(y[0, :, :, 2:], x[0, :, :, :-2])            # shifted right by two pixels
    assert torch.allclose(y[0, :, :, 0], x[0, :, :, 2])               # reflection about the centre of the border pixel
    assert torch.allclose(y[1, :, 1:, :], x[1, :, :-1, :])            # shifted down by one pixel
    r = W.rotate(x, torch.tensor([90.0, 180.0]))
    assert torch.allclose(r[0], torch.rot90(x[0], 1, (1, 2)), atol=1e-4)   # positive angle = anti-clockwise
    assert torch.allclose(r[1], torch.rot90(x[1], 2, (1, 2)), atol=1e-4)
    assert torch.allclose(W.scale(x, torch.ones(2)), x, atol=1e-5)
