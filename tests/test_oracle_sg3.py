"""CPU checks of the StyleGAN3 oracle restatement (oracle/sg3.py; PARITY UNPINNED, see its header):
internal identities of the reference-semantics ops and the committed golden vectors."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import sg3 as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sg3_tiny.pt")


def test_filtered_lrelu_identity_filters_is_bias_act():
    torch.manual_seed(0)
    x = torch.randn(2, 3, 9, 11)
    b = torch.randn(3)
    y = O.filtered_lrelu_ref(x, None, None, b, up=1, down=1, padding=[0, 0, 0, 0], gain=np.sqrt(2), slope=0.2, clamp=1.5)
    want = (F.leaky_relu(x + b[None, :, None, None], 0.2) * np.sqrt(2)).clamp(-1.5, 1.5)
    assert torch.allclose(y, want, atol=1e-6)


def test_upfirdn2d_shapes_and_dc_gain():
    f = O.design_lowpass_filter(12, 2.0, 4.0, 16.0)
    assert abs(float(f.sum()) - 1.0) < 1e-6
    x = torch.ones(1, 1, 20, 20)
    y = O.upfirdn2d_ref(x, f, up=2, padding=[9, 8, 9, 8], gain=4)
    assert y.shape == (1, 1, 20 * 2 + 17 - 11, 20 * 2 + 17 - 11)
    assert torch.allclose(y[0, 0, 12:-12, 12:-12], torch.ones(1), atol=1e-5)  # unit DC gain after up^2 compensation


def test_modulated_conv_equals_per_sample_conv():
    torch.manual_seed(1)
    x, w, s = torch.randn(2, 5, 7, 6), torch.randn(4, 5, 3, 3), torch.randn(2, 5)
    y = O.modulated_conv2d_ref(x, w, s, demodulate=True, padding=2, input_gain=torch.tensor(0.5))
    wn = w * w.square().mean([1, 2, 3], keepdim=True).rsqrt()
    for b in range(2):
        sb = s[b] * s[b].square().mean().rsqrt()
        wb = wn * sb[None, :, None, None]
        wb = wb * (wb.square().sum([1, 2, 3], keepdim=True) + 1e-8).rsqrt() * 0.5
        assert torch.allclose(y[b], F.conv2d(x[b:b + 1], wb, padding=2)[0], atol=1e-5)


def test_golden_vectors():
    gold = torch.load(GOLD)
    cases = {"T64": dict(config="T", img_resolution=64, channel_base=1024, channel_max=32),
             "R64": dict(config="R", img_resolution=64, channel_base=2048, channel_max=48)}
    for name, kw in cases.items():
        kw = dict(kw)
        net = O.make_synthesis(kw.pop("config"), seed=3, **kw)
        img, acts = net(gold[name]["ws"], return_activations=True)
        assert [tuple(a.shape) for a in acts] == gold[name]["act_shapes"]
        assert torch.allclose(img, gold[name]["img"].float(), atol=2e-3, rtol=2e-3)
        rms = torch.tensor([a.square().mean().sqrt() for a in acts])
        assert torch.allclose(rms, gold[name]["act_rms"], rtol=1e-3)


def test_synthesis_input_is_translation_equivariant_in_phase():
    """Shifting the user transform translates the Fourier features: the input layer only sees freqs/phases."""
    torch.manual_seed(2)
    inp = O.SynthesisInput(w_dim=512, channels=16, size=36, sampling_rate=16, bandwidth=2)
    w = torch.randn(1, 512)
    a = inp(w)
    inp.transform[0, 2] = 1.0 / 16 * 4  # 4 pixels at sampling rate 16
    b = inp(w)
    assert torch.allclose(a[..., :, 4:], b[..., :, :-4], atol=2e-4)
