"""CPU: the StyleGAN2 oracle against golden vectors produced by the REFERENCE's own in-tree network
(tests/golden/make_sg2_golden.py imports maua/GAN/wrappers/inference/{ops,stylegan2}.py), and the host-side module
surface (state-dict keys, num_ws, layer names) against the same reference state dict."""
import os

import torch

from oracle import sg2 as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sg2.pt")


def test_oracle_reproduces_reference_images():
    gold = torch.load(GOLD)
    g = gold["sg2_32"]
    net = O.make_synthesis(32, seed=0, **g["kw"])
    net.load_state_dict(g["state"], strict=True)
    img = net(g["ws"])
    assert img.shape == g["img"].shape == (2, 3, 32, 32)
    assert float((img - g["img"]).abs().max()) <= 1e-4 * float(g["img"].abs().max())


def test_oracle_seeded_init_matches_reference_init_order():
    """Same seed => same parameters as the reference constructor (64^2 golden stores only the image)."""
    gold = torch.load(GOLD)
    g = gold["sg2_64"]
    net = O.make_synthesis(64, seed=g["seed"], **g["kw"])
    img = net(g["ws"])
    assert float((img - g["img"]).abs().max()) <= 1e-4 * float(g["img"].abs().max())


def test_host_modules_accept_reference_state_dict():
    from maua_b200.GAN.networks import stylegan2 as N

    gold = torch.load(GOLD)
    g = gold["sg2_32"]
    net = N.SynthesisNetwork(w_dim=512, img_resolution=32, img_channels=3, **g["kw"])
    net.load_state_dict(g["state"], strict=True)
    assert net.num_ws == g["ws"].shape[1] == 2 * 4 - 1 + 1
    assert net.block_resolutions == [4, 8, 16, 32]
    assert net.bs[0].conv0 is None and net.bs[1].conv0.up == 2


def test_wrapper_surface():
    from maua_b200.GAN.wrappers import get_generator_class
    from maua_b200.GAN.wrappers.stylegan2 import StyleGAN2Synthesizer

    assert get_generator_class("stylegan2").SynthesizerCls is StyleGAN2Synthesizer
    torch.manual_seed(0)
    S = StyleGAN2Synthesizer.__new__(StyleGAN2Synthesizer)
    torch.nn.Module.__init__(S)
    from maua_b200.GAN.networks import stylegan2 as N
    S.G_synth = N.SynthesisNetwork(w_dim=512, img_resolution=32, img_channels=3, channel_base=512, channel_max=32)
    names = [f"bs.{c//2}.conv{1 if b == 4 else c % 2}" for c, b in enumerate(sorted(S.G_synth.block_resolutions * 2))]
    assert names[:4] == ["bs.0.conv1", "bs.0.conv1", "bs.1.conv0", "bs.1.conv1"]
