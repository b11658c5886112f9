"""Host side of the output-size hooks (maua/GAN/wrappers/stylegan3.py:62-117): the size arithmetic of the library
(mb_sg3_resized_output, no device needed) must predict the shape the hooked oracle network produces, and
install_hook must pick the module / size / padding the reference's get_hook would."""
import numpy as np
import pytest
import torch
from torch.nn.functional import interpolate, pad

from oracle import sg3 as O

KW = dict(channel_base=1024, channel_max=16)  # tiny channel counts: the oracle runs in well under a second


def reference_hook(G_synth, layer, size, strategy):
    size = np.flip(size)
    if strategy == "stretch":
        return lambda module, input, output: interpolate(output, tuple(int(s) for s in size), mode="bicubic", align_corners=False)
    original_size = getattr(G_synth, G_synth.layer_names[max(layer - 1, 0)]).out_size
    pad_h, pad_w = (size - original_size).astype(int) // 2
    return lambda module, input, output: pad(output, (int(pad_w), int(pad_w), int(pad_h), int(pad_h)), mode="constant", value=0)


class FakeSynth:
    """Records set_resize calls; carries the attributes install_hook reads."""

    def __init__(self, onet):
        self.layer_names, self.num_layers = onet.layer_names, onet.num_layers
        for n in onet.layer_names:
            setattr(self, n, getattr(onet, n))
        self.calls = []

    def set_resize(self, module, strategy=None, a=0, b=0):
        self.calls.append((module, strategy, a, b))


@pytest.mark.parametrize("layer,output_size,strategy", [
    (0, (320, 192), "stretch"), (5, (288, 240), "stretch"), (9, (200, 260), "stretch"), (14, (300, 256), "stretch"),
    (3, (384, 256), "pad-zero"), (0, (192, 320), "pad-zero"), (14, (280, 300), "pad-zero"), (7, (250, 254), "pad-zero"),
])
def test_predicted_output_shape_matches_hooked_oracle(layer, output_size, strategy):
    from maua_b200.GAN.networks import stylegan3 as N
    from maua_b200.GAN.wrappers.stylegan3 import install_hook, layer_multipliers

    onet = O.make_synthesis("T", img_resolution=256, seed=0, **KW)
    size = np.round(np.array(output_size) / layer_multipliers[256][layer] + 20).astype(int)
    module = getattr(onet, "input" if layer == 0 else onet.layer_names[layer - 1])
    module.register_forward_hook(reference_hook(onet, layer, size, strategy))
    want = tuple(onet(torch.randn(1, onet.num_ws, 512)).shape[2:])

    fake = FakeSynth(onet)
    install_hook(fake, layer, size, strategy)
    (mod, strat, a, b), = fake.calls
    assert mod == layer and strat == strategy
    net = N.SynthesisNetwork(w_dim=512, img_resolution=256, img_channels=3, **KW)
    net._resize = (mod, {"stretch": 1, "pad-zero": 2}[strat], a, b)   # what set_resize stores; no handle, no device
    assert net.output_hw() == want
    if strategy == "stretch" and layer < 13:
        # layers that still carry the 10-pixel margin (the reference's "+ 20", :68) land exactly on the requested size
        assert want == (output_size[1], output_size[0])


def test_native_size_without_hook():
    from maua_b200.GAN.networks import stylegan3 as N

    assert N.SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3).output_hw() == (1024, 1024)
    assert N.SynthesisNetwork(w_dim=512, img_resolution=256, img_channels=3, **KW).output_hw() == (256, 256)


def test_install_hook_rejects_unknown_strategy_and_image_layer():
    from maua_b200.GAN.wrappers.stylegan3 import install_hook

    onet = O.make_synthesis("T", img_resolution=256, seed=0, **KW)
    fake = FakeSynth(onet)
    with pytest.raises(Exception, match="Resize strategy not found"):
        install_hook(fake, 2, np.array([40, 40]), "mirror")
    with pytest.raises(NotImplementedError):
        install_hook(fake, 15, np.array([300, 300]), "stretch")
