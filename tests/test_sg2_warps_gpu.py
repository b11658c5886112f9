"""Network-bending feature-map warps of the StyleGAN2 wrapper (maua/GAN/wrappers/stylegan2.py:65-80 and
:153-194): translation / zoom / rotation forward hooks on layer_names[layer].  The oracle side registers the same hooks
on the oracle's torch modules with oracle/warp.py standing in for kornia (absent, un-pinned: parity unpinned for the
warp arithmetic itself; affine_grid + grid_sample are torch's).  Tolerance: relative 5e-3 of the image range, as the
other StyleGAN2 network tests."""
from collections import OrderedDict

import pytest
import torch

from oracle import sg2 as O
from oracle import warp as W

pytestmark = pytest.mark.gpu


def make_pair(res, seed=0, **kw):
    from maua_b200.GAN.networks import stylegan2 as N

    onet = O.make_synthesis(res, seed=seed, **kw)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=res, img_channels=3, **kw)
    net.load_state_dict(onet.state_dict(), strict=True)
    return onet, net


def make_wrapper(net):
    from maua_b200.GAN.wrappers.stylegan2 import StyleGAN2Synthesizer

    S = StyleGAN2Synthesizer.__new__(StyleGAN2Synthesizer)
    torch.nn.Module.__init__(S)
    S.G_synth = net
    S.layer_names = [f"bs.{c//2}.conv{1 if bs == 4 else c % 2}" for c, bs in enumerate(sorted(net.block_resolutions * 2))]
    S.translate_hook, S.rotate_hook, S.zoom_hook = None, None, None
    S._warp_hooks = OrderedDict()
    return S


def oracle_layer(onet, names, layer):
    _, block, conv = names[layer].split(".")
    return getattr(onet.bs[int(block)], conv)


def rel_err(a, b):
    return float((a.float().cpu() - b).abs().max() / b.abs().max())


def test_translation_zoom_rotation_match_hooked_oracle(cuda):
    onet, net = make_pair(64, channel_base=4096, channel_max=128)
    S = make_wrapper(net)
    B = 3
    torch.manual_seed(11)
    ws = torch.randn(B, net.num_ws, 512)
    translation = torch.tensor([[0.10, -0.05], [0.0, 0.25], [-0.4, 0.3]])
    zoom = torch.tensor([[1.3], [0.7], [1.0]])
    rotation = torch.tensor([[15.0], [-80.0], [200.0]])
    t_layer, z_layer, r_layer = 5, 5, 6

    def t_hook(module, input, output):  # stylegan2.py:157-164
        _, _, h, w = output.shape
        return W.translate(output, translation * torch.tensor([[h, w]]))

    def z_hook(module, input, output):  # stylegan2.py:187-190
        return W.scale(output, zoom.squeeze(), None)

    def r_hook(module, input, output):  # stylegan2.py:174-177
        return W.rotate(output, rotation.squeeze(), None)

    handles = [oracle_layer(onet, S.layer_names, t_layer).register_forward_hook(t_hook),
               oracle_layer(onet, S.layer_names, z_layer).register_forward_hook(z_hook),
               oracle_layer(onet, S.layer_names, r_layer).register_forward_hook(r_hook)]
    ref = onet(ws)
    out = S.forward(ws.to(cuda), translation=translation, translation_layer=t_layer, zoom=zoom, zoom_layer=z_layer,
                    rotation=rotation, rotation_layer=r_layer)
    plain = net(ws.to(cuda))
    print("warped vs oracle:", rel_err(out, ref), " warped vs unwarped:", rel_err(plain, ref))
    assert rel_err(out, ref) < 5e-3
    assert rel_err(plain, ref) > 5e-2  # the warps do change the image
    # hooks stay installed (the reference never removes them): a call without arguments renders the same frames
    assert torch.equal(S.forward(ws.to(cuda)), out)
    # re-applying one warp moves its hook to the end of the layer's hook list: zoom -> translate order on layer 5
    for h in handles:
        h.remove()
    handles = [oracle_layer(onet, S.layer_names, z_layer).register_forward_hook(z_hook),
               oracle_layer(onet, S.layer_names, r_layer).register_forward_hook(r_hook),
               oracle_layer(onet, S.layer_names, t_layer).register_forward_hook(t_hook)]
    ref2 = onet(ws)
    out2 = S.forward(ws.to(cuda), translation=translation, translation_layer=t_layer)
    assert rel_err(out2, ref2) < 5e-3
    S.remove_warps()
    assert torch.equal(S.forward(ws.to(cuda)), plain)


@pytest.mark.parametrize("layer", [1, 2, 9])
def test_single_warp_on_first_const_and_last_layer(cuda, layer):
    """layer 1 = bs.0.conv1 (the 4x4 const block), 2 = a conv0 (up-sampling conv + FIR), 9 = the last conv1 (its ToRGB
    sees the warped map)."""
    onet, net = make_pair(64, channel_base=4096, channel_max=128)
    S = make_wrapper(net)
    torch.manual_seed(5)
    ws = torch.randn(2, net.num_ws, 512)
    rotation = torch.tensor([[33.0], [-120.0]])
    centre = torch.tensor([[1.0, 2.0], [1.5, 0.5]])

    def r_hook(module, input, output):
        return W.rotate(output, rotation.squeeze(), centre)

    oracle_layer(onet, S.layer_names, layer).register_forward_hook(r_hook)
    ref = onet(ws)
    out = S.forward(ws.to(cuda), rotation=rotation, rotation_layer=layer, rotation_center=centre)
    assert rel_err(out, ref) < 5e-3
    u8 = S.forward(ws.to(cuda), out_fmt="u8").cpu()
    want8 = (((ref + 1) / 2).clamp(0, 1) * 255).round().permute(0, 2, 3, 1)
    assert float((u8.float() - want8).abs().max()) <= 2
