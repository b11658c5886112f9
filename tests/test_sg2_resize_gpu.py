"""Output-size hooks of the StyleGAN2 wrapper (maua/GAN/wrappers/stylegan2.py:104-151, get_hook :216-340) on the device
against the same hooks on the CPU oracle network, with the noise maps the wrapper drew handed to the oracle."""
from collections import OrderedDict

import pytest
import torch

from oracle import sg2 as O
from oracle import sg2_hooks as OH

pytestmark = pytest.mark.gpu
KW = dict(channel_base=2048, channel_max=64)


def make(res=64):
    from maua_b200.GAN.networks import stylegan2 as N
    from maua_b200.GAN.wrappers.stylegan2 import StyleGAN2Synthesizer

    onet = O.make_synthesis(res, seed=2, **KW)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=res, img_channels=3, **KW)
    net.load_state_dict(onet.state_dict(), strict=True)
    S = StyleGAN2Synthesizer.__new__(StyleGAN2Synthesizer)
    torch.nn.Module.__init__(S)
    S.G_synth, S._hook_handles, S._warp_hooks = net, [], OrderedDict()
    S.w_dim, S.num_ws = net.w_dim, net.num_ws
    S.layer_names = OH.layer_names(net)
    S.output_size = (res, res)
    return onet, net, S


def pix(x):
    return ((x.float().cpu() + 1) / 2).clamp(0, 1)


CASES = [
    # layer, output size (W, H), strategy
    (5, (96, 80), "stretch"),            # conv1 of the 16^2 block -> 24 x 20 features, image hooks on that block
    (4, (96, 80), "stretch"),            # conv0 of the 16^2 block: conv1 and ToRGB run resized
    (3, (96, 80), "pad-reflect-out"),    # conv1 of the 8^2 block, 8 -> 12 x 10
    (6, (96, 72), "pad-0.5-left"),       # conv0 of the 32^2 block, constant border on the left
    (7, (80, 96), "pad-replicate-bottom"),
    (5, (80, 72), "pad-circular-top"),
    (1, (96, 80), "stretch"),            # post-hook on bs.0.conv1
    (0, (96, 80), "stretch"),            # the reference's PRE-hook: the constant input is resized
    (0, (96, 80), "pad-reflect-out"),
    (5, (48, 40), "stretch"),            # shrinking
    (0, (32, 16), "stretch"),            # the constant input shrinks to 1 x 2: what sample.generate's default downscale does
]


@pytest.mark.parametrize("layer,size,strategy", CASES)
def test_resize_hook_matches_oracle(cuda, layer, size, strategy):
    onet, net, S = make()
    S.change_output_resolution(size, strategy, layer)
    assert S.output_size == size
    torch.manual_seed(11)
    ws = torch.randn(2, net.num_ws, 512)
    out = S.forward(ws.to(cuda))
    assert tuple(out.shape) == (2, 3, size[1], size[0])
    # the oracle gets the maps the wrapper drew
    _, _, _, noise, _ = net._resize
    names = S.layer_names
    later = {names[l]: getattr(net.bs[int(names[l].split(".")[1])], names[l].split(".")[2]).noise_const for l in range(layer + 1, len(names))}
    if layer == 0:
        # the device path receives the resized constant with its noise already added: recover the noise map for the oracle
        cst = onet.bs[0].const.detach()[None]
        hooks = OH.get_hook(4, (noise.shape[1], noise.shape[2]), strategy, None, pre=True)
        resized = hooks[0](None, (cst,))[0][0]
        feat_noise = noise.cpu() - resized
    else:
        feat_noise = noise.cpu()
    OH.install(onet, layer, size, strategy, feat_noise, later)
    ref = onet(ws)
    assert ref.shape == out.shape
    err = float((pix(out) - pix(ref)).abs().max())
    rel = float((out.cpu() - ref).abs().max() / ref.abs().max())
    print(f"layer {layer} {strategy} {size}: pixel err {err:.3e} rel {rel:.3e}")
    assert err <= 1e-3
    # the noise map carries the statistics of the resized features it was drawn for (per-channel mean / std, :236-249)
    if layer > 0:
        assert noise.shape[1:] == (ref.shape[2] // (onet.img_resolution // onet.bs[layer // 2].resolution),
                                   ref.shape[3] // (onet.img_resolution // onet.bs[layer // 2].resolution))
    # rgb24 output and removal of the hook
    u8 = S.forward(ws.to(cuda), out_fmt="u8")
    assert tuple(u8.shape) == (2, size[1], size[0], 3)
    S.change_output_resolution((64, 64), "stretch", 0)
    onet2 = O.make_synthesis(64, seed=2, **KW)
    plain = S.forward(ws.to(cuda))
    assert tuple(plain.shape) == (2, 3, 64, 64)
    assert float((pix(plain) - pix(onet2(ws))).abs().max()) <= 1e-3


def test_resize_noise_statistics_and_seed(cuda):
    """The feature noise is N(mean_c, std_c) of the resized features per channel; the same seed gives the same render."""
    _, net, S = make()
    S.change_output_resolution((96, 80), "stretch", 5)
    noise = net._resize[3]
    torch.manual_seed(3)
    ws = torch.randn(1, net.num_ws, 512).to(cuda)
    a = S.forward(ws).clone()
    assert tuple(noise.shape[1:]) == (20, 24) and torch.isfinite(noise).all() and float(noise.std(dim=(1, 2)).min()) > 0
    # the stats tap measures the resized features of any forward: mean / std per channel are what the map was scaled by
    stats = torch.zeros(2, noise.shape[0], device=cuda)
    net.set_resize(5, "stretch", (20, 24), noise=None, stats=stats)
    gen = torch.Generator().manual_seed(S.resize_seed)
    for _ in range(4):                                # the handle's draws before the probe latents: four later noise maps
        torch.randn(1, generator=gen)
    net.set_resize(5, "stretch", (20, 24), noise=noise, stats=None)
    S.change_output_resolution((96, 80), "stretch", 5)
    b = S.forward(ws)
    assert torch.equal(a, b)
    S.resize_seed = 1
    S.change_output_resolution((96, 80), "stretch", 5)
    c = S.forward(ws)
    assert not torch.equal(a, c)
    # with warps on a layer behind the hook the combination still runs (non-square warp)
    S.apply_translation(7, torch.tensor([[0.1, -0.05]]))
    d = S.forward(ws)
    assert d.shape == c.shape and torch.isfinite(d).all() and not torch.equal(c, d)
