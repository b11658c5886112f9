"""Output-size hooks: maua/GAN/wrappers/stylegan3.py:62-117 resizes one module's output with a torch
forward hook; here the resize is a kernel inside mb_net_forward.  The oracle side registers the reference's hook
(restated below from :96-117) on the oracle's torch modules.  Tolerance: 1e-3 max-abs on fp32 pixels."""
import warnings

import numpy as np
import pytest
import torch
from torch.nn.functional import interpolate, pad

from oracle import sg3 as O

pytestmark = pytest.mark.gpu
PIX_TOL = 1e-3
KW = dict(channel_base=8192, channel_max=128)


def pix(x):
    return ((x.float().cpu() + 1) / 2).clamp(0, 1)


def make_pair(seed=0):
    from maua_b200.GAN.networks import stylegan3 as N

    onet = O.make_synthesis("T", img_resolution=256, seed=seed, **KW)
    torch.manual_seed(seed)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=256, img_channels=3, **KW)
    net.load_state_dict(onet.state_dict())
    return onet, net


def reference_hook(G_synth, layer, size, strategy):
    """get_hook of maua/GAN/wrappers/stylegan3.py:96-117."""
    size = np.flip(size)  # W,H --> H,W
    if strategy == "stretch":
        return lambda module, input, output: interpolate(output, tuple(int(s) for s in size), mode="bicubic", align_corners=False)
    original_size = getattr(G_synth, G_synth.layer_names[max(layer - 1, 0)]).out_size
    pad_h, pad_w = (size - original_size).astype(int) // 2
    padding = (int(pad_w), int(pad_w), int(pad_h), int(pad_h))
    return lambda module, input, output: pad(output, padding, mode="constant", value=0)


def hooked_oracle(onet, layer, size, strategy):
    module = getattr(onet, "input" if layer == 0 else onet.layer_names[layer - 1])
    return module.register_forward_hook(reference_hook(onet, layer, np.asarray(size), strategy))


@pytest.mark.parametrize("layer,output_size,strategy", [
    (0, (320, 192), "stretch"),      # input module, non-square, multiplier 16
    (5, (288, 240), "stretch"),      # mid network, multiplier 8
    (9, (200, 260), "stretch"),      # multiplier 2: shrink one axis, grow the other
    (14, (300, 256), "stretch"),     # last layer before ToRGB (planar path)
    (3, (384, 256), "pad-zero"),
    (0, (192, 320), "pad-zero"),     # negative padding (crop) on one axis
    (14, (280, 300), "pad-zero"),
])
def test_hooked_network_matches_hooked_oracle(cuda, layer, output_size, strategy):
    from maua_b200.GAN.wrappers.stylegan3 import install_hook, layer_multipliers

    onet, net = make_pair()
    mult = layer_multipliers[256][layer]
    size = np.round(np.array(output_size) / mult + 20).astype(int)  # (W, H), stylegan3.py:68-69
    handle_o = hooked_oracle(onet, layer, size, strategy)
    handle_n = install_hook(net, layer, size, strategy)
    torch.manual_seed(3)
    ws = torch.randn(2, net.num_ws, 512)
    ref = onet(ws)
    out = net(ws.to(cuda))
    assert tuple(out.shape) == tuple(ref.shape), (out.shape, ref.shape)
    assert net.output_hw() == tuple(ref.shape[2:])
    err = float((pix(out) - pix(ref)).abs().max())
    print(f"layer {layer} {strategy} -> {tuple(ref.shape[2:])}: max-abs pixel error {err:.3e}")
    assert err <= PIX_TOL, err
    u8 = net(ws.to(cuda), out_fmt="u8").cpu()
    want8 = (pix(ref) * 255).round().permute(0, 2, 3, 1)
    assert float((u8.float() - want8).abs().max()) <= 1
    # removing the hook restores the native geometry bit for bit
    handle_o.remove()
    handle_n.remove()
    base = net(ws.to(cuda))
    assert tuple(base.shape) == (2, 3, 256, 256)
    _, fresh = make_pair()
    assert torch.equal(base, fresh(ws.to(cuda)))


def test_hook_on_the_radial_config(cuda):
    """StyleGAN3-R (1x1 convs, radial down filters) with a non-square stretch: the generic filter fallbacks see H != W too."""
    from maua_b200.GAN.networks import stylegan3 as N
    from maua_b200.GAN.wrappers.stylegan3 import install_hook

    kw = dict(O.SG3_R_KWARGS, channel_base=8192, channel_max=160)
    onet = O.make_synthesis("T", img_resolution=256, seed=0, **kw)
    torch.manual_seed(0)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=256, img_channels=3, **kw)
    net.load_state_dict(onet.state_dict())
    size = np.array([44, 30])  # (W, H) at layer 4
    hooked_oracle(onet, 4, size, "stretch")
    install_hook(net, 4, size, "stretch")
    torch.manual_seed(6)
    ws = torch.randn(2, net.num_ws, 512)
    ref = onet(ws)
    out = net(ws.to(cuda))
    assert tuple(out.shape) == tuple(ref.shape)
    err = float((pix(out) - pix(ref)).abs().max())
    print(f"R config, layer 4 stretch -> {tuple(ref.shape[2:])}: max-abs pixel error {err:.3e}")
    assert err <= PIX_TOL, err


def test_wrapper_change_output_resolution(cuda):
    """StyleGAN3Synthesizer(output_size=(W, H), strategy, layer) end to end at the reference's default 1024^2 network:
    rounding warning (:70-73), output shape, hook removal through refresh_model_hooks."""
    from maua_b200.GAN.wrappers.stylegan3 import StyleGAN3Synthesizer

    torch.manual_seed(0)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        S = StyleGAN3Synthesizer(model_file=None, inference=False, output_size=(1920, 1080), strategy="stretch", layer=0)
    assert any("rounded" in str(w.message) for w in rec)  # 1080 / 64 is not an integer
    lat = torch.randn(1, 16, 512, device=cuda)
    img = S(latents=lat)
    assert tuple(img.shape) == (1, 3, 1088, 1920)
    assert torch.isfinite(img).all()
    S.change_output_resolution((1024, 1024), "stretch", 0)
    assert tuple(S(latents=lat).shape) == (1, 3, 1024, 1024)
    with pytest.raises(Exception, match="Resize strategy not found"):
        S.change_output_resolution((2048, 1024), "mirror", 0)


def test_per_frame_transforms_match_oracle(cuda):
    """One input transform per frame (mb_net_forward_xf) == the oracle run frame by frame with that transform, and
    bit-identical to the shared-buffer path (input.transform) run frame by frame, also for the rank-2 matrices the
    reference's make_transform_mat produces (pseudo-inverse path, stylegan3.py:86-92)."""
    from maua_b200.GAN.wrappers.stylegan3 import make_transform_mats

    onet, net = make_pair()
    torch.manual_seed(4)
    ws = torch.randn(3, net.num_ws, 512)

    def rigid(deg, tx, ty):
        c, s = np.cos(np.deg2rad(deg)), np.sin(np.deg2rad(deg))
        return torch.tensor([[c, s, tx], [-s, c, ty], [0.0, 0.0, 1.0]], dtype=torch.float32)

    mats = torch.stack([rigid(10.0, 0.1, 0.2), rigid(200.0, 0.3, -0.1), rigid(-45.0, 0.0, 0.05)])
    out = net(ws.to(cuda), transforms=mats.to(cuda)).clone()
    for i in range(3):
        onet.input.transform.copy_(mats[i])
        ref = onet(ws[i:i + 1])
        assert float((pix(out[i:i + 1]) - pix(ref)).abs().max()) <= PIX_TOL, i
        net.input.transform.copy_(mats[i])
        assert torch.equal(net(ws[i:i + 1].to(cuda)), out[i:i + 1]), i

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pinv = make_transform_mats(torch.tensor([[0.1, 0.2], [0.3, -0.1], [0.0, 0.05]]), torch.tensor([[10.0], [200.0], [-45.0]]))
    out = net(ws.to(cuda), transforms=pinv.to(cuda)).clone()
    for i in range(3):
        net.input.transform.copy_(pinv[i])
        assert torch.equal(net(ws[i:i + 1].to(cuda)), out[i:i + 1]), i
