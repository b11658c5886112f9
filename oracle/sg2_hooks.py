"""Oracle: the output-size hooks of the reference's StyleGAN2 wrapper on the oracle network.  TEST INFRASTRUCTURE ONLY.

Restates maua/GAN/wrappers/stylegan2.py:104-151 (change_output_resolution) and :216-340 (get_hook) as forward hooks on
oracle/sg2.py's torch modules.  Two things are passed in instead of drawn from torch's global RNG, so that the device path
can be compared value for value: the feature noise map (the reference draws torch.normal(mean_c, std_c) per channel on the
hook's first call, :236-249) and the noise_const maps of the layers behind the hook (fresh torch.randn, :137-147).
PARITY: the reference's get_hook allocates its noise with .cuda() and cannot run in a CPU-only container; this file
restates it line for line (resize / pad arithmetic, hook kinds and their order) and is checked against torch's own
interpolate / pad semantics only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def layer_names(net):
    return [f"bs.{c // 2}.conv{1 if r == 4 else c % 2}" for c, r in enumerate(sorted(net.block_resolutions * 2))]


def padding_of(strategy, layer_size, target_hw):
    """(left, right, top, bottom), mode, value of a "pad-<how>-<where>" strategy (:261-283).  Trailing pads are the remainder
    (the reference rounds pad / 2 half-to-even on both sides, which loses a pixel for pads = 1 mod 4)."""
    _, how, where = strategy.split("-")
    pad_h, pad_w = int(target_hw[0] - layer_size), int(target_hw[1] - layer_size)
    left = {"out": pad_w // 2, "left": pad_w, "right": 0, "top": pad_w // 2, "bottom": pad_w // 2}[where]
    top = {"out": pad_h // 2, "left": pad_h // 2, "right": pad_h // 2, "top": pad_h, "bottom": 0}[where]
    pad = (left, pad_w - left, top, pad_h - top)
    if how in ("reflect", "replicate", "circular"):
        return pad, how, 0.0
    return pad, "constant", float(how)


def get_hook(layer_size, target_hw, strategy, noise, pre=False):
    """feat_hook, img_hook, rgb_hook of get_hook (:216-340); `noise` [1, C, h, w] or None is the map the reference would draw."""
    target_hw = tuple(int(t) for t in target_hw)
    if strategy == "stretch":
        def resize(x, feat=False):
            x = F.interpolate(x, target_hw, mode="bicubic", align_corners=False)
            return x + noise.to(x) if (feat and noise is not None) else x

        def inverse(x):
            return F.interpolate(x, (layer_size, layer_size), mode="bicubic", align_corners=False)
    elif strategy.startswith("pad"):
        pad, how, value = padding_of(strategy, layer_size, target_hw)

        def resize(x, feat=False):
            x = F.pad(x, pad, mode=how, value=value) if how == "constant" else F.pad(x, pad, mode=how)
            return x + noise.to(x) if (feat and noise is not None) else x

        def inverse(x):
            return x[..., pad[2]: x.shape[-2] - pad[3], pad[0]: x.shape[-1] - pad[1]]
    else:
        raise Exception(f"Resize strategy not found: {strategy}")

    if pre:
        def feat_hook(module, inputs):
            return (resize(inputs[0], feat=True), *inputs[1:])
    else:
        def feat_hook(module, inputs, output):
            return resize(output, feat=True)

    def img_hook(module, inputs, output):
        return (output[0], resize(output[1], feat=False))

    def rgb_hook(module, inputs, output):
        return inverse(output)

    return feat_hook, img_hook, rgb_hook


def install(net, layer, output_size, strategy, feat_noise, later_noise):
    """change_output_resolution (:104-151) on an oracle SynthesisNetwork.  output_size = (W, H); feat_noise [C, h, w] or None;
    later_noise: {layer name: [h, w] map} for the layers behind the hook.  Returns the hook handles."""
    names = layer_names(net)
    _, block, conv = names[layer].split(".")
    blk = net.bs[int(block)]
    synth_layer = getattr(blk, conv)
    layer_size = synth_layer.resolution
    lay_mult = net.img_resolution // layer_size
    target = np.round(np.array(output_size) / lay_mult).astype(int)
    target_hw = (int(target[1]), int(target[0]))
    pre = layer == 0
    noise = None if feat_noise is None else feat_noise[None].float().cpu()
    feat_hook, img_hook, rgb_hook = get_hook(layer_size, target_hw, strategy, noise, pre=pre)
    handles = [synth_layer.register_forward_pre_hook(feat_hook) if pre else synth_layer.register_forward_hook(feat_hook)]
    if not pre:
        handles.append(blk.register_forward_hook(img_hook))
        handles.append(blk.torgb.register_forward_hook(rgb_hook))
    for name, nz in later_noise.items():
        _, b, c = name.split(".")
        getattr(net.bs[int(b)], c).noise_const = nz.float().cpu()
    return handles
