"""StyleGAN2 wrapper: mirrors maua/GAN/wrappers/stylegan2.py:21-213 on top of the sm_100a network (csrc/sg2.cu).

Kept in behaviour: ctor arguments, ``layer_names`` (:48-51), ``modulation_targets``, ``forward(latents, ..., **noise)``
swapping each SynthesisLayer's ``noise_const`` for the per-frame map of the batch (:81-96, bicubic resize + warning on
a shape mismatch) and ``make_noise_pyramid`` (:196-213), and the network-bending feature warps (translation / zoom /
rotation, :65-80 and :153-194): the reference registers kornia forward hooks on ``layer_names[layer]`` that stay installed
until the same kind of warp is applied again; here each installed hook is a (layer, matrices) record and the warps run
as kernels between that layer's activation and its consumers (``mb_sg2_set_warps``), in the order torch would call the
hooks (re-applying a warp moves it to the end, as remove + register does).  Non-native output sizes (:104-151): the
feature / image / ToRGB hooks of ``get_hook`` run inside the network's forward (``mb_sg2_set_resize``), their noise maps come
from a seeded generator instead of ``torch.normal`` on the global RNG.
"""
from collections import OrderedDict
import warnings
from typing import Optional, Tuple

import numpy as np
import torch

from ... import ops
from torch import Tensor

from ..networks import stylegan2
from . import _warp
from .stylegan import StyleGAN, StyleGANMapper, StyleGANSynthesizer, load_network


class StyleGAN2Mapper(StyleGANMapper):
    MapperClsFn = lambda inference: stylegan2.MappingNetwork


class StyleGAN2Synthesizer(StyleGANSynthesizer):
    __constants__ = ["w_dim", "num_ws", "layer_names"]
    _warp_hooks = None  # kind -> (layer, inverse matrices [B,2,3]); insertion order = hook order

    def __init__(
        self, model_file: str, inference: bool, output_size: Optional[Tuple[int, int]], strategy: str, layer: int
    ) -> None:
        super().__init__()
        if model_file is None or model_file == "None":
            self.G_synth = stylegan2.SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3)
        else:
            self.G_synth = load_network(model_file, inference).synthesis
        if output_size is None:
            output_size = (self.G_synth.img_resolution, self.G_synth.img_resolution)
        self.w_dim, self.num_ws = self.G_synth.w_dim, self.G_synth.num_ws
        self.layer_names = [
            f"bs.{c//2}.conv{1 if block_size == 4 else c % 2}"
            for c, block_size in enumerate(sorted(self.G_synth.block_resolutions * 2))
        ]
        self.modulation_targets = {
            "latent_w": (self.w_dim,),
            "latent_w_plus": (self.num_ws, self.w_dim),
            "translation": (2,),
            "rotation": (1,),
        }
        self.translate_hook, self.rotate_hook, self.zoom_hook = None, None, None
        self._warp_hooks = OrderedDict()
        self.change_output_resolution(output_size, strategy, layer)

    def forward(
        self,
        latents: Tensor,
        translation: Optional[Tensor] = None,
        translation_layer: int = 7,
        zoom: Optional[Tensor] = None,
        zoom_layer: int = 7,
        zoom_center: Optional[int] = None,
        rotation: Optional[Tensor] = None,
        rotation_layer: int = 7,
        rotation_center: Optional[int] = None,
        out_fmt: str = "f32",
        **noise,
    ) -> Tensor:
        if translation is not None:
            self.apply_translation(translation_layer, translation)
        if zoom is not None:
            self.apply_zoom(zoom_layer, zoom, zoom_center)
        if rotation is not None:
            self.apply_rotation(rotation_layer, rotation, rotation_center)
        if noise:
            noises, l = list(noise.values()), 0
            for block in self.G_synth.bs:
                if l >= len(noises):
                    continue
                for c in ([block.conv0] if getattr(block, "conv0", None) is not None else []) + [block.conv1]:
                    if l >= len(noises):
                        break
                    noise_l = noises[l].to(c.noise_const, non_blocking=True)
                    if (noise_l.shape[-2], noise_l.shape[-1]) != (c.noise_const.shape[-2], c.noise_const.shape[-1]):
                        warnings.warn(
                            f"Supplied noise for SynthesisLayer {l} has shape {noise_l.shape} while the expected "
                            f"shape is {c.noise_const.shape}. Resizing the supplied noise to match..."
                        )
                        h, w = c.noise_const.shape[-2], c.noise_const.shape[-1]
                        noise_l = ops.resize_bicubic(noise_l, (h, w), align_corners=False)
                    setattr(c, "noise_const", noise_l)
                    l += 1
        return self.G_synth.forward(latents, noise_mode="const", out_fmt=out_fmt, warps=list((self._warp_hooks or {}).values()))

    def _install_warp(self, kind, layer, inv_mats):
        if self._warp_hooks is None:
            self._warp_hooks = OrderedDict()
        self._warp_hooks.pop(kind, None)  # hook.remove() + register_forward_hook: the new hook runs last
        self._warp_hooks[kind] = (int(layer), inv_mats)
        return kind

    def remove_warps(self):
        """Drop every installed translation / zoom / rotation warp (the reference keeps them until overwritten)."""
        self._warp_hooks = OrderedDict()
        self.translate_hook, self.rotate_hook, self.zoom_hook = None, None, None

    def apply_translation(self, layer, translation):
        r = self.G_synth.layer_resolution(layer)
        # kT.translate(output, translation * [[h, w]], padding_mode="reflection")  (stylegan2.py:161-163)
        pixels = translation.detach().cpu().double().reshape(-1, 2) * torch.tensor([[r, r]], dtype=torch.float64)
        self.translate_hook = self._install_warp("translate", layer, _warp.inverse_2x3(_warp.translation_matrix(pixels)))

    def apply_rotation(self, layer, angle, center):
        r = self.G_synth.layer_resolution(layer)
        # kT.rotate(output, angle.squeeze(), center, padding_mode="reflection")  (stylegan2.py:175-177)
        m = _warp.rotation_scale_matrix(angle, torch.ones(1), center, r, r)
        self.rotate_hook = self._install_warp("rotate", layer, _warp.inverse_2x3(m))

    def apply_zoom(self, layer, zoom, center):
        r = self.G_synth.layer_resolution(layer)
        # kT.scale(output, zoom.squeeze(), center, padding_mode="reflection")  (stylegan2.py:188-190)
        m = _warp.rotation_scale_matrix(torch.zeros(1), zoom, center, r, r)
        self.zoom_hook = self._install_warp("zoom", layer, _warp.inverse_2x3(m))

    resize_seed = 0   # seed of the hook's noise maps and of the probe latents (the reference draws them from the global RNG)

    def change_output_resolution(self, output_size: Tuple[int, int], strategy: str, layer: int):
        """Re-target the output size (W, H) (stylegan2.py:104-151): the output of ``layer_names[layer]`` is resized to
        ``output_size / (img_resolution / layer size)`` ("stretch": bicubic; "pad-<how>-<where>": constant value / reflect /
        replicate / circular border on out / left / right / top / bottom), a noise map drawn once per channel from
        N(mean_c, std_c) of the resized features is added to it, the hooked block's ToRGB output is mapped back before it
        joins the skip image and the block's image is resized like the features (get_hook, :216-340); every later layer gets a
        fresh noise_const of its new size (:137-147).  All of it runs inside the network's forward (``mb_sg2_set_resize``).
        Noise maps and probe latents come from a generator seeded with ``resize_seed`` instead of the global RNG."""
        self.refresh_model_hooks()
        res = self.G_synth.img_resolution
        if tuple(output_size) != (res, res):
            _, block, conv = self.layer_names[layer].split(".")
            bk = int(block)
            layer_size = self.G_synth.block_resolutions[bk]
            lay_mult = res // layer_size
            unrounded_size = np.array(output_size) / lay_mult
            target_size = np.round(unrounded_size).astype(int)
            if sum(abs(unrounded_size - target_size)) > 1e-10:
                warnings.warn(f"Layer {layer} resizes to multiples of {lay_mult}. --output-size rounded to {lay_mult * target_size}")
            tw, th = int(target_size[0]), int(target_size[1])      # (W, H) -> rows, columns
            mode, pads, value = parse_resize_strategy(strategy, layer_size, (th, tw))
            self._hook_handles.append(_ResizeHandle(self, layer, bk, layer_size, mode, (th, tw), pads, value))
        self.output_size = output_size

    def make_noise_pyramid(self, noise, layer_limit=8):
        noises = {}
        for l, layer in enumerate(self.layer_names[1:]):
            if l > layer_limit:
                continue
            _, block, conv = layer.split(".")
            synth_layer = getattr(self.G_synth.bs[int(block)], conv)
            h, w = synth_layer.noise_const.shape[-2], synth_layer.noise_const.shape[-1]
            # bicubic resize and unit-std normalisation on the device (stylegan2.py:203-212), returned on the host like the reference
            noises[f"noise{l}"] = ops.std_normalize_(ops.resize_bicubic(noise.cuda(), (h, w), align_corners=False)).cpu()
        return noises


def parse_resize_strategy(strategy, layer_size, target_hw):
    """'stretch' | 'pad-<how>-<where>' -> (mode, (pad_top, pad_left), value) with the padding arithmetic of get_hook
    (stylegan2.py:261-283): how in reflect / replicate / circular or a number (constant border), where in
    out / left / right / top / bottom."""
    if strategy == "stretch":
        return "stretch", (0, 0), 0.0
    if not strategy.startswith("pad"):
        raise Exception(f"Resize strategy not found: {strategy}")
    _, how, where = strategy.split("-")
    pad_h, pad_w = int(round(target_hw[0] - layer_size)), int(round(target_hw[1] - layer_size))
    if pad_h < 0 or pad_w < 0:
        raise NotImplementedError("negative padding (an output smaller than the layer) is a TODO of the reference as well; use 'stretch'")
    left = {"out": pad_w // 2, "left": pad_w, "right": 0, "top": pad_w // 2, "bottom": pad_w // 2}
    top = {"out": pad_h // 2, "left": pad_h // 2, "right": pad_h // 2, "top": pad_h, "bottom": 0}
    if where not in left:
        raise Exception(f"Resize strategy not found: {strategy}")
    if how in ("reflect", "replicate", "circular"):
        return how, (top[where], left[where]), 0.0
    return "constant", (top[where], left[where]), float(how)


class _ResizeHandle:
    """What change_output_resolution installs (stands in for the torch RemovableHandles of the reference's four kinds of
    hooks, stylegan2.py:124-147): configures the network's output-size hook, swaps the noise maps of the later layers for
    maps of their new size, and undoes both on remove()."""

    def __init__(self, synth, layer, bk, layer_size, mode, target_hw, pads, value):
        net = synth.G_synth
        self.net = net
        dev = torch.device("cuda")
        gen = torch.Generator(device="cpu").manual_seed(int(synth.resize_seed))
        th, tw = target_hw
        # noise maps of every layer behind the hook (noise_adjust, :137-147): randn of the layer's new size
        self.saved = {}
        for l in range(layer + 1, len(synth.layer_names)):
            _, b, c = synth.layer_names[l].split(".")
            j = int(b)
            mod = getattr(net.bs[j], c)
            if mod in self.saved:
                continue
            h, w = (th, tw) if j == bk else (th << (j - bk), tw << (j - bk))
            self.saved[mod] = mod.noise_const
            mod.noise_const = torch.randn(h, w, generator=gen)
        try:
            if layer == 0:
                # forward PRE-hook on bs.0.conv1 (:122-128): the constant input itself is resized, noise included
                cst = self._resize(net.bs[0].const.detach().to(dev, torch.float32)[None], mode, target_hw, pads, value, layer_size)[0]
                noise = self._draw(cst.mean(dim=(1, 2)), cst.std(dim=(1, 2)), target_hw, gen, dev)
                net.set_resize(0, mode, target_hw, pads, value, noise=(cst + noise).contiguous())
            else:
                c = getattr(net.bs[bk], "conv0" if layer % 2 == 0 else "conv1").out_channels
                stats = torch.zeros(2, c, device=dev)
                net.set_resize(layer, mode, target_hw, pads, value, noise=None, stats=stats)
                # the reference runs one forward on random latents right away (:149), which is when its hook measures the
                # per-channel statistics and fixes the noise map
                probe = torch.randn(1, net.num_ws, net.w_dim, generator=gen).to(dev)
                net.forward(probe)
                torch.cuda.synchronize()
                noise = self._draw(stats[0], stats[1], target_hw, gen, dev)
                net.set_resize(layer, mode, target_hw, pads, value, noise=noise, stats=None)
        except Exception:
            self.remove()
            raise

    @staticmethod
    def _draw(mean, std, hw, gen, dev):
        """torch.normal(mean_c, std_c, size=(1, 1, h, w)) per channel (:238-247) -> [C, h, w]."""
        z = torch.randn(mean.numel(), hw[0], hw[1], generator=gen).to(dev)
        std = torch.nan_to_num(std.to(dev), nan=0.0)     # a 1 x 1 map has no spread (the reference's x.std() is NaN there)
        return (mean.to(dev)[:, None, None] + std[:, None, None] * z).contiguous()

    @staticmethod
    def _resize(x, mode, hw, pads, value, layer_size):
        if mode == "stretch":
            return ops.resize_bicubic(x, hw, align_corners=False)
        pt, pl = pads
        pad = (pl, hw[1] - layer_size - pl, pt, hw[0] - layer_size - pt)
        return torch.nn.functional.pad(x, pad, mode=mode, value=value) if mode == "constant" else torch.nn.functional.pad(x, pad, mode=mode)

    def remove(self):
        for mod, buf in self.saved.items():
            mod.noise_const = buf
        self.saved = {}
        self.net.set_resize(None, None, None)


class StyleGAN2(StyleGAN):
    SynthesizerCls = StyleGAN2Synthesizer
    MapperCls = StyleGAN2Mapper
