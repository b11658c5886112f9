"""StyleGAN2 wrapper: mirrors maua/GAN/wrappers/stylegan2.py:21-213 on top of the sm_100a network (csrc/sg2.cu).

Kept in behaviour: ctor arguments, ``layer_names`` (:48-51), ``modulation_targets``, ``forward(latents, ..., **noise)``
swapping each SynthesisLayer's ``noise_const`` for the per-frame map of the batch (:81-96, bicubic resize + warning on
a shape mismatch) and ``make_noise_pyramid`` (:196-213), and the network-bending feature warps (translation / zoom /
rotation, :65-80 and :153-194): the reference registers kornia forward hooks on ``layer_names[layer]`` that stay installed
until the same kind of warp is applied again; here each installed hook is a (layer, matrices) record and the warps run
as kernels between that layer's activation and its consumers (``mb_sg2_set_warps``), in the order torch would call the
hooks (re-applying a warp moves it to the end, as remove + register does).  Non-native output sizes (:100-151, which
inject ``torch.normal`` noise into the resized map) are not built.
"""
from collections import OrderedDict
import warnings
from typing import Optional, Tuple

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

    def change_output_resolution(self, output_size: Tuple[int, int], strategy: str, layer: int):
        self.refresh_model_hooks()
        if tuple(output_size) != (self.G_synth.img_resolution, self.G_synth.img_resolution):
            raise NotImplementedError(
                "non-native output sizes (feature-map resize hooks, maua/GAN/wrappers/stylegan2.py:100-151) "
                "are not built yet (SURVEY §8f N2)"
            )
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


class StyleGAN2(StyleGAN):
    SynthesizerCls = StyleGAN2Synthesizer
    MapperCls = StyleGAN2Mapper
