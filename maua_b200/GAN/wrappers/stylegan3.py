"""StyleGAN3 wrapper: mirrors maua/GAN/wrappers/stylegan3.py:15-132 on top of the sm_100a network.

Kept verbatim in behaviour: ``layer_multipliers``, ctor arguments, ``forward(latents, translation,
rotation)`` including the stabilisation trick (:54-57) and ``make_transform_mat`` (:82-93, note the
reference's matrix has a zero bottom-right entry, so it always takes the pseudo-inverse path).
Arbitrary output sizes (:62-117): the reference registers a torch forward hook that resizes one module's output; here
the same arithmetic decides module and size and the resize runs as a kernel inside ``mb_net_forward``
(``SynthesisNetwork.set_resize``); ``_hook_handles`` keeps a removable handle so ``refresh_model_hooks`` works unchanged.
``forward`` additionally accepts a whole batch of translations / rotations (one transform per frame), which the
reference's ``make_transform_mat`` cannot express (it squeezes its arguments).
"""
import warnings
from typing import Optional, Tuple

import numpy as np
import torch

from ..networks import stylegan3
from .stylegan import StyleGAN, StyleGANMapper, StyleGANSynthesizer, load_network

layer_multipliers = {
    1024: {0: 64, 1: 64, 2: 64, 3: 32, 4: 32, 5: 16, 6: 8, 7: 8, 8: 4, 9: 4, 10: 2, 11: 1, 12: 1, 13: 1, 14: 1, 15: 1},
    512: {0: 32, 1: 32, 2: 32, 3: 16, 4: 16, 5: 8, 6: 8, 7: 4, 8: 4, 9: 2, 10: 2, 11: 1, 12: 1, 13: 1, 14: 1},
    256: {0: 16, 1: 16, 2: 16, 3: 16, 4: 8, 5: 8, 6: 4, 7: 4, 8: 2, 9: 2, 10: 2, 11: 1, 12: 1, 13: 1, 14: 1},
}


class StyleGAN3Mapper(StyleGANMapper):
    MapperClsFn = lambda inference: stylegan3.MappingNetwork


class StyleGAN3Synthesizer(StyleGANSynthesizer):
    def __init__(
        self, model_file: str, inference: bool, output_size: Optional[Tuple[int, int]], strategy: str, layer: int
    ) -> None:
        super().__init__()
        if model_file is None or model_file == "None":
            self.G_synth = stylegan3.SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3)
        else:
            self.G_synth = load_network(model_file).synthesis
        if output_size is None:
            output_size = (self.G_synth.img_resolution, self.G_synth.img_resolution)
        self.w_dim, self.num_ws = self.G_synth.w_dim, self.G_synth.num_ws
        self.modulation_targets = {
            "latent_w": (self.w_dim,),
            "latent_w_plus": (self.num_ws, self.w_dim),
            "translation": (2,),
            "rotation": (1,),
        }
        self.change_output_resolution(output_size, strategy, layer)

    def forward(
        self, latents: torch.Tensor = None, translation: torch.Tensor = None, rotation: torch.Tensor = None, out_fmt: str = "f32"
    ) -> torch.Tensor:
        # out_fmt is an extension of the reference signature: "f32_01" fuses render()'s (x+1)/2 .clamp(0,1) into the
        # last kernel, "u8" also the uint8 conversion of tensor2bytes; the default is the reference's raw ~[-1,1] output
        # a [B,2] translation with a tensor rotation is one transform per frame, for B == 1 too (the last batch of a render
        # whose length is not a multiple of the batch size; MemMap's batch size of one)
        batched = torch.is_tensor(translation) and torch.is_tensor(rotation) and translation.ndim == 2
        if batched:
            # (not expressible in the reference for B > 1: make_transform_mat squeezes to one matrix)
            return self.G_synth.forward(latents, out_fmt=out_fmt, transforms=make_transform_mats(translation, rotation))
        scalar_zero = lambda v: v is not None and (not torch.is_tensor(v) or v.numel() == 1) and float(v) == 0.0  # noqa: E731
        if scalar_zero(translation) and scalar_zero(rotation):
            # stabilization trick by @RiversHaveWings and @nshepperd1
            with torch.no_grad():
                bias = self.G_synth.input.affine.bias
                bias.add_(self.avg_shift.to(device=bias.device, dtype=bias.dtype))
                self.G_synth.input.affine.weight.zero_()
        elif not (translation is None or rotation is None):
            self.G_synth.input.transform.copy_(make_transform_mat(translation, rotation))
        return self.G_synth.forward(latents, out_fmt=out_fmt)

    def change_output_resolution(self, output_size: Tuple[int, int], strategy: str, layer: int):
        """Re-target the output size (W, H): drops the installed hook, and unless the size is the native one installs a new
        hook on module `layer` (stylegan3.py:62-79; size arithmetic in hooked_feature_size)."""
        self.refresh_model_hooks()
        native = (self.G_synth.img_resolution, self.G_synth.img_resolution)
        if tuple(output_size) != native:
            size = hooked_feature_size(self.G_synth.img_resolution, output_size, layer)
            self._hook_handles.append(install_hook(self.G_synth, layer, size, strategy))
        self.output_size = output_size


def hooked_feature_size(img_resolution, output_size, layer):
    """(W, H) of the feature map module `layer` must be resized to so that the image comes out at `output_size`: the
    module's up-sampling factor to the output divides the size, the 10-pixel margin on each side is added back, and sizes
    that do not divide evenly are rounded with the reference's warning (stylegan3.py:66-73)."""
    factor = layer_multipliers[img_resolution][layer]
    exact = np.array(output_size) / factor + 20
    size = np.round(exact).astype(int)
    if sum(abs(exact - size)) > 1e-10:
        warnings.warn(f"Layer {layer} resizes to multiples of {factor}. --output-size rounded to {factor * (size - 20)}")
    return size


class _ResizeHandle:
    """Stands in for the torch RemovableHandle the reference keeps in ``_hook_handles`` (wrappers/__init__.py:34-37)."""

    def __init__(self, G_synth):
        self.G_synth = G_synth

    def remove(self):
        self.G_synth.set_resize(0, None)


def install_hook(G_synth, layer, size, strategy):
    """get_hook of the reference (stylegan3.py:96-117) as a device-side resize of module `layer`'s output
    (0 = input, i = layer_names[i-1], exactly the module the reference hooks at :77)."""
    size = np.flip(size)  # W,H --> H,W
    if layer > G_synth.num_layers:
        raise NotImplementedError(
            "resizing the finished image (layer > num_layers) is not a network hook here: resample the frames instead "
            "(maua_b200.ops.resample / MauaPatch.force_output_size)"
        )
    if strategy == "stretch":
        G_synth.set_resize(layer, "stretch", int(size[0]), int(size[1]))
    elif strategy == "pad-zero":
        original_size = getattr(G_synth, G_synth.layer_names[max(layer - 1, 0)]).out_size
        pad_h, pad_w = (size - original_size).astype(int) // 2
        G_synth.set_resize(layer, "pad-zero", int(pad_h), int(pad_w))
    else:
        raise Exception(f"Resize strategy not found: {strategy}")
    return _ResizeHandle(G_synth)


def make_transform_mats(translate: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    """make_transform_mat for every row of translate [B,2] / angle [B] or [B,1]: float32 [B,3,3]."""
    t = translate.detach().cpu().double().reshape(-1, 2)
    a = angle.detach().cpu().double().reshape(-1)
    mats = [make_transform_mat(t[i], a[i]) for i in range(t.shape[0])]
    return torch.stack(mats).to(torch.float32)


def make_transform_mat(translate: Tuple[float, float], angle: float) -> torch.Tensor:
    """User transform of the Fourier-feature input from a translation (x, y) and an angle in degrees (stylegan3.py:82-93):
    the inverse of [[cos, sin, tx], [-sin, cos, ty], [0, 0, 0]].  That matrix has a zero last row, so the inverse never
    exists and the reference always lands in its pseudo-inverse branch (with its warning); both branches are kept."""
    turn = angle.squeeze().cpu() / 360.0 * np.pi * 2
    sin, cos = np.sin(turn), np.cos(turn)
    shift = translate.squeeze().cpu()
    forward = np.array([[cos, sin, shift[0]], [-sin, cos, shift[1]], [0, 0, 0]])
    try:
        inverse = np.linalg.inv(forward)
    except np.linalg.LinAlgError:
        warnings.warn(
            "Singular transform matrix, continuing with pseudo-inverse of transform matrix which might not give expected "
            "results! (If you want no translation or rotation, set them to None rather than 0)"
        )
        inverse = np.linalg.pinv(forward)
    return torch.from_numpy(inverse)


class StyleGAN3(StyleGAN):
    SynthesizerCls = StyleGAN3Synthesizer
    MapperCls = StyleGAN3Mapper

    def __init__(self, **kwargs) -> None:
        super().__init__(**kwargs)
        self.synthesizer.register_buffer(
            "avg_shift", self.synthesizer.G_synth.input.affine(self.mapper.G_map.w_avg.unsqueeze(0)).squeeze(0).detach()
        )

    def forward(self, z, c=None, truncation=1, translation=None, rotation=None):
        w = self.mapper(z, c, truncation=truncation)
        return self.synthesizer(w, translation=translation, rotation=rotation)
