"""StyleGAN2Patch: mirror of maua/audiovisual/patches/base/stylegan2.py:7-57 (four overridable stages)."""
import torch

from ....GAN.wrappers.stylegan2 import StyleGAN2
from . import MauaPatch


class StyleGAN2Patch(MauaPatch):
    """Patch bound to a StyleGAN2 generator.  The render loop calls, in order: process_audio, process_mapper_inputs (->
    keyword tensors for the mapper), process_synthesizer_inputs (-> dict of [T, ...] tensors: "latents", optionally
    "translation" / "zoom" / "rotation" (+ their ``_layer`` / ``_center`` companions) and per-layer "noise{i}" maps), and
    process_outputs on the rendered video."""

    def __init__(self, model_file, audio_file, fps=24, offset=0, duration=-1, output_size=(1024, 1024),
                 resize_strategy="pad-zero", resize_layer=0, inference=False):
        super().__init__(audio_file, fps, offset, duration)
        self.stylegan2 = StyleGAN2(model_file, inference, output_size, resize_strategy, resize_layer)
        self.mapper = self.stylegan2.mapper
        self.synthesizer = self.stylegan2.synthesizer

    def process_mapper_inputs(self):
        return {"latent_z": torch.randn((1, 512))}

    def process_synthesizer_inputs(self, latent_w):
        return latent_w

    def process_outputs(self, video):
        return video

    process_outputs.stock = True   # un-overridden: generate.py may declare the postprocess chain pure (render/ffmpeg.py)
