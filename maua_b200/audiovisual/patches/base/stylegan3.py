"""StyleGAN3Patch: mirror of maua/audiovisual/patches/base/stylegan3.py:5-58 (four overridable stages).
The reference passes its ctor arguments positionally to a keyword-only constructor; keywords here."""
from ....GAN.wrappers.stylegan3 import StyleGAN3
from . import MauaPatch


class StyleGAN3Patch(MauaPatch):
    def __init__(self, model_file, audio_file, fps=24, offset=0, duration=-1, output_size=(1024, 1024),
                 resize_strategy="pad-zero", resize_layer=0):
        super().__init__(audio_file, fps, offset, duration)
        self.stylegan3 = StyleGAN3(model_file=model_file, output_size=output_size, strategy=resize_strategy, layer=resize_layer)
        self.mapper = self.stylegan3.mapper
        self.synthesizer = self.stylegan3.synthesizer

    def process_mapper_inputs(self):
        """-> {"latent_z", "truncation", "class_conditioning"}"""
        return {}

    def process_synthesizer_inputs(self, latent_w):
        """-> {"latents", "translation", "rotation"} tensors of [T, ...]"""
        return latent_w

    def process_outputs(self, video):
        return video
