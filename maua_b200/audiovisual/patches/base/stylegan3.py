"""StyleGAN3Patch: mirror of maua/audiovisual/patches/base/stylegan3.py:5-58 (four overridable stages).
The reference passes its ctor arguments positionally to a keyword-only constructor; keywords here."""
from ....GAN.wrappers.stylegan3 import StyleGAN3
from . import MauaPatch


class StyleGAN3Patch(MauaPatch):
    """Patch bound to a StyleGAN3 generator.  The render loop calls, in order: process_audio, process_mapper_inputs
    (-> keyword tensors for the mapper, or {} to skip it), process_synthesizer_inputs (-> dict of [T, ...] tensors:
    "latents", optionally "translation" / "rotation"), and process_outputs on the rendered video."""

    def __init__(self, model_file, audio_file, fps=24, offset=0, duration=-1, output_size=(1024, 1024),
                 resize_strategy="pad-zero", resize_layer=0):
        super().__init__(audio_file, fps, offset, duration)
        generator = StyleGAN3(model_file=model_file, output_size=output_size, strategy=resize_strategy, layer=resize_layer)
        self.stylegan3, self.mapper, self.synthesizer = generator, generator.mapper, generator.synthesizer

    def process_mapper_inputs(self):
        return {}

    def process_synthesizer_inputs(self, latent_w):
        return latent_w

    def process_outputs(self, video):
        return video

    process_outputs.stock = True   # un-overridden: generate.py may declare the postprocess chain pure (render/ffmpeg.py)
