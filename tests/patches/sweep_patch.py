"""A user patch file as the reference expects them (maua/audiovisual/patches/examples/*): a MauaPatch subclass with
the four stage methods.  Onsets drive an interpolation between two key latents."""
import numpy as np
import torch

from maua_b200.audiovisual import audioreactive as ar
from maua_b200.audiovisual.patches.base.stylegan3 import StyleGAN3Patch


class SweepPatch(StyleGAN3Patch):
    def process_audio(self):
        sr = 1024 * self.fps                      # one hop per video frame (selfsupervised/sample.py:29-30)
        n = self.n_frames * 1024
        t = np.arange(n) / sr
        y = np.interp(t, np.arange(len(self.audio)) / self.sr, self.audio).astype(np.float32)
        self.onsets = ar.onsets(torch.from_numpy(y).to(self.device), sr)[:, 0]

    def process_synthesizer_inputs(self, latent_w):
        w = self.stylegan3.get_z_latents("1-3").float().to(self.device)[:, None, :].repeat(1, self.synthesizer.num_ws, 1)
        lat = ar.single_weighted(w[0], w[1], ar.gaussian_filter(self.onsets, 1.0))
        return {"latents": lat}
