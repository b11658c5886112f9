"""A patch file written against the REFERENCE's import surface (``maua.*`` module paths, the classic ``ar.*`` functions on
the host arrays MauaPatch holds, plain torch arithmetic on the results), as maua/audiovisual/patches/examples/*.py are."""
import torch

from maua.audiovisual import audioreactive as ar
from maua.audiovisual.patches.base.stylegan3 import StyleGAN3Patch


class ImportSurfacePatch(StyleGAN3Patch):
    def process_audio(self):
        lows = ar.low_pass(self.audio, self.sr, 4000, 12)
        self.kick = ar.resample(ar.onsets(self.audio, self.sr, type="rosa", prepercussive=2).reshape(-1, 1), self.n_frames)
        self.kick = ar.gaussian_filter(ar.normalize(self.kick), 1).reshape(-1, 1, 1)
        self.loud = ar.resample(ar.volume(lows, self.sr).reshape(-1, 1), self.n_frames).reshape(-1, 1, 1)
        self.notes = ar.resample(torch.from_numpy(ar.chroma(self.audio, self.sr, notes=4)), self.n_frames)
        ar.plot_signals([self.kick, self.loud])
        for name in ("kick", "loud", "notes"):
            assert torch.isfinite(getattr(self, name)).all(), name

    def process_mapper_inputs(self):
        return {"latent_z": self.stylegan3.get_z_latents("1-7")}

    def process_synthesizer_inputs(self, latent_w):
        chroma_latents = ar.multi_weighted(latent_w[:4], self.notes)
        loops = ar.spline_loops(latent_w[4:], self.n_frames, n_loops=2)
        latents = (1 - self.loud) * loops + self.loud * chroma_latents
        latents = (1 - self.kick) * latents + self.kick * latent_w[[5]]
        return {"latents": latents}
