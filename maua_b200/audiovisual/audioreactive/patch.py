"""Random audio-reactive patch generator: mirror of maua/audiovisual/audioreactive/selfsupervised/patch.py:11-197
(first part).  A ``Patch`` draws a random stack of latent and noise sub-patches from a seeded
``torch.Generator`` and applies them (``latent_patch`` / ``noise_patch``, device kernels) to a latent palette.

The order and kind of every random draw follows the reference, so a given (seed, generator device) yields the same
sub-patch tables: pinned on a CPU generator against the reference's own class by tests/golden/make_patch_golden.py.
JSON ``save`` / ``load`` use the reference's keys, so patch files written by either side load in the other.
"""
from __future__ import annotations

import json
import math

import numpy as np
import torch

from . import noise as _noise
from .selfsupervised import latent_patch, noise_patch, spline_loop_latents

# selfsupervised/mir.py:9-11
UNITFEATS = ["rms", "drop_strength", "onsets", "spectral_flatness"]
ALLFEATS = ["chromagram", "tonnetz", "mfcc", "spectral_contrast"] + UNITFEATS

_NOISE_SIZES = [4, 8, 8, 16, 16, 32, 32, 64, 64, 128, 128, 256, 256, 512, 512, 1024, 1024]
_DEPTHS, _DEPTH_WEIGHTS = ["low", "mid", "high", "lowmid", "midhigh", "all"], [3, 3, 3, 2, 2, 1]
_BARS, _BAR_WEIGHTS = [4, 8, 16, 32], [2, 2, 2, 1]


def random_choice(rng, options, weights=None, n=1, replacement=False):
    """patch.py:11-19: one multinomial draw on the generator's device."""
    if weights is None:
        p = torch.ones(len(options), device=rng.device) / len(options)
    else:
        p = torch.tensor(weights, device=rng.device) / np.sum(weights)
    return options[p.multinomial(num_samples=n, replacement=replacement, generator=rng)]


def skewnorm(rng, a, loc, scale, size=()):
    """patch.py:22-31: skew-normal sample from two standard normals (scipy's construction)."""
    u0 = torch.randn(size, generator=rng, device=rng.device)
    v = torch.randn(size, generator=rng, device=rng.device)
    d = a / math.sqrt(1 + a ** 2)
    u1 = d * u0 + v * math.sqrt(1 - d ** 2)
    return loc + scale * torch.where(u0 >= 0, u1, -u1)


class Patch(torch.nn.Module):
    def __init__(self, features, segmentations, tempo, fps=24, seed=42, min_subpatches=2, max_subpatches=20, device="cuda"):
        super().__init__()
        self.seed, self.rng = seed, torch.Generator(device).manual_seed(seed)
        self.fps, self.tempo = fps, tempo
        self.features, self.segmentations = features, segmentations
        self.length = next(iter(features.values())).shape[0]
        rng = self.rng
        self.n_base_latents = torch.randint(3, 15, size=(), generator=rng, device=rng.device).item()
        self.sigma_base_noise = 1 + 9 * torch.rand((), generator=rng, device=rng.device).item()
        self.loops_base_noise = random_choice(rng, [1, 2, 4, 8, 16, 32, 64])
        self.ks = np.unique([k for (_, k) in segmentations]).tolist()
        self.min_subpatches, self.max_subpatches = min_subpatches, max_subpatches
        self.randomize_latent_patches()
        self.randomize_noise_patches()

    def __getstate__(self):  # a torch.Generator cannot be pickled
        state = {k: v for k, v in self.__dict__.items() if k != "rng"}
        state["device"] = self.rng.device
        return state

    def __setstate__(self, d):
        self.__dict__ = d
        self.rng = torch.Generator(d["device"]).manual_seed(d["seed"])

    def _count(self):
        return int(torch.randint(self.min_subpatches, self.max_subpatches, size=(), generator=self.rng, device=self.rng.device))

    def randomize_latent_patches(self):
        self.latent_patches = [self.random_latent_patch() for _ in range(self._count())]

    def randomize_noise_patches(self):
        self.noise_patches = [self.random_noise_patch() for _ in range(self._count())]

    def update_intensity(self, val):
        draw = lambda: skewnorm(self.rng, a=5, loc=val, scale=0.5).item()
        for p in self.latent_patches:
            p["seq_feat_weight"] = draw()
            p["mod_feat_weight"] = draw()
        for p in self.noise_patches:
            p["seq_feat_weight"] = draw()
            p["mod_feat_weight"] = draw()
            p["noise_std"] = draw()

    def random_latent_patch(self):
        r = self.rng
        return dict(
            patch_type=random_choice(r, ["segmentation", "feature", "loop"]),
            segments=random_choice(r, self.ks),
            loop_bars=random_choice(r, _BARS, weights=_BAR_WEIGHTS),
            seq_feat=random_choice(r, ALLFEATS),
            seq_feat_weight=1,
            mod_feat=random_choice(r, UNITFEATS),
            mod_feat_weight=1,
            merge_type=random_choice(r, ["average", "modulate"], weights=[1, 3]),
            merge_depth=random_choice(r, _DEPTHS, weights=_DEPTH_WEIGHTS),
        )

    def random_noise_patch(self):
        r = self.rng
        return dict(
            patch_type=random_choice(r, ["blend", "multiply", "loop"]),
            loop_bars=random_choice(r, _BARS, weights=_BAR_WEIGHTS),
            seq_feat=random_choice(r, ALLFEATS),
            seq_feat_weight=1,
            mod_feat=random_choice(r, UNITFEATS),
            mod_feat_weight=1,
            merge_type=random_choice(r, ["average", "modulate"], weights=[1, 3]),
            merge_depth=random_choice(r, _DEPTHS, weights=_DEPTH_WEIGHTS),
            noise_mean=0,
            noise_std=1,
        )

    def forward(self, latent_palette, downscale_factor=1, aspect_ratio=1):
        """-> (latents [T, num_ws, w_dim] on the device, list of 17 lazy per-layer noise sequencers)."""
        if not latent_palette.is_cuda:
            raise RuntimeError("Patch.forward: the latent palette must be a CUDA tensor (no CPU fallback)")
        self.rng.manual_seed(self.seed)
        base = torch.randperm(len(latent_palette), generator=self.rng, device=self.rng.device)[: self.n_base_latents]
        latents = spline_loop_latents(latent_palette[base.to(latent_palette.device)], self.length)
        for sub in self.latent_patches:
            latents = latent_patch(self.rng, latents, latent_palette, self.segmentations, self.features, self.tempo, self.fps, **sub)
        noise = [
            _noise.Loop(rng=self.rng, length=self.length, size=(round(aspect_ratio * s / downscale_factor), round(s / downscale_factor)),
                        n_loops=self.loops_base_noise, sigma=self.sigma_base_noise, device=latent_palette.device)
            for s in _NOISE_SIZES
        ]
        for sub in self.noise_patches:
            noise = noise_patch(self.rng, noise, self.features, self.tempo, self.fps, **sub)
        return latents, noise

    def __repr__(self):
        blocks = []
        for patches in (self.latent_patches, self.noise_patches):
            header = [""] + list(patches[0])
            rows = [[str(i + 1)] + [(f"{v:.4f}" if isinstance(v, float) else f"{v}").replace("spectral_", "") for v in p.values()]
                    for i, p in enumerate(patches)]
            widths = [max(len(r[n]) for r in [header] + rows) for n in range(len(header))]
            table = [header, ["-" * w for w in widths]] + rows
            blocks.append([" | ".join(cell.ljust(w) for cell, w in zip(r, widths)) for r in table])
        return ("Patch(\n  Latent(\n    " + "\n    ".join(blocks[0]) + "\n  ),\n  Noise(\n    " + "\n    ".join(blocks[1]) + "\n  )\n)")

    _SAVED = ("seed", "latent_patches", "noise_patches", "n_base_latents", "sigma_base_noise", "loops_base_noise")

    def save(self, path):
        with open(path, mode="w") as f:
            f.write(json.dumps({k: getattr(self, k) for k in self._SAVED}))

    @staticmethod
    def load(path, features, segmentations, tempo, fps, device):
        # the reference passes `device` in the seed position (patch.py:192); the loaded seed overrides it either way
        patch = Patch(features, segmentations, tempo, fps, device=device)
        with open(path, mode="r") as f:
            for key, val in json.loads(f.read()).items():
                setattr(patch, key, val)
        return patch
