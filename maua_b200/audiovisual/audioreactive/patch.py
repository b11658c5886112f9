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
        """Re-draw every sub-patch weight around `val` (skew-normal, a = 5, scale 0.5), latent patches first (patch.py:87-95)."""
        for patches, keys in ((self.latent_patches, ("seq_feat_weight", "mod_feat_weight")),
                              (self.noise_patches, ("seq_feat_weight", "mod_feat_weight", "noise_std"))):
            for sub in patches:
                for key in keys:
                    sub[key] = skewnorm(self.rng, a=5, loc=val, scale=0.5).item()

    # One sub-patch = one dict; the tables list its keys in the reference's order (patch.py:97-125).  A tuple entry is drawn
    # from the generator (options, weights), "ks" stands for this track's segment counts, anything else is a constant
    # (the reference has its skew-normal weight draws commented out and uses these defaults).
    _COMMON_TAIL = (
        ("loop_bars", (_BARS, _BAR_WEIGHTS)),
        ("seq_feat", (ALLFEATS, None)),
        ("seq_feat_weight", 1),
        ("mod_feat", (UNITFEATS, None)),
        ("mod_feat_weight", 1),
        ("merge_type", (["average", "modulate"], [1, 3])),
        ("merge_depth", (_DEPTHS, _DEPTH_WEIGHTS)),
    )
    _LATENT_TABLE = (("patch_type", (["segmentation", "feature", "loop"], None)), ("segments", ("ks", None))) + _COMMON_TAIL
    _NOISE_TABLE = (("patch_type", (["blend", "multiply", "loop"], None)),) + _COMMON_TAIL + (("noise_mean", 0), ("noise_std", 1))

    def _draw_subpatch(self, table):
        sub = {}
        for key, spec in table:
            if isinstance(spec, tuple):
                options, weights = spec
                sub[key] = random_choice(self.rng, self.ks if options == "ks" else options, weights=weights)
            else:
                sub[key] = spec
        return sub

    def random_latent_patch(self):
        return self._draw_subpatch(self._LATENT_TABLE)

    def random_noise_patch(self):
        return self._draw_subpatch(self._NOISE_TABLE)

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

    @staticmethod
    def _table(patches):
        """Fixed-width text table of a list of sub-patch dicts (floats with four decimals, "spectral_" dropped)."""
        def cell(v):
            return (f"{v:.4f}" if isinstance(v, float) else f"{v}").replace("spectral_", "")

        grid = [[""] + list(patches[0])] + [[str(n)] + [cell(v) for v in sub.values()] for n, sub in enumerate(patches, start=1)]
        widths = [max(len(row[c]) for row in grid) for c in range(len(grid[0]))]
        grid.insert(1, ["-" * w for w in widths])
        return [" | ".join(text.ljust(w) for text, w in zip(row, widths)) for row in grid]

    def __repr__(self):
        indent = "\n    "
        latent, noise = indent.join(self._table(self.latent_patches)), indent.join(self._table(self.noise_patches))
        return f"Patch(\n  Latent({indent}{latent}\n  ),\n  Noise({indent}{noise}\n  )\n)"

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
