"""Shared StyleGAN2 / StyleGAN3 facade: the API surface of maua/GAN/wrappers/stylegan.py:11-77 (``StyleGANMapper``,
``StyleGANSynthesizer``, ``StyleGAN`` with ``get_z_latents`` / ``get_w_latents``) re-stated over this package's networks.

Behaviour kept: ``model_file`` of ``None`` or the string ``"None"`` means a randomly initialised network built with the
reference's constructor arguments (z_dim 512, c_dim 0, w_dim 512, num_ws 18), anything else goes through the checkpoint
loaders; seeds select key latents exactly as the reference does ("1-12,24": half-open ranges, one
``numpy.random.RandomState(seed).randn(1, z_dim)`` row per seed).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
from torch import Tensor

from . import MauaGenerator, MauaMapper, MauaSynthesizer


def load_network(model_file, inference=False):
    """Checkpoint loaders of maua/GAN/load.py:192-207 (imported lazily: the networks import this package)."""
    from ..load import load_network as _load

    return _load(model_file, inference)


def is_random_init(model_file) -> bool:
    return model_file is None or model_file == "None"


def parse_seeds(spec: str) -> List[int]:
    """"3,7-10,42" -> [3, 7, 8, 9, 42]: comma separated seeds, "a-b" = range(a, b) (stylegan.py:59-66)."""
    seeds: List[int] = []
    for token in spec.split(","):
        if "-" in token:
            first, last = token.split("-")[:2]
            seeds.extend(range(int(first), int(last)))
        else:
            seeds.append(int(token))
    return seeds


class StyleGANMapper(MauaMapper):
    MapperClsFn = lambda: None  # set by the StyleGAN2 / StyleGAN3 subclasses: inference flag -> MappingNetwork class

    def __init__(self, model_file: str, inference: bool) -> None:
        super().__init__()
        if is_random_init(model_file):
            mapping_cls = type(self).MapperClsFn(inference)
            self.G_map = mapping_cls(z_dim=512, c_dim=0, w_dim=512, num_ws=18)
        else:
            self.G_map = load_network(model_file, inference).mapping
        self.z_dim = self.G_map.z_dim
        self.c_dim = self.G_map.c_dim
        targets = {"latent_z": (self.z_dim,), "truncation": (1,)}
        if self.c_dim > 0:
            targets["class_conditioning"] = (self.c_dim,)
        self.modulation_targets = targets

    def forward(self, latent_z: Tensor, class_conditioning: Optional[Tensor] = None, truncation: float = 1.0):
        return self.G_map.forward(latent_z, class_conditioning, truncation_psi=truncation)


class StyleGANSynthesizer(MauaSynthesizer):
    """Marker base of the two synthesizer wrappers (stylegan2.py / stylegan3.py)."""


class StyleGAN(MauaGenerator):
    __constants__ = ["z_dim", "c_dim", "w_dim", "num_ws", "res", "model_file"]
    MapperCls = StyleGANMapper
    SynthesizerCls = StyleGANSynthesizer

    def __init__(self, model_file=None, inference=False, output_size=None, strategy="stretch", layer=0) -> None:
        shared = dict(model_file=model_file, inference=inference)
        super().__init__(mapper_kwargs=dict(shared),
                         synthesizer_kwargs=dict(shared, output_size=output_size, strategy=strategy, layer=layer))
        mapping = self.mapper.G_map
        for name in ("z_dim", "c_dim", "w_dim", "num_ws"):
            setattr(self, name, getattr(mapping, name))
        self.res = self.synthesizer.G_synth.img_resolution
        self.model_file = model_file

    def get_z_latents(self, seeds: str) -> Tensor:
        """One float64 row of z per seed of the spec (see parse_seeds)."""
        rows = [np.random.RandomState(seed).randn(1, self.mapper.z_dim) for seed in parse_seeds(seeds)]
        return torch.from_numpy(np.concatenate(rows, axis=0))

    def get_w_latents(self, seeds: str, truncation=1):
        z = self.get_z_latents(seeds).to(self.mapper.G_map.w_avg.device)
        return self.mapper(z, truncation=truncation)

    def forward(self, z, *args, c=None, **kwargs):
        return self.synthesizer(self.mapper(z, c))
