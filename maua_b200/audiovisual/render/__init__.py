"""Renderers: mirror of maua/audiovisual/render/__init__.py:1-19 (Renderer, get_output_class)."""
import torch


class Renderer:
    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError("maua_b200 renderers need a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda")


def get_output_class(renderer):
    if renderer == "memmap":
        from .memmap import MemMap

        return MemMap
    if renderer == "ffmpeg":
        from .ffmpeg import FFMPEG

        return FFMPEG
    raise NotImplementedError
