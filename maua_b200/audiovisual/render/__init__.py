"""Frame sinks of the render loop: the ``Renderer`` base and the name -> class lookup the entry point uses
(API of maua/audiovisual/render/__init__.py:1-19).  Unlike the reference, a renderer refuses to exist without a GPU:
this package has no CPU path."""
import importlib

import torch

# renderer name (the `renderer` argument of generate_audiovisal_from_patch) -> (module, class); imported on first use
_RENDERERS = {"memmap": (".memmap", "MemMap"), "ffmpeg": (".ffmpeg", "FFMPEG")}


class Renderer:
    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError("maua_b200 renderers need a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda")


def get_output_class(renderer):
    try:
        module, cls = _RENDERERS[renderer]
    except KeyError:
        raise NotImplementedError(f"unknown renderer '{renderer}' (available: {sorted(_RENDERERS)})") from None
    return getattr(importlib.import_module(module, __name__), cls)
