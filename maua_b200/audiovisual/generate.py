"""Top-level render entry: mirror of maua/audiovisual/generate.py:15-54 (generate_audiovisal_from_patch)."""
from typing import Tuple

import torch

from .patches.base import get_patch_from_file
from .render import get_output_class


@torch.inference_mode()
def generate_audiovisal_from_patch(audio_file: str, model_file: str, patch_file: str, patch_name: str, renderer: str,
                                   renderer_kwargs: dict, fps: float, out_size: Tuple[int], resize_strategy: str,
                                   resize_layer: int):
    patch = get_patch_from_file(patch_file, patch_name)(
        model_file, audio_file, fps=fps, offset=0, duration=-1, output_size=out_size, resize_strategy=resize_strategy,
        resize_layer=resize_layer,
    )
    patch.process_audio()
    mapper_inputs = patch.process_mapper_inputs()
    mapped_inputs = patch.mapper(**mapper_inputs) if mapper_inputs else None
    synthesizer_inputs = patch.process_synthesizer_inputs(mapped_inputs)
    postprocess = lambda video: patch.force_output_size(patch.process_outputs(video))
    renderer_kwargs = dict(renderer_kwargs)
    if renderer == "ffmpeg":
        renderer_kwargs["fps"] = patch.fps
        renderer_kwargs["audio_file"] = patch.audio_file
    video = get_output_class(renderer)(**renderer_kwargs)(patch.synthesizer, synthesizer_inputs, postprocess)
    return video, (patch.audio, patch.sr)
