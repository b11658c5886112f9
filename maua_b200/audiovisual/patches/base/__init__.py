"""Patch plugin API: mirror of maua/audiovisual/patches/base/__init__.py:7-45 (MauaPatch, get_patch_from_file)."""
import torch


from ...audioreactive.audio import load_audio  # noqa: E402  (ar.load_audio, audioreactive/audio.py:15-48)


class MauaPatch:
    """Base of the user patch files (API of maua/audiovisual/patches/base/__init__.py:7-25): holds the decoded track
    (`audio` numpy float32 mono, `sr`, `duration` in seconds), the frame count of the render and the device; subclasses add
    a generator (`mapper`, `synthesizer`) and override the processing stages."""

    def __init__(self, audio_file, fps=24, offset=0, duration=-1) -> None:
        track, self.sr, self.duration = load_audio(audio_file, offset, duration)
        self.audio = track.numpy()
        self.audio_file, self.fps = audio_file, fps
        self.n_frames = round(self.duration * fps)
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def process_audio(self):
        """Stage 1 (optional): derive envelopes / features from `self.audio`."""

    def force_output_size(self, video):
        t, c, h, w = video.shape
        if (w, h) != tuple(self.synthesizer.output_size):
            # Lanczos-prefiltered bicubic resampling on the device (patches/base/__init__.py:21-25, ops/image.py:214-240)
            from .... import ops

            ow, oh = self.synthesizer.output_size
            if not torch.is_tensor(video):
                import numpy as np

                video = torch.from_numpy(np.ascontiguousarray(video))
            frames = video.to(self.device)
            scale = 255.0 if frames.dtype == torch.uint8 else 1.0
            out = ops.resample(frames.float() / scale, (oh, ow))
            video = (out.clamp(0, 1) * 255).round().to(torch.uint8) if scale == 255.0 else out
        return video


def get_patch_from_file(filepath, class_name=None):
    """The MauaPatch subclass DEFINED in the python file `filepath` (path relative to the working directory, as the
    reference's CLI passes it), optionally the one called `class_name` (patches/base/__init__.py:28-45)."""
    import importlib
    import inspect

    module_name = filepath.replace(".py", "").replace("/", ".")
    module = importlib.import_module(module_name)
    for name, cls in inspect.getmembers(module, inspect.isclass):
        defined_here = cls.__module__ == module_name      # not merely imported into the patch file
        wanted = class_name is None or name == class_name
        if defined_here and wanted and issubclass(cls, MauaPatch):
            return cls
    raise Exception(
        "Patch not found! Are you sure there is a class that extends MauaPatch in the file you specified and that the name you (might have) specified is correct?"
    )
