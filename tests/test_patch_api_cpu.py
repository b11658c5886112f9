"""Patch plugin API host logic (no GPU): same lookup rule and errors as maua/audiovisual/patches/base/__init__.py:28-45."""
import wave

import numpy as np
import pytest

from maua_b200.audiovisual.patches.base import MauaPatch, get_patch_from_file, load_audio
from maua_b200.audiovisual.render import get_output_class


def test_patch_lookup_by_file_and_name():
    cls = get_patch_from_file("tests/patches/sweep_patch.py")
    assert cls.__name__ == "SweepPatch" and issubclass(cls, MauaPatch)
    assert get_patch_from_file("tests/patches/sweep_patch.py", "SweepPatch") is cls
    with pytest.raises(Exception, match="Patch not found"):
        get_patch_from_file("tests/patches/sweep_patch.py", "Nope")


def test_load_audio_and_frame_count(tmp_path):
    p = str(tmp_path / "a.wav")
    y = (0.5 * np.sin(np.arange(48000 * 3) * 0.01)).astype(np.float32)
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes((np.stack([y, y], 1) * 32767).astype("<i2").tobytes())
    a, sr, dur = load_audio(p)
    assert sr == 48000 and abs(dur - 3.0) < 1e-6 and a.shape == (144000,)
    assert np.allclose(a.numpy(), y, atol=1e-4)
    a2, _, dur2 = load_audio(p, offset=1, duration=1.5)
    assert abs(dur2 - 1.5) < 1e-6
    patch = MauaPatch(p, fps=24)
    assert patch.n_frames == 72


def test_renderer_registry():
    assert get_output_class("memmap").__name__ == "MemMap" and get_output_class("ffmpeg").__name__ == "FFMPEG"
    with pytest.raises(NotImplementedError):
        get_output_class("gl")
