"""sample.generate (maua/audiovisual/audioreactive/selfsupervised/sample.py:34-101) end to end: WAV -> music information ->
random Patch -> StyleGAN2 1024^2 (random init) with per-frame noise -> rgb24 frames into a sink."""
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class CountingSink:
    def __init__(self):
        self.nbytes, self.first = 0, None

    def write(self, b):
        if self.first is None:
            self.first = bytes(b[: 3 * 1024 * 4])
        self.nbytes += len(b)


def test_generate_random_patch_video(cuda, tmp_path):
    from maua_b200.audiovisual.audioreactive.sample import generate

    sr, seconds, fps = 48000, 14, 24
    t = np.arange(seconds * sr) / sr
    rng = np.random.RandomState(0)
    y = 0.3 * np.where((t % 7.0) < 3.5, np.sin(2 * np.pi * 220 * t), np.sin(2 * np.pi * 330 * t)) + ((t % 0.5) < 0.01) * rng.randn(len(t)) * 0.6
    wav = str(tmp_path / "track.wav")
    with wave.open(wav, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes((np.clip(y, -1, 1) * 32767).astype("<i2").tobytes())
    sink = CountingSink()
    torch.manual_seed(0)
    out_file, written, patch = generate(wav, seed=5, fps=fps, downscale_factor=1, batch_size=16, device="cuda", sink=sink)
    T = seconds * fps
    assert written == ((T - 1) // 16) * 16 == 320          # the reference's loop drops the last partial batch (sample.py:85)
    assert sink.nbytes == written * 1024 * 1024 * 3
    assert out_file.endswith("track_RandomPatches++_seed5_1024x1024.mp4")
    assert len(set(sink.first)) > 8                          # not a constant image
    assert 2 <= len(patch.latent_patches) < 20 and patch.length == T
    # a non-native size through the StyleGAN2 wrapper's output-size hook (the reference's defaults: "stretch" on layer 0):
    # half the size, 3:2 -> the 4 x 4 constant input becomes 2 x 3 and every block runs non-square
    sink2 = type(sink)()
    out2, written2, _ = generate(wav, seed=5, fps=fps, downscale_factor=2, aspect_ratio=1.5, batch_size=16, device="cuda", sink=sink2)
    assert written2 == written and sink2.nbytes == written2 * 768 * 512 * 3
    assert out2.endswith("track_RandomPatches++_seed5_768x512.mp4") and len(set(sink2.first)) > 8
