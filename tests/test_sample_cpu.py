"""Host pieces of the random-patch entry point (audioreactive/sample.py; reference selfsupervised/sample.py:16-32)."""
import wave

import numpy as np
import pytest
import torch


def _wav(path, seconds, sr=44100, channels=2):
    t = np.arange(int(seconds * sr)) / sr
    y = 0.5 * np.sin(2 * np.pi * 440 * t)
    data = np.stack([y, 0.5 * y][:channels], axis=1)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes((data * 32767).astype("<i2").tobytes())


def test_load_audio_resamples_to_1024_samples_per_frame(tmp_path):
    from maua_b200.audiovisual.audioreactive.sample import load_audio

    _wav(tmp_path / "a.wav", 3.0)
    audio, sr = load_audio(str(tmp_path / "a.wav"), offset=1, duration=1.5, fps=24, device="cpu")
    assert sr == 1024 * 24 and audio.dtype == torch.float32
    assert len(audio) % 1024 == 0 and abs(len(audio) - 1.5 * sr) < 1024          # whole video frames of the cropped span
    # mono mix of (y, y/2) keeps the 440 Hz tone: dominant FFT bin at 440 Hz
    spec = torch.fft.rfft(audio).abs()
    assert abs(int(spec.argmax()) * sr / len(audio) - 440.0) < 2.0
    assert 0.3 < float(audio.abs().max()) < 0.45


def test_resize_strategy_parsing_matches_get_hook():
    """Strategy strings of StyleGAN2Synthesizer.change_output_resolution (maua/GAN/wrappers/stylegan2.py:216-283): mode, leading
    pads (top, left) and border value; the oracle restatement of get_hook agrees on the padding tuple."""
    from maua_b200.GAN.wrappers.stylegan2 import parse_resize_strategy
    from oracle.sg2_hooks import padding_of

    assert parse_resize_strategy("stretch", 16, (20, 24)) == ("stretch", (0, 0), 0.0)
    for how, mode, value in [("reflect", "reflect", 0.0), ("replicate", "replicate", 0.0), ("circular", "circular", 0.0), ("0.5", "constant", 0.5)]:
        for where in ("out", "left", "right", "top", "bottom"):
            got = parse_resize_strategy(f"pad-{how}-{where}", 16, (22, 26))
            (left, right, top, bottom), omode, ovalue = padding_of(f"pad-{how}-{where}", 16, (22, 26))
            assert got == (mode, (top, left), value) and (omode, ovalue) == (mode, value)
            assert left + right == 10 and top + bottom == 6
    assert parse_resize_strategy("pad-reflect-left", 16, (22, 26))[1] == (3, 10)
    assert parse_resize_strategy("pad-reflect-bottom", 16, (22, 26))[1] == (0, 5)
    with pytest.raises(Exception):
        parse_resize_strategy("squash", 16, (20, 24))
    with pytest.raises(NotImplementedError):
        parse_resize_strategy("pad-reflect-out", 16, (12, 24))
