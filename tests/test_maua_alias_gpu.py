"""The reference's import surface end to end on the device: a patch file that imports ``maua.*`` and calls the classic
``ar.*`` functions on host arrays, rendered through ``maua.audiovisual.generate``; the classic filters against scipy (the
reference's own call); pulse against vectors from the reference's own plp."""
import os
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_wav(path, seconds):
    from maua_b200.workload import sine_sweep

    y, sr = sine_sweep(seconds, tremolo_hz=4.0)
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes((np.clip(y, -1, 1) * 32767).astype("<i2").tobytes())


def test_reference_style_patch_renders(cuda, tmp_path):
    from maua.audiovisual.generate import generate_audiovisal_from_patch

    wav = str(tmp_path / "sweep.wav")
    _write_wav(wav, 3.0)
    torch.manual_seed(0)
    video, (audio, sr) = generate_audiovisal_from_patch(
        audio_file=wav, model_file=None, patch_file="tests/patches/maua_import_patch.py", patch_name=None, renderer="memmap",
        renderer_kwargs=dict(cache_file=str(tmp_path / "frames.npy"), batch_size=4), fps=4, out_size=(1024, 1024),
        resize_strategy="pad-zero", resize_layer=0)
    assert video.shape == (12, 3, 1024, 1024) and video.dtype == np.uint8 and sr == 48000
    assert 20 < float(video.mean()) < 235 and float(video[0].std()) > 5
    assert float(np.abs(video[0].astype(np.int16) - video[-1].astype(np.int16)).mean()) > 0.5   # the audio moves the image


@pytest.mark.parametrize("kind", ["low", "high", "band"])
def test_butterworth_filters_match_scipy(cuda, kind):
    """low_pass / high_pass / band_pass (audio.py:96-110) = scipy.signal.sosfilt(butter(...)): the reference's own call is the
    oracle.  The device kernel chains chunk states through A^256 in double: equal to scipy's serial recurrence to ~1e-12."""
    from scipy import signal

    from maua.audiovisual import audioreactive as ar

    sr = 48000
    g = np.random.default_rng(3)
    x = (g.standard_normal(3 * sr + 137) * 0.3).astype(np.float32)
    if kind == "low":
        got, sos = ar.low_pass(x, sr, 200, 12), signal.butter(12, 200, "low", fs=sr, output="sos")
    elif kind == "high":
        got, sos = ar.high_pass(x, sr, 3000, 12), signal.butter(12, 3000, "high", fs=sr, output="sos")
    else:
        got, sos = ar.band_pass(x, sr, 200, 3000, 12), signal.butter(12, [200, 3000], "band", fs=sr, output="sos")
    want = signal.sosfilt(sos, x)
    assert isinstance(got, np.ndarray) and got.dtype == np.float64 and got.shape == want.shape
    assert float(np.abs(got - want).max()) <= 1e-9 * max(1.0, float(np.abs(want).max()))
    dev = ar.low_pass(torch.from_numpy(x).to(cuda), sr) if kind == "low" else None
    if dev is not None:
        assert dev.is_cuda and dev.dtype == torch.float32


def test_classic_features_shapes_and_ranges(cuda):
    from maua.audiovisual import audioreactive as ar
    from maua_b200.workload import sine_sweep

    y, sr = sine_sweep(4.0, tremolo_hz=4.0)
    y = (y + 0.01 * np.random.default_rng(1).standard_normal(len(y))).astype(np.float32)
    T = (len(y) + 1023) // 1024
    on = ar.onsets(y, sr, type="rosa")
    vol = ar.volume(y, sr)
    ch = ar.chroma(y, sr)
    ton = ar.tonnetz(y, sr)
    pul = ar.pulse(y, sr)
    assert on.shape == (T,) and vol.shape == (T,) and pul.shape == (T,) and not on.is_cuda
    assert isinstance(ch, np.ndarray) and ch.shape == (T, 12) and ton.shape == (T, 6)
    for v in (on, vol, torch.from_numpy(ch), ton, pul):
        assert torch.isfinite(v).all() and float(v.min()) >= -1e-6 and float(v.max()) <= 1 + 1e-6
    assert ar.chroma(y, sr, notes=4).shape == (T, 4)
    h = ar.harmonic(y, sr, margin=4)
    assert isinstance(h, np.ndarray) and h.shape == y.shape
    assert ar.pitch_dominance(y, sr).shape == (12,) and ar.spectral_max(y, sr).shape == (T,)
    # device tensors stay on the device
    assert ar.volume(torch.from_numpy(y).to(cuda), sr).is_cuda


def test_pulse_matches_reference_vector(cuda):
    from maua.audiovisual.audioreactive.selfsupervised.features.audio import pulse

    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "audio.pt"))
    got = pulse(G["audio_exact"].to(cuda), G["sr"])[:, 0].cpu()
    # the tempogram keeps one bin per frame (the arg-max of the magnitude): identical bins -> identical pulse up to FFT rounding
    assert got.shape == G["pulse"].shape and float((got - G["pulse"]).abs().max()) < 2e-3


def test_load_audio_conventions(tmp_path):
    from maua.audiovisual import audioreactive as ar

    wav = str(tmp_path / "s.wav")
    _write_wav(wav, 2.0)
    audio, sr, dur = ar.load_audio(wav)
    assert sr == 48000 and abs(dur - 2.0) < 1e-6 and audio.shape == (96000,) and audio.dtype == torch.float32
    part, _, d2 = ar.load_audio(wav, offset=0.5, duration=1.0)
    assert d2 == 1.0 and part.shape == (48000,) and torch.equal(part, audio[24000:72000])
