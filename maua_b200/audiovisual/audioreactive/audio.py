"""The classic ``ar.*`` audio functions: mirror of maua/audiovisual/audioreactive/audio.py:15-110 (same names, arguments
and return conventions), computed on the device.

The reference takes host arrays (MauaPatch hands over ``self.audio`` as numpy, patches/base/__init__.py:14) and returns
host arrays: librosa HPSS (:84-93) and scipy Butterworth filtering (:96-110).  Here a host array is moved to the GPU,
processed by the library's kernels, and comes back in the container it arrived in (numpy in -> numpy out, CUDA tensor in ->
CUDA tensor out), so a patch file written against the reference runs unchanged.  There is no CPU implementation: without a
GPU these raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from ... import _lib
from . import features as _f

HOP = 1024


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("maua_b200.audioreactive: the audio functions need a CUDA device (no CPU fallback)")
    return torch.device("cuda")


def _to_device(audio):
    """(CUDA float tensor [N], restore) where restore() converts a device result back to the caller's container."""
    if torch.is_tensor(audio):
        if audio.is_cuda:
            return audio.reshape(-1), lambda t: t
        return audio.reshape(-1).to(_device()), lambda t: t.cpu()
    arr = np.asarray(audio).reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(_device()), lambda t: t.cpu().numpy()


def _pad_to_hop(y):
    """The STFT kernels take whole hops: zero-pad the tail (at most 1023 samples of silence)."""
    n = y.numel()
    m = (n + HOP - 1) // HOP * HOP
    return torch.nn.functional.pad(y, (0, m - n)) if m != n else y


def load_audio(audio_file, offset=0, duration=-1, cache=True):
    """(audio float32 [N] mono tensor, sr, duration in seconds) -- audio.py:15-48.  ``cache`` is accepted for signature
    compatibility: the reference's joblib cache of the decoded file is a host-side convenience this build does not keep.
    Decoding: torchaudio when it has a backend for the file, else the standard library for PCM / float WAV."""
    audio, sr = _decode(audio_file)
    total = audio.shape[-1] / sr
    if duration == -1 or total < duration:      # :29-33
        duration = total
        if offset != 0:
            duration -= offset
    audio = audio[:, int(offset * sr): int((offset + duration) * sr)].mean(0)
    return audio.contiguous(), sr, duration


def _decode(audio_file):
    try:
        import torchaudio

        audio, sr = torchaudio.load(audio_file)
        return audio.to(torch.float32), int(sr)
    except Exception:
        pass
    import wave

    with wave.open(audio_file, "rb") as w:
        sr, ch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise NotImplementedError(f"load_audio: {8 * width}-bit WAV (and non-WAV files without a torchaudio backend) not supported")
    return torch.from_numpy(a.reshape(-1, ch).T.copy()), sr


def harmonic(audio, sr, margin=8):
    """Harmonic component of a median-filtering HPSS (audio.py:84-87: librosa.effects.harmonic(y, margin)) on the device:
    the in-tree torch-native twin's arithmetic (features/audio.py:13-17: n_fft 2048, hop 1024, 31-tap medians, soft masks)."""
    y, restore = _to_device(audio)
    n = y.numel()
    return restore(_f.harmonic(_pad_to_hop(y.float()), margin=float(margin))[:n])


def percussive(audio, sr, margin=8):
    """Percussive component (audio.py:90-93), see harmonic."""
    y, restore = _to_device(audio)
    n = y.numel()
    return restore(_f.percussive(_pad_to_hop(y.float()), margin=float(margin))[:n])


def _butter_sosfilt(audio, sos):
    y, restore = _to_device(audio)
    x = y.to(torch.float64).contiguous()
    n = x.numel()
    out = torch.empty_like(x)
    scratch = torch.empty(4 * ((n + 255) // 256), device=x.device, dtype=torch.float64)
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_sosfilt(_lib.ptr(x), _lib.ptr(out), n, sos.ctypes.data_as(C.POINTER(C.c_double)), sos.shape[0],
                                          _lib.ptr(scratch), _lib.stream_ptr()))
    # scipy returns float64 for a float32 signal (the coefficient array is float64); tensors stay float32 for the kernels
    return restore(out) if not torch.is_tensor(audio) else restore(out.to(torch.float32))


def _butter(order, freqs, kind, sr):
    """Filter DESIGN (a dozen scalars) is scipy's on the host, exactly the call the reference makes (audio.py:98,104,110);
    the filtering of the signal is the device kernel."""
    from scipy import signal

    return signal.butter(order, freqs, kind, fs=sr, output="sos")


def low_pass(audio, sr, fmax=200, db_per_octave=12):
    """audio.py:96-99 (the reference passes ``db_per_octave`` as the Butterworth ORDER; kept)."""
    return _butter_sosfilt(audio, _butter(db_per_octave, fmax, "low", sr))


def high_pass(audio, sr, fmin=3000, db_per_octave=12):
    """audio.py:102-105."""
    return _butter_sosfilt(audio, _butter(db_per_octave, fmin, "high", sr))


def band_pass(audio, sr, fmin=200, fmax=3000, db_per_octave=12):
    """audio.py:108-111."""
    return _butter_sosfilt(audio, _butter(db_per_octave, [fmin, fmax], "band", sr))


def unmix(audio, sr):
    raise NotImplementedError("unmix / unmixed need the openunmix source-separation model (an absent third-party network, "
                              "outside SURVEY §8); pass the stems in yourself")


unmixed = unmix


def spleeted(audio, sr):
    raise NotImplementedError()   # as in the reference (audio.py:79-80)
