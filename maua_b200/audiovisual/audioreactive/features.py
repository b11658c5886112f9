"""Torch-native audio features on the device: host wrappers over mb_audio_onsets_rms.

Mirrors maua/audiovisual/audioreactive/selfsupervised/features/audio.py (``onsets(audio, sr)``,
``rms(y, sr)``, ``percussive(audio)``): same names, arguments and output shapes ([T,1] envelopes).
The audio must already be at sr = 1024 * fps (selfsupervised/sample.py:29-30), i.e. one hop per video frame.
CUDA tensors only -- there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ... import _lib

N_FFT, HOP, N_MELS = 2048, 1024, 128


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float32)
    logstep = np.float32(math.log(6.4) / 27.0)
    with np.errstate(divide="ignore"):
        log = np.float32(15.0) + np.log(np.maximum(f, np.float32(1e-30)) / np.float32(1000.0)) / logstep
    return np.where(f >= 1000.0, log, f / np.float32(200.0 / 3)).astype(np.float32)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float32)
    logstep = np.float32(math.log(6.4) / 27.0)
    return np.where(m >= 15.0, np.float32(1000.0) * np.exp(logstep * (m - np.float32(15.0))),
                    np.float32(200.0 / 3) * m).astype(np.float32)


def mel_filterbank(sr, n_fft=N_FFT, n_mels=N_MELS, fmin=0.0, fmax=None):
    """Slaney mel filterbank [n_mels, n_fft/2+1] (rosa/spectral.py:81-110), designed on the host in fp32."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    fft_f = torch.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mels = torch.linspace(float(_hz_to_mel(fmin)), float(_hz_to_mel(fmax)), n_mels + 2)
    mel_f = torch.from_numpy(_mel_to_hz(mels.numpy()))
    fdiff = torch.diff(mel_f)
    ramps = mel_f.reshape(-1, 1) - fft_f
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = torch.clamp_min(torch.minimum(lower, upper), 0.0)
    return (w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]).contiguous()


_ws_cache = {}


def _features(audio, sr, margin=8.0, want_percussive=False):
    if not audio.is_cuda:
        raise RuntimeError("maua_b200 audio features need a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    y = audio.detach().to(torch.float32).contiguous().reshape(-1)
    n = y.numel()
    if n % HOP:
        raise ValueError(f"audio length must be a multiple of {HOP} (resample to sr = 1024 * fps first)")
    T = n // HOP
    dev = y.device
    with torch.cuda.device(dev):
        fb = mel_filterbank(sr, fmax=11025.0).to(dev)
        onsets = torch.empty(T, device=dev)
        rms = torch.empty(T, device=dev)
        peaks = torch.empty(T, device=dev, dtype=torch.int32)
        n_peaks = torch.zeros(1, device=dev, dtype=torch.int32)
        perc = torch.empty(n, device=dev) if want_percussive else None
        nbytes = lib.mb_audio_workspace_bytes(n)
        ws = _ws_cache.get((nbytes, str(dev)))
        if ws is None:
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            _ws_cache.clear()
            _ws_cache[(nbytes, str(dev))] = ws
        off = (-ws.data_ptr()) % 256
        _lib.check(lib.mb_audio_onsets_rms(_lib.ptr(y), n, _lib.ptr(fb), float(margin), _lib.ptr(onsets), _lib.ptr(rms),
                                           _lib.ptr(peaks), _lib.ptr(n_peaks), _lib.ptr(perc),
                                           C.c_void_p(ws.data_ptr() + off), nbytes, _lib.stream_ptr()))
    return onsets, rms, peaks, n_peaks, perc


def onsets(audio, sr):
    """normalize(onset_strength(percussive(audio), sr)) -> [T,1] in [0,1] (features/audio.py:27-28)."""
    return _features(audio, sr)[0].unsqueeze(-1)


def rms(y, sr, frame_length=N_FFT, hop_length=HOP, center=True, pad_mode="reflect"):
    """Per-frame RMS -> [T,1] (features/audio.py:31-37); only the reference's default framing is built."""
    if (frame_length, hop_length, center, pad_mode) != (N_FFT, HOP, True, "reflect"):
        raise NotImplementedError("rms: only frame_length=2048, hop_length=1024, center=True, reflect padding")
    return _features(y, sr)[1].unsqueeze(-1)


def _hpss_component(audio, margin, which):
    if not audio.is_cuda:
        raise RuntimeError("maua_b200 audio features need a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    y = audio.detach().to(torch.float32).contiguous().reshape(-1)
    n = y.numel()
    if n % HOP:
        raise ValueError(f"audio length must be a multiple of {HOP} (resample to sr = 1024 * fps first)")
    with torch.cuda.device(y.device):
        out = torch.empty_like(y)
        nbytes = lib.mb_audio_workspace_bytes(n)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=y.device)
        off = (-ws.data_ptr()) % 256
        _lib.check(lib.mb_audio_hpss_component(_lib.ptr(y), n, float(margin), int(which), _lib.ptr(out),
                                               C.c_void_p(ws.data_ptr() + off), nbytes, _lib.stream_ptr()))
    return out


def harmonic(audio, margin=8.0):
    """HPSS harmonic component (features/audio.py:13-17)."""
    return _hpss_component(audio, margin, 0)


def percussive(audio, margin=8.0):
    """HPSS percussive component (features/audio.py:20-24)."""
    return _hpss_component(audio, margin, 1)


def onset_peaks(audio, sr):
    """Frame indices of the strict local maxima of onsets(audio, sr) (peak rule of signal.py:69-76), int64."""
    _, _, peaks, n_peaks, _ = _features(audio, sr)
    return peaks[: int(n_peaks.item())].to(torch.int64)


def onsets_rms(audio, sr):
    """Both envelopes from one device pass -> ([T,1], [T,1])."""
    o, r, _, _, _ = _features(audio, sr)
    return o.unsqueeze(-1), r.unsqueeze(-1)


# ---- spectral descriptors of the torch-native feature list (features/audio.py:59-133) ---------------------------------
def _spectrogram(y, sr, want_mag, want_mel, mel_fmax=None):
    """Device magnitude spectrogram [T,1025] and / or mel power spectrogram [T,128] (frame-major)."""
    if not y.is_cuda:
        raise RuntimeError("maua_b200 audio features need a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    y = y.detach().to(torch.float32).contiguous().reshape(-1)
    n = y.numel()
    if n % HOP:
        raise ValueError(f"audio length must be a multiple of {HOP} (resample to sr = 1024 * fps first)")
    T = n // HOP
    dev = y.device
    with torch.cuda.device(dev):
        mag = torch.empty(T, N_FFT // 2 + 1, device=dev) if want_mag else None
        mel = torch.empty(T, N_MELS, device=dev) if want_mel else None
        fb = mel_filterbank(sr, fmax=mel_fmax).to(dev) if want_mel else None
        nbytes = lib.mb_audio_workspace_bytes(n)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        _lib.check(lib.mb_audio_spectrogram(_lib.ptr(y), n, _lib.ptr(fb), _lib.ptr(mag), _lib.ptr(mel),
                                            C.c_void_p(ws.data_ptr() + off), nbytes, _lib.stream_ptr()))
    return mag, mel


def spectral_flatness(y, sr, n_fft=N_FFT, hop_length=HOP, amin=1e-10, power=2.0):
    """Geometric / arithmetic mean of the thresholded power spectrum per frame -> [T,1] (features/audio.py:123-133)."""
    if (n_fft, hop_length) != (N_FFT, HOP):
        raise NotImplementedError("spectral_flatness: only n_fft=2048, hop_length=1024")
    mag, _ = _spectrogram(y, sr, True, False)
    out = torch.empty(mag.shape[0], device=mag.device)
    with torch.cuda.device(mag.device):
        _lib.check(_lib.load().mb_spectral_flatness(_lib.ptr(mag), mag.shape[0], float(amin), float(power), _lib.ptr(out), _lib.stream_ptr()))
    return out.unsqueeze(-1)


def contrast_bands(sr, n_fft=N_FFT, fmin=200.0, n_bands=6, quantile=0.02):
    """Bin ranges [lo, hi) and quantile sizes of the octave bands, exactly as the reference's loop builds them
    (features/audio.py:81-108): host arithmetic on the FFT bin frequencies."""
    freq = torch.linspace(0, float(sr) / 2, int(1 + n_fft // 2))
    octa = torch.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** torch.arange(0, n_bands + 1))
    lo, hi, cnt = [], [], []
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        band = torch.logical_and(freq >= f_low, freq <= f_high)
        idx = band.flatten().nonzero()
        if k > 0:
            band[idx[0] - 1] = True
        if k == n_bands:
            band[idx[-1] + 1:] = True
        where = band.nonzero().flatten()
        first, last = int(where[0]), int(where[-1]) + 1
        if k < n_bands:
            last -= 1  # sub_band[:-1]
        q = torch.round(quantile * torch.sum(band))
        lo.append(first); hi.append(last); cnt.append(int(torch.maximum(q, torch.ones(()))))
    return lo, hi, cnt


def spectral_contrast(y, sr, n_fft=N_FFT, hop_length=HOP, fmin=200.0, n_bands=6, quantile=0.02, linear=False):
    """Octave-band spectral contrast -> [T, n_bands + 1] (features/audio.py:69-120)."""
    if (n_fft, hop_length) != (N_FFT, HOP):
        raise NotImplementedError("spectral_contrast: only n_fft=2048, hop_length=1024")
    mag, _ = _spectrogram(y, sr, True, False)
    lo, hi, cnt = contrast_bands(sr, n_fft, fmin, n_bands, quantile)
    nb, T = len(lo), mag.shape[0]
    arr = lambda v: (C.c_int32 * nb)(*v)
    out = torch.empty(T, nb, device=mag.device)
    scratch = torch.empty(2 * nb * T, device=mag.device)
    with torch.cuda.device(mag.device):
        _lib.check(_lib.load().mb_spectral_contrast(_lib.ptr(mag), T, nb, arr(lo), arr(hi), arr(cnt), int(bool(linear)), _lib.ptr(scratch),
                                                    _lib.ptr(out), _lib.stream_ptr()))
    return out


def mfcc(y, sr, n_mfcc=20, norm=False):
    """Mel-frequency cepstral coefficients -> [T, n_mfcc] (features/audio.py:59-64): dB mel spectrum, orthonormal DCT-II."""
    _, mel = _spectrogram(y, sr, False, True)
    out = torch.empty(mel.shape[0], int(n_mfcc), device=mel.device)
    with torch.cuda.device(mel.device):
        _lib.check(_lib.load().mb_mfcc(_lib.ptr(mel), mel.shape[0], int(n_mfcc), _lib.ptr(out), _lib.stream_ptr()))
    if norm is True:
        out = out / out.norm(p=2)
    return out
