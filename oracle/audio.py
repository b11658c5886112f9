"""Oracle: torch-native audio features of the reference, fp32 on CPU.  TEST INFRASTRUCTURE ONLY.

Restates maua/audiovisual/audioreactive/selfsupervised/features/{audio.py,processing.py} and
rosa/{spectral,beat,convert,helpers}.py for the onset / rms path.
PINNED: tests/golden/make_audio_golden.py imports the reference's own modules (with the import stubs of
SURVEY Appendix C.2) and checks these functions against them bit-for-bit before writing the fixtures.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

N_FFT, HOP = 2048, 1024


def _win(n, device):
    return torch.hann_window(n, device=device)  # periodic hann, rosa/spectral.py:11,17


def stft(y, n_fft=N_FFT, hop=HOP):
    """rosa/spectral.py:10-21 -> complex [n_fft/2+1, 1 + len(y)//hop] (centered, reflect padded)."""
    return torch.stft(y, n_fft=n_fft, hop_length=hop, center=True, window=_win(n_fft, y.device), pad_mode="reflect",
                      return_complex=True)


def istft(spec, length, n_fft=N_FFT, hop=HOP):
    """rosa/spectral.py:24-32."""
    return torch.istft(spec, n_fft=n_fft, hop_length=hop, center=True, window=_win(n_fft, spec.device), length=length)


def spectrogram(y, power=1.0):
    """rosa/spectral.py:59-62: the last STFT column is dropped."""
    return stft(y)[:, :-1].abs() ** power


def hz_to_mel(f):
    """Slaney scale, rosa/convert.py:15-41 (htk=False)."""
    f = torch.as_tensor(f, dtype=torch.float32)
    lin = f / (200.0 / 3)
    logstep = math.log(6.4) / 27.0
    log = 15.0 + torch.log(f.clamp_min(1e-30) / 1000.0) / logstep
    return torch.where(f >= 1000.0, log, lin)


def mel_to_hz(m):
    """rosa/convert.py:44-66."""
    logstep = math.log(6.4) / 27.0
    return torch.where(m >= 15.0, 1000.0 * torch.exp(logstep * (m - 15.0)), (200.0 / 3) * m)


def mel_filterbank(sr, n_fft=N_FFT, n_mels=128, fmin=0.0, fmax=None):
    """rosa/spectral.py:73-110: triangular Slaney filters, area-normalised."""
    fmax = float(sr) / 2 if fmax is None else fmax
    fft_f = torch.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(torch.linspace(float(hz_to_mel(fmin)), float(hz_to_mel(fmax)), n_mels + 2))
    fdiff = torch.diff(mel_f)
    ramps = mel_f.reshape(-1, 1) - fft_f
    w = torch.zeros(n_mels, 1 + n_fft // 2)
    for i in range(n_mels):
        w[i] = torch.maximum(torch.zeros(()), torch.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    return w * (2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels]))[:, None]


def melspectrogram(y, sr, fmax=None):
    """rosa/spectral.py:65-70 (power 2)."""
    return mel_filterbank(sr, fmax=fmax) @ spectrogram(y, power=2.0)


def power_to_db(s, top_db=80.0):
    """rosa/convert.py:7-12 with ref 1, amin 1e-10; the floor uses the GLOBAL maximum."""
    db = 10.0 * torch.log10(torch.maximum(torch.tensor(1e-10), s))
    db -= 10.0 * torch.log10(torch.maximum(torch.tensor(1e-10), torch.ones(())))
    return torch.maximum(db, db.max() - top_db)


def median_time(s, k=31):
    """processing.py:75-85 with k=(1,ks): median over ks neighbours along time, reflect padded."""
    x = F.pad(s[None, None], (k // 2, k // 2, 0, 0), mode="reflect")[0, 0]
    return x.unfold(1, k, 1).median(dim=-1).values


def median_freq(s, k=31):
    """processing.py:75-85 with k=(ks,1): median over ks neighbours along frequency."""
    x = F.pad(s[None, None], (0, 0, k // 2, k // 2), mode="reflect")[0, 0]
    return x.unfold(0, k, 1).median(dim=-1).values


def softmask(x, x_ref, power=2.0):
    """rosa/spectral.py:120-142 (finite power, split_zeros False)."""
    z = torch.maximum(x, x_ref)
    bad = z < torch.finfo(torch.float32).tiny
    z = torch.where(bad, torch.ones(()), z)
    m, r = (x / z) ** power, (x_ref / z) ** power
    return torch.where(bad, torch.zeros(()), m / (m + r))


def hpss(d, margin=1.0, power=2.0, k=31):
    """rosa/spectral.py:145-161 -> (harmonic, percussive) complex spectra."""
    mag = d.abs()
    phase = torch.exp(1.0j * torch.angle(d))
    harm, perc = median_time(mag, k), median_freq(mag, k)
    if margin == 1:
        raise NotImplementedError("split_zeros branch (margin == 1) is not on the render path")
    return (mag * softmask(harm, perc * margin, power)) * phase, (mag * softmask(perc, harm * margin, power)) * phase


def percussive(y, margin=8.0):
    """features/audio.py:20-24."""
    return istft(hpss(stft(y), margin=margin)[1], length=len(y))


def harmonic(y, margin=8.0):
    """features/audio.py:13-17."""
    return istft(hpss(stft(y), margin=margin)[0], length=len(y))


def onset_strength(y, sr):
    """rosa/beat.py:10-23: mel dB flux, mean over bands, shifted by 1 + n_fft // (2 hop) = 2 frames."""
    s = power_to_db(melspectrogram(y, sr, fmax=11025.0).abs())
    env = torch.clamp_min(s[:, 1:] - s[:, :-1], 0.0).mean(dim=0)
    env = F.pad(env, (1 + N_FFT // (2 * HOP), 0))
    return env[: s.shape[1]]


def normalize(x):
    """processing.py:53-56."""
    x = x - x.min()
    return x / (x.max() + 1e-8)


def onsets(y, sr):
    """features/audio.py:27-28 -> [T, 1] in [0, 1]."""
    return normalize(onset_strength(percussive(y), sr).unsqueeze(-1))


def rms(y, frame=N_FFT, hop=HOP):
    """features/audio.py:31-37 -> [T, 1]."""
    x = F.pad(y[None, None], (frame // 2, frame // 2), mode="reflect")[0, 0].unfold(0, frame, hop)[:-1]
    return x.abs().pow(2).mean(dim=1).sqrt().unsqueeze(-1)


def peak_indices(env):
    """Strict local maxima with index-clamped neighbours (signal.py:69-76, processing.py:108-116): int64 indices."""
    e = env.reshape(-1)
    n = len(e)
    idx = torch.arange(n)
    m = (e > e[(idx + 1).clamp(0, n - 1)]) & (e > e[(idx - 1).clamp(0, n - 1)])
    return idx[m]
