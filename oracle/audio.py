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


# ---------------------------------------------------------------------------------------------------------
# constant-Q chroma.  Restates rosa/constantq.py:13-269, rosa/spectral.py:286-325
# (chroma_cqt) and rosa/convert.py:69-125 (cq_to_chroma, hz_to_midi).  PINNED bit-for-bit against the reference's own
# functions by tests/golden/make_audio_golden.py (tuning passed explicitly: the reference's tuning=None default runs
# rosa/pitch.py:estimate_tuning first, which is not on the restated path).
# ---------------------------------------------------------------------------------------------------------
C1_HZ = 32.70319566257483  # librosa.note_to_hz("C1"), the reference's fmin default (convert.py:129-130)


def cqt_frequencies(n_bins, fmin, bins_per_octave=12):
    """constantq.py:205-208."""
    return fmin * 2.0 ** (torch.arange(0, n_bins, dtype=torch.float) / bins_per_octave)


def constant_q_lengths(sr, fmin, n_bins=84, bins_per_octave=12, filter_scale=1, gamma=0):
    """constantq.py:211-216."""
    alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
    q = float(filter_scale) / alpha
    freq = fmin * (2.0 ** (torch.arange(n_bins, dtype=torch.float) / bins_per_octave))
    return q * sr / (freq + gamma / alpha)


def constant_q(sr, fmin, n_bins, bins_per_octave, filter_scale=1, gamma=0):
    """constantq.py:219-262 with pad_fft=True: hann-windowed complex exponentials, L1-normalised, centre-padded to
    the next power of two of the longest filter."""
    lengths = constant_q_lengths(sr, fmin, n_bins=n_bins, bins_per_octave=bins_per_octave, filter_scale=filter_scale, gamma=gamma)
    freqs = fmin * (2.0 ** (torch.arange(n_bins, dtype=torch.float) / bins_per_octave))
    filters = []
    for ilen, freq in zip(lengths, freqs):
        ilen2 = torch.div(ilen, 2, rounding_mode="floor")
        sig = torch.exp(torch.arange(-ilen2, ilen2, dtype=torch.float) * 1j * 2 * torch.pi * freq / sr)
        sig = sig * torch.hann_window(len(sig))
        sig = sig / sig.norm(p=1, dim=0)
        filters.append(sig)
    max_len = int(2.0 ** (torch.ceil(torch.log2(max(lengths)))))
    out = []
    for f in filters:
        lpad = int((max_len - f.shape[-1]) // 2)
        out.append(F.pad(f, (lpad, int(max_len - f.shape[-1] - lpad)), mode="constant"))
    return torch.stack(out), lengths


def sparsify_rows_dense(x, quantile=0.01):
    """constantq.py:146-163, returned as a DENSE matrix with the dropped entries zeroed (the reference builds a
    sparse COO tensor of the kept entries; `sparse.mm(D)` sums exactly those products)."""
    mags = torch.abs(x)
    norms = torch.sum(mags, axis=1, keepdims=True)
    mag_sort = torch.sort(mags, axis=1).values
    cumulative_mag = torch.cumsum(mag_sort / norms, axis=1)
    threshold_idx = torch.argmin((cumulative_mag < quantile).to(torch.uint8), axis=1)
    keep = mags >= mag_sort[torch.arange(x.shape[0]), threshold_idx][:, None]
    return torch.where(keep, x, torch.zeros_like(x)), keep


def cqt_filter_fft(sr, fmin, n_bins, bins_per_octave, filter_scale=1, sparsity=0.01, gamma=0.0):
    """constantq.py:119-143 -> (dense-with-zeros one-sided FFT basis [n_bins, n_fft/2+1], keep mask, n_fft)."""
    basis, lengths = constant_q(sr, fmin, n_bins, bins_per_octave, filter_scale, gamma)
    n_fft = basis.shape[1]
    basis = basis * (lengths[:, None] / float(n_fft))
    fft_basis = torch.fft.fft(basis, n=n_fft, axis=1)[:, : (n_fft // 2) + 1]
    dense, keep = sparsify_rows_dense(fft_basis, quantile=sparsity)
    return dense, keep, n_fft


def cqt(y, sr, hop_length=1024, fmin=None, n_bins=84, bins_per_octave=12, tuning=0.0, filter_scale=1, sparsity=0.01):
    """constantq.py:13-116 with gamma=0 (cqt = vqt special case): recursive octave-by-octave transform."""
    import numpy as np
    from torchaudio.functional import resample

    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    fmin = torch.tensor(C1_HZ).float() if fmin is None else torch.as_tensor(fmin).float()
    fmin = fmin * 2.0 ** (tuning / bins_per_octave)
    freqs = cqt_frequencies(n_bins, fmin, bins_per_octave=bins_per_octave)[-bins_per_octave:]
    fmin_t = torch.min(freqs)
    my_y, my_sr, my_hop = y, sr, hop_length
    resp = []
    for i in range(n_octaves):
        if i > 0:
            my_y = resample(my_y, my_sr, my_sr / 2, resampling_method="sinc_interp_kaiser")
            my_y = my_y * np.sqrt(2)
            my_sr /= 2.0
            my_hop //= 2
        dense, keep, n_fft = cqt_filter_fft(my_sr, fmin_t * 2.0**-i, n_filters, bins_per_octave, filter_scale, sparsity, gamma=0)
        dense = dense * np.sqrt(2**i)
        d = torch.stft(my_y, n_fft=n_fft, hop_length=my_hop, center=True, window=None, pad_mode="reflect", return_complex=True)[:, :-1]
        # the reference multiplies a sparse COO matrix: the same sparse op keeps the restatement bit-comparable
        resp.append(_sparse_mm_like_reference(dense, keep, d))
    max_col = min(c.shape[-1] for c in resp)
    out = torch.empty((n_bins, max_col), dtype=resp[0].dtype)
    end = n_bins
    for c in resp:
        n_oct = c.shape[0]
        if end < n_oct:
            out[:end] = c[-end:, :max_col]
        else:
            out[end - n_oct: end] = c[:, :max_col]
        end -= n_oct
    lengths = constant_q_lengths(sr, fmin, n_bins=n_bins, bins_per_octave=bins_per_octave, filter_scale=filter_scale, gamma=0)
    return out / torch.sqrt(lengths[:, None])


def _sparse_mm_like_reference(dense, keep, d):
    """fft_basis.mm(D) for the reference's sparse COO basis (constantq.py:160-163, 188)."""
    idx = keep.nonzero().permute(1, 0)
    sp = torch.sparse_coo_tensor(idx, dense[keep], size=dense.shape, dtype=dense.dtype)
    return sp.mm(d)


def hz_to_midi(f):
    import numpy as np
    return 12 * (np.log2(f) - np.log2(440.0)) + 69


def cq_to_chroma(n_input, bins_per_octave=12, n_chroma=12, fmin=None):
    """convert.py:69-117 (base_c=True, window=None)."""
    import numpy as np
    n_merge = float(bins_per_octave) / n_chroma
    fmin = torch.tensor(C1_HZ).float() if fmin is None else fmin
    m = torch.repeat_interleave(torch.eye(n_chroma), round(n_merge), dim=1)
    m = torch.roll(m, -int(n_merge // 2), dims=1)
    n_octaves = np.ceil(float(n_input) / bins_per_octave)
    m = torch.tile(m, (1, int(n_octaves)))[:, :n_input]
    midi_0 = hz_to_midi(fmin) % 12
    roll = int(torch.round(midi_0 * (n_chroma / 12.0)))
    return torch.roll(m, roll, dims=0).to(torch.float)


def chroma_cqt(y, sr, hop_length=1024, fmin=None, threshold=0.0, tuning=0.0, n_chroma=12, n_octaves=7, bins_per_octave=36, norm=True):
    """spectral.py:286-325 -> [12, T]."""
    c = torch.abs(cqt(y, sr=sr, hop_length=hop_length, fmin=fmin, n_bins=n_octaves * bins_per_octave,
                      bins_per_octave=bins_per_octave, tuning=tuning))
    chroma = cq_to_chroma(c.shape[0], bins_per_octave=bins_per_octave, n_chroma=n_chroma, fmin=fmin) @ c
    if threshold is not None:
        chroma[chroma < threshold] = 0.0
    if norm:
        chroma = chroma / chroma.max()
    return chroma
