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


def onset_strength_median(y, sr):
    """rosa/beat.py:10-23 with aggregate = median over the mel bins (torch.median: the lower middle value), as plp passes it
    (:41-43)."""
    s = power_to_db(melspectrogram(y, sr, fmax=11025.0))
    env = torch.clamp(s[:, 1:] - s[:, :-1], min=0).median(dim=0).values
    env = torch.nn.functional.pad(env, (2, 0))[: s.shape[1]]
    return env


def plp(y, sr, hop_length=1024, win_length=1024, tempo_min=60, tempo_max=180):
    """Predominant local pulse, rosa/beat.py:41-75: Fourier tempogram of the (median-aggregated) onset envelope (hann window
    of min(T, 1024) frames, hop 1), kept to [tempo_min, tempo_max] bpm, only the per-frame peak bin survives, unit magnitude,
    inverse transform, positive part, min-max normalised."""
    env = onset_strength_median(y, sr)
    n = min(len(env), win_length)
    win = torch.hann_window(n)
    ft = torch.stft(env, n_fft=n, hop_length=1, center=True, window=win, pad_mode="reflect", return_complex=True)
    freqs = torch.linspace(0, float(sr * 60 / float(hop_length)) / 2, int(1 + n // 2))
    ft[freqs < tempo_min] = 0
    ft[freqs > tempo_max] = 0
    mag = torch.log1p(1e6 * torch.abs(ft))
    ft[mag < mag.max(dim=0, keepdim=True).values] = 0
    ft = ft / (torch.finfo(ft.dtype).tiny ** 0.5 + torch.abs(ft.abs().max(dim=0, keepdim=True).values))
    pulse = torch.istft(ft, n_fft=n, hop_length=1, center=True, window=win, length=len(env))
    pulse = torch.clamp(pulse, torch.zeros(()), pulse.max())
    pulse = pulse - pulse.min()
    return pulse / (pulse.max() + 1e-8)


def pulse(y, sr):
    """features/audio.py:67-68 -> [T, 1]."""
    return plp(percussive(y), sr).unsqueeze(-1)


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


# ---------------------------------------------------------------------------------------------------------
# tuning estimate (rosa/pitch.py:9-120) and CENS chroma (rosa/spectral.py:164-280): the rest of chromagram()
# (features/audio.py:44).  estimate_tuning is PINNED bit-for-bit against the reference's own function by
# tests/golden/make_audio_golden.py.  chroma_cens evaluates a natural cubic spline whose coefficients the reference
# takes from the third-party torchcubicspline (un-pinned in setup.py:104, absent from the checkout): the published
# algorithm is restated (natural_cubic_coeffs) and checked against scipy's CubicSpline(bc_type="natural") --
# PARITY UNPINNED against the reference's own call for that stage.
# ---------------------------------------------------------------------------------------------------------
def localmax(x):
    """pitch.py:88-97 (axis 0)."""
    xp = F.pad(x, (0, 0, 1, 1))
    return (x > xp[:-2]) & (x >= xp[2:])


def piptrack(y, sr, n_fft=2048, fmin=150.0, fmax=4000.0, threshold=0.1):
    """pitch.py:27-85 with the reference defaults (hop_length=None -> torch.stft's n_fft // 4)."""
    s = torch.stft(y, n_fft=n_fft, hop_length=None, center=True, window=_win(n_fft, y.device), pad_mode="reflect",
                   return_complex=True)[:, :-1].abs()
    fmin = max(fmin, 0)
    fmax = min(fmax, float(sr) / 2)
    fft_freqs = torch.linspace(0, float(sr) / 2, int(1 + n_fft // 2))
    avg = 0.5 * (s[2:] - s[:-2])
    shift = 2 * s[1:-1] - s[2:] - s[:-2]
    shift = avg / (shift + (torch.abs(shift) < torch.finfo(shift.dtype).tiny))
    avg = F.pad(avg, [0, 0, 1, 1], mode="constant")
    shift = F.pad(shift, [0, 0, 1, 1], mode="constant")
    dskew = 0.5 * avg * shift
    pitches = torch.zeros_like(s)
    mags = torch.zeros_like(s)
    freq_mask = ((fmin <= fft_freqs) & (fft_freqs < fmax)).reshape((-1, 1))
    ref_value = threshold * torch.max(s, axis=0).values
    idx = torch.argwhere(freq_mask & localmax(s * (s > ref_value)))
    pitches[idx[:, 0], idx[:, 1]] = (idx[:, 0] + shift[idx[:, 0], idx[:, 1]]) * float(sr) / n_fft
    mags[idx[:, 0], idx[:, 1]] = s[idx[:, 0], idx[:, 1]] + dskew[idx[:, 0], idx[:, 1]]
    return pitches, mags


def pitch_tuning(frequencies, resolution=0.01, bins_per_octave=12):
    """pitch.py:100-120: histogram peak of the pitch residuals relative to the bin grid."""
    import numpy as np

    frequencies = torch.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not torch.any(frequencies):
        return 0.0
    residual = (bins_per_octave * torch.log2(frequencies / (440.0 / 16))) % 1.0
    residual[residual >= 0.5] -= 1.0
    bins = int(np.ceil(1.0 / resolution))
    counts = torch.histc(residual, bins=bins, min=-0.5, max=0.5)
    tuning = torch.linspace(-0.5, 0.5, bins + 1)
    return tuning[torch.argmax(counts)]


def estimate_tuning(y, sr, n_fft=2048, resolution=0.01, bins_per_octave=12):
    """pitch.py:9-24."""
    pitch, mag = piptrack(y, sr, n_fft=n_fft)
    pitch_mask = pitch > 0
    threshold = torch.median(mag[pitch_mask]) if pitch_mask.any() else 0.0
    return pitch_tuning(pitch[(mag >= threshold) & pitch_mask], resolution=resolution, bins_per_octave=bins_per_octave)


def natural_cubic_coeffs(x, y):
    """Natural cubic spline through (x_i, y_i), float64: per interval S(t) = a + b f + c f^2 + d f^3, f = t - x_i
    (the form torchcubicspline's NaturalCubicSpline evaluates; spectral.py:189,192-203 consume (x, a, b, c, d))."""
    import numpy as np

    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    h = np.diff(x)
    m = np.zeros(n)  # second derivatives, m_0 = m_{n-1} = 0
    if n > 2:
        a_ = np.zeros((n - 2, n - 2))
        rhs = 6.0 * ((y[2:] - y[1:-1]) / h[1:] - (y[1:-1] - y[:-2]) / h[:-1])
        for i in range(n - 2):
            a_[i, i] = 2.0 * (h[i] + h[i + 1])
            if i > 0:
                a_[i, i - 1] = h[i]
            if i < n - 3:
                a_[i, i + 1] = h[i + 1]
        m[1:-1] = np.linalg.solve(a_, rhs)
    a = y[:-1]
    b = (y[1:] - y[:-1]) / h - h * (2.0 * m[:-1] + m[1:]) / 6.0
    c = m[:-1] / 2.0
    d = (m[1:] - m[:-1]) / (6.0 * h)
    return x, a, b, c, d


def cens_quantiser_knots():
    """spectral.py:164-188: the knots of the smooth 4-step quantisation curve."""
    import numpy as np

    steps = [0.4, 0.2, 0.1, 0.05]
    p1, p2, p3, p4 = np.diff(list(reversed(steps + [0])))
    xs = [torch.linspace(-0.1, 0.025, 101)[:-1], torch.linspace(0.025, p1, 11)[:-1], torch.linspace(p1, p1 + p2, 11)[:-1],
          torch.linspace(p1 + p2, p1 + p2 + p3, 11)[:-1], torch.linspace(p1 + p2 + p3, 0.5, 11)[:-1], torch.linspace(0.5, 1.1, 100)]
    ys = torch.cat((0.5 * torch.ones(len(xs[0])), xs[1] / p1, (xs[2] - p1) / p2 + 1, (xs[3] - p1 - p2) / p3 + 2,
                    (xs[4] - p1 - p2 - p3) / p4 + 3, 4.5 * torch.ones(len(xs[5]))))
    return torch.cat(xs), ys


def spline_quantize(chroma):
    """spectral.py:192-219: spline_eval on the quantiser knots, then the smooth step function (h = 0.25, alpha = 20)."""
    import numpy as np

    xs, ys = cens_quantiser_knots()
    x, a, b, c, d = (torch.from_numpy(np.asarray(v)).float() for v in natural_cubic_coeffs(xs.numpy(), ys.numpy()))
    index = (torch.bucketize(chroma, x) - 1).clamp(0, len(b) - 1)
    f = chroma - x[index]
    w = a[index] + (b[index] + (c[index] + d[index] * f) * f) * f
    alpha, hq = 20, 0.25
    r = (w - 0.5) - torch.floor(w - 0.5) - 0.5
    m = 1 / (1 + np.exp(-alpha)) - 0.5
    return hq * (torch.floor(w - 0.5) + 1 / (2 * m) * 1 / (1 + torch.exp(-2 * alpha * r)))


def chroma_cens(y, sr, hop_length=1024, tuning=None, win_len_smooth=41):
    """spectral.py:239-280 -> [12, T]."""
    if tuning is None:
        tuning = float(estimate_tuning(y, sr, bins_per_octave=36))
    chroma = chroma_cqt(y, sr, hop_length=hop_length, tuning=tuning, norm=False)
    chroma = chroma / torch.norm(chroma, p=1, dim=0)
    q = spline_quantize(chroma)
    win = torch.hann_window(win_len_smooth + 2)
    win = win / torch.sum(win)
    cens = F.conv1d(q.unsqueeze(0), win.tile(12, 1, 1), groups=12, padding="same").squeeze(0)
    return cens / torch.norm(cens, p=2, dim=0)


def chromagram(audio, sr):
    """features/audio.py:44."""
    return chroma_cens(harmonic(audio), sr).T


# ---- spectral descriptors (features/audio.py:59-133) -------------------------------------------------------------------
def dct(x, norm=None):
    """rosa/spectral.py:35-56 (FFT-based DCT-II over the last axis)."""
    import numpy as np

    x_shape = x.shape
    N = x_shape[-1]
    x = x.contiguous().view(-1, N)
    v = torch.cat([x[:, ::2], x[:, 1::2].flip([1])], dim=1)
    Vc = torch.view_as_real(torch.fft.fft(v, dim=1))
    k = -torch.arange(N, dtype=x.dtype)[None, :] * np.pi / (2 * N)
    V = Vc[:, :, 0] * torch.cos(k) - Vc[:, :, 1] * torch.sin(k)
    if norm == "ortho":
        V[:, 0] /= np.sqrt(N) * 2
        V[:, 1:] /= np.sqrt(N / 2) * 2
    return 2 * V.view(*x_shape)


def mfcc(y, sr, n_mfcc=20, norm=False):
    """features/audio.py:59-64."""
    S = power_to_db(melspectrogram(y, sr))
    M = dct(S.permute(1, 0), norm="ortho").permute(1, 0)[:n_mfcc]
    if norm is True:
        M = M / M.norm(p=2)
    return M.T


def spectral_contrast(y, sr, fmin=200.0, n_bands=6, quantile=0.02, linear=False):
    """features/audio.py:69-120."""
    S = spectrogram(y)
    freq = torch.linspace(0, float(sr) / 2, int(1 + N_FFT // 2))
    octa = torch.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** torch.arange(0, n_bands + 1))
    valley = torch.zeros((n_bands + 1, S.shape[1]))
    peak = torch.zeros_like(valley)
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        current_band = torch.logical_and(freq >= f_low, freq <= f_high)
        idx = current_band.flatten().nonzero()
        if k > 0:
            current_band[idx[0] - 1] = True
        if k == n_bands:
            current_band[idx[-1] + 1:] = True
        sub_band = S[current_band]
        if k < n_bands:
            sub_band = sub_band[:-1]
        idx = torch.round(quantile * torch.sum(current_band))
        idx = int(torch.maximum(idx, torch.ones(())))
        sortedr = torch.sort(sub_band, dim=0).values
        valley[k] = torch.mean(sortedr[:idx], dim=0)
        peak[k] = torch.mean(sortedr[-idx:], dim=0)
    if linear:
        return (peak - valley).T
    return (power_to_db(peak) - power_to_db(valley)).T


def spectral_flatness(y, sr, amin=1e-10, power=2.0):
    """features/audio.py:123-133."""
    S = spectrogram(y, power=1.0)
    S_thresh = torch.maximum(torch.tensor(amin), S ** power)
    gmean = torch.exp(torch.mean(torch.log(S_thresh), axis=0))
    amean = torch.mean(S_thresh, axis=0)
    return (gmean / amean).unsqueeze(-1)
