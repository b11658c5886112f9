"""Tempo estimate and beat tracker for the once-per-track pre-pass: what
maua/audiovisual/audioreactive/selfsupervised/mir.py:27-32 asks of librosa,

    tempo = librosa.beat.tempo(onset_envelope=env, max_tempo=240, prior=lognorm(loc=0, scale=400, s=1), ac_size=120, hop_length=1024)
    beats = librosa.beat.beat_track(onset_envelope=env, trim=False, hop_length=1024, bpm=tempo)[1]

PARITY UNPINNED: librosa is an absent, un-pinned dependency (setup.py:60) and the reference holds no test for these calls.
The published algorithms are restated: the autocorrelation tempogram with a log-prior over tempo (librosa.beat.tempo /
feature.tempogram) and the dynamic-programming beat tracker of Ellis (2007) as librosa implements it.  Host numpy on a
[T] envelope, as in the reference (nothing here is GPU work: T is the number of video frames).  Note the reference does
not pass `sr`, so librosa's default of 22050 Hz sets the frame rate both calls assume (22050 / 1024 = 21.5 frames/s)
whatever the real one is; `sr` defaults to that here too, so tempi come out in the same (reference) units.
"""
from __future__ import annotations

import numpy as np
import scipy.signal
import scipy.stats

DEFAULT_SR = 22050


def tempogram(onset_envelope, win_length):
    """Local autocorrelation of the onset envelope: [win_length, T], each column normalised by its maximum."""
    env = np.asarray(onset_envelope, dtype=np.float64)
    n = len(env)
    padded = np.pad(env, int(win_length // 2), mode="linear_ramp", end_values=[0, 0])
    frames = np.lib.stride_tricks.sliding_window_view(padded, win_length)[:n].T      # [win_length, T], hop 1
    frames = frames * scipy.signal.get_window("hann", win_length, fftbins=True)[:, None]
    n_pad = scipy.fft.next_fast_len(2 * win_length - 1, real=True)
    power = np.abs(scipy.fft.rfft(frames, n=n_pad, axis=0)) ** 2
    ac = scipy.fft.irfft(power, n=n_pad, axis=0)[:win_length]
    peak = np.abs(ac).max(axis=0, keepdims=True)
    peak[peak < np.finfo(ac.dtype).tiny] = 1.0
    return ac / peak


def tempo_frequencies(n_bins, hop_length, sr):
    bpm = np.zeros(n_bins)
    bpm[0] = np.inf
    bpm[1:] = 60.0 * sr / (hop_length * np.arange(1.0, n_bins))
    return bpm


def tempo(onset_envelope, sr=DEFAULT_SR, hop_length=1024, start_bpm=120.0, std_bpm=1.0, ac_size=120.0, max_tempo=240.0, prior=None):
    """Global tempo (BPM in the units of `sr / hop_length` frames per second) -> float."""
    win_length = int(np.floor(ac_size * sr / hop_length))
    tg = tempogram(onset_envelope, win_length).mean(axis=1)
    bpms = tempo_frequencies(len(tg), hop_length, sr)
    if prior is None:
        with np.errstate(divide="ignore"):
            logprior = -0.5 * ((np.log2(bpms) - np.log2(start_bpm)) / std_bpm) ** 2
    else:
        logprior = prior.logpdf(bpms)
    if max_tempo is not None:
        logprior[: int(np.argmax(bpms < max_tempo))] = -np.inf
    return float(bpms[int(np.argmax(np.log1p(1e6 * tg) + logprior))])


def reference_prior():
    """scipy.stats.lognorm(loc=0, scale=400, s=1), the prior of mir.py:28."""
    return scipy.stats.lognorm(loc=0, scale=400, s=1)


def _local_score(onset_envelope, period):
    env = np.asarray(onset_envelope, dtype=np.float64)
    norm = env.std(ddof=1)
    if norm > 0:
        env = env / norm
    window = np.exp(-0.5 * (np.arange(-period, period + 1) * 32.0 / period) ** 2)
    return scipy.signal.convolve(env, window, "same")


def _beat_dp(localscore, period, tightness):
    """Ellis' recursion: cumscore[i] = localscore[i] + max over the previous beat's position of
    cumscore[j] - tightness * log((i - j) / period)^2, j between two periods and half a period back."""
    n = len(localscore)
    backlink = np.zeros(n, dtype=int)
    cumscore = np.zeros(n)
    window = np.arange(-2 * period, -int(np.round(period / 2)) + 1, dtype=int)
    txwt = -tightness * (np.log(-window / period) ** 2)
    first_beat = True
    threshold = 0.01 * localscore.max()
    for i in range(n):
        z_pad = int(np.maximum(0, min(-window[0], len(window))))   # window positions that fall before the start
        candidates = txwt.copy()
        candidates[z_pad:] = candidates[z_pad:] + cumscore[window[z_pad:]]
        best = int(np.argmax(candidates))
        cumscore[i] = localscore[i] + candidates[best]
        if first_beat and localscore[i] < threshold:
            backlink[i] = -1
        else:
            backlink[i] = window[best]
            first_beat = False
        window = window + 1
    return backlink, cumscore


def _local_max(x):
    m = np.zeros(len(x), dtype=bool)
    m[1:-1] = (x[1:-1] > x[:-2]) & (x[1:-1] >= x[2:])
    m[-1] = x[-1] > x[-2] if len(x) > 1 else False
    return m


def beat_track(onset_envelope, bpm, sr=DEFAULT_SR, hop_length=1024, tightness=100.0, trim=False):
    """Beat positions (frame indices, ascending) for a given tempo."""
    env = np.asarray(onset_envelope, dtype=np.float64)
    if not env.any():
        return np.array([], dtype=int)
    period = int(round(60.0 * (float(sr) / hop_length) / bpm))
    localscore = _local_score(env, period)
    backlink, cumscore = _beat_dp(localscore, period, tightness)
    maxes = _local_max(cumscore)
    med = np.median(cumscore[np.argwhere(maxes)])
    beats = [int(np.argwhere(cumscore * maxes * 2 > med).max())]
    while backlink[beats[-1]] >= 0:
        beats.append(int(backlink[beats[-1]]))
    beats = np.array(beats[::-1], dtype=int)
    smooth = scipy.signal.convolve(localscore[beats], scipy.signal.windows.hann(5), "same")
    threshold = 0.5 * np.sqrt((smooth ** 2).mean()) if trim else 0.0
    valid = np.argwhere(smooth > threshold)
    return beats[int(valid.min()): int(valid.max())]


def tempo_and_beats(onset_envelope):
    """mir.py:27-32: (tempo, beats) with the reference's arguments; a leading beat at frame 0 is dropped (:31-32)."""
    env = np.asarray(onset_envelope, dtype=np.float64).reshape(-1)
    bpm = tempo(env, max_tempo=240, prior=reference_prior(), ac_size=120, hop_length=1024)
    beats = list(int(b) for b in beat_track(env, bpm, hop_length=1024, trim=False))
    if beats and beats[0] == 0:
        del beats[0]
    return bpm, beats
