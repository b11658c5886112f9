"""Tempo / beat tracker restatement (maua_b200/audiovisual/audioreactive/beat.py; librosa absent: parity unpinned) on
synthetic click tracks with a known answer, in the frame-rate units the reference's calls imply (22050 / 1024 frames/s)."""
import numpy as np
import pytest

from maua_b200.audiovisual.audioreactive import beat as B

FPS = B.DEFAULT_SR / 1024.0


def click_track(bpm, seconds, jitter=0.0, seed=0):
    rng = np.random.RandomState(seed)
    n = int(seconds * FPS)
    env = 0.02 * rng.rand(n)
    period = 60.0 * FPS / bpm
    pos = np.arange(period / 2, n - 1, period)
    pos = np.round(pos + jitter * rng.randn(len(pos))).astype(int).clip(0, n - 1)
    env[pos] += 1.0
    env[(pos + 1).clip(0, n - 1)] += 0.4
    return env, pos


@pytest.mark.parametrize("bpm", [90.0, 120.0, 150.0])
def test_tempo_of_a_click_track(bpm):
    env, _ = click_track(bpm, 60)
    est = B.tempo(env, max_tempo=240, prior=B.reference_prior(), ac_size=120, hop_length=1024)
    # the estimate is quantised to integer autocorrelation lags: it must be one of the two lags around the true period
    period = 60.0 * FPS / bpm
    allowed = [60.0 * FPS / np.floor(period), 60.0 * FPS / np.ceil(period)]
    assert min(abs(est - a) for a in allowed) < 1e-6, (est, allowed)
    est8 = B.tempo(env, ac_size=8.0, max_tempo=320.0)                                # librosa's default prior / window
    assert min(abs(est8 - a) for a in allowed) < 1e-6, (est8, allowed)


def test_beats_land_on_the_clicks():
    env, clicks = click_track(120.0, 45, jitter=0.3)
    bpm, beats = B.tempo_and_beats(env)
    beats = np.array(beats)
    assert abs(bpm - 120.0) / 120.0 < 0.04
    assert np.all(np.diff(beats) > 0) and beats[0] > 0
    period = 60.0 * FPS / 120.0
    assert abs(np.median(np.diff(beats)) - period) < 1.0
    inner = clicks[(clicks > beats[0] - 2) & (clicks < beats[-1] + 2)]
    hit = [np.abs(beats - c).min() <= 2 for c in inner]
    assert np.mean(hit) > 0.9, np.mean(hit)


def test_tempogram_shape_and_normalisation():
    env, _ = click_track(100.0, 20)
    tg = B.tempogram(env, 128)
    assert tg.shape == (128, len(env)) and np.allclose(np.abs(tg).max(axis=0), 1.0)
    assert np.allclose(tg[0], 1.0)                       # lag 0 is the maximum of an autocorrelation
    assert B.tempo_frequencies(4, 1024, B.DEFAULT_SR)[0] == np.inf
    assert B.beat_track(np.zeros(100), 120.0).size == 0
