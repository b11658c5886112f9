"""Oracle (oracle/audio.py, oracle/signal.py) against the golden vectors generated FROM THE REFERENCE'S OWN CODE
(tests/golden/make_audio_golden.py asserts oracle == reference bit-for-bit before writing them)."""
import os

import torch

from oracle import audio as OA
from oracle import signal as OS

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "audio.pt"))


def test_onsets_rms_peaks_match_reference_vectors():
    y, sr = G["audio_exact"], G["sr"]
    on = OA.onsets(y, sr)[:, 0]
    assert torch.allclose(on, G["onsets"], atol=1e-6)
    assert torch.allclose(OA.rms(y)[:, 0], G["rms"], atol=1e-7)
    assert torch.equal(OA.peak_indices(on), G["peaks"])
    assert float(G["peak_margins"].min()) > 1e-4  # the bit-exact peak test is meaningful: no near-ties


def test_spectra_match_reference_vectors():
    y = G["audio_exact"]
    d = OA.stft(y)
    assert torch.allclose(d.abs()[:, ::16], G["stft_abs"].float(), rtol=2e-3, atol=1e-3)
    perc = OA.hpss(d, margin=8.0)[1]
    assert torch.allclose(perc.abs()[:, ::16], G["perc_abs"].float(), rtol=2e-3, atol=1e-3)


def test_signal_ops_match_reference_vectors():
    env = G["onsets"].clone()
    assert torch.allclose(OS.gaussian_filter(env, 2.0), G["gauss2"], atol=1e-7)
    assert torch.allclose(OS.percentile_clip(env.clone(), 90)[:, 0], G["pclip90"], atol=1e-7)
    assert torch.allclose(OS.resample(env, 57), G["resample57"], atol=1e-7)
    assert torch.allclose(OS.multi_weighted(G["keys"], G["chroma"].clone()), G["multi_weighted"], atol=1e-6)
    assert torch.allclose(OS.slerp_loops(G["keys"], 60, 2), G["slerp_loops"], atol=1e-6)


def test_mel_filterbank_of_the_host_facade_matches_oracle():
    from maua_b200.audiovisual.audioreactive.features import mel_filterbank

    for sr in (24576, 61440):
        assert torch.allclose(mel_filterbank(sr, fmax=11025.0), OA.mel_filterbank(sr, fmax=11025.0), atol=1e-6)
