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


def test_constant_q_chroma_matches_reference_vectors():
    """oracle chroma_cqt / cqt / harmonic vs the vectors the REFERENCE's own functions produced (make_audio_golden.py)."""
    import warnings

    warnings.filterwarnings("ignore")
    y, sr = G["audio_exact"], G["sr"]
    assert torch.equal(OA.chroma_cqt(y.clone(), sr, tuning=0.0), G["chroma_cqt"])
    assert torch.equal(OA.cqt(y.clone(), sr, n_bins=252, bins_per_octave=36, tuning=0.0).abs(), G["cqt_abs"])
    assert torch.equal(OA.chroma_cqt(OA.harmonic(y), sr, tuning=0.0), G["chroma_cqt_harmonic"])


def test_chroma_design_of_the_host_facade_matches_oracle_and_torchaudio():
    """The product's host-side constant design (decimation taps, sparse filter bank, fold matrix) is the reference's."""
    from torchaudio.functional import functional as TF

    from maua_b200.audiovisual.audioreactive import chroma as CH

    k, w = CH.kaiser_decimation_kernel()
    rk, rw = TF._get_sinc_resample_kernel(2, 1, 1, 6, 0.99, "sinc_interp_kaiser", None)
    assert w == rw and torch.equal(k, rk.reshape(-1))
    for sr in (24 * 1024, 60 * 1024):
        f0 = torch.tensor(CH.C1_HZ).float()
        top = (f0 * 2.0 ** (torch.arange(0, 252, dtype=torch.float) / 36))[-36:]
        rowptr, col, val, n_fft = CH.octave_filter_bank(sr, torch.min(top), 36)
        dense, keep, nf = OA.cqt_filter_fft(sr, torch.min(top), 36, 36)
        assert n_fft == nf and int(rowptr[-1]) == int(keep.sum()) and torch.equal(val, dense[keep])
        assert torch.equal(col.long(), keep.nonzero()[:, 1])
    assert torch.equal(CH.cq_to_chroma(252, 36, 12), OA.cq_to_chroma(252, 36, 12))


def test_sequencers_match_reference_vectors_and_scipy():
    """select_modulo / noise sequencers vs vectors of the reference's own code; the natural spline (third-party in the
    reference, absent here) vs scipy's CubicSpline(bc_type="natural")."""
    import numpy as np
    from scipy.interpolate import CubicSpline

    from oracle import noise as ON

    assert torch.equal(OS.select_modulo(G["keys"], G["env"].clone(), smooth=2), G["select_modulo"])
    n = G["noise"]
    assert torch.equal(ON.blend(n["blend_noise"], n["mod"], n["i"], n["b"]), n["blend"])
    assert torch.equal(ON.multiply(n["mult_noise"], n["mod"], n["i"], n["b"]), n["multiply"])
    assert torch.equal(ON.loop(n["loop_noise"], n["loop_idx"], 5, n["i"], n["b"]), n["loop"])
    keys, size, n_loops = G["keys"], 97, 3
    got = OS.spline_loops(keys, size, n_loops)
    y = torch.cat([keys] * n_loops + [keys[[0]]]).reshape(len(keys) * n_loops + 1, -1).double().numpy()
    ref = CubicSpline(np.linspace(0, 1, len(y)), y, bc_type="natural")(np.linspace(0, 1, size))
    assert got.shape == (size, 4, 8)
    assert float(np.abs(got.reshape(size, -1).numpy() - ref).max()) < 1e-5


def test_tuning_estimate_and_cens_quantiser():
    """estimate_tuning vs the reference's own value; the CENS quantiser spline (third-party in the reference) vs scipy's
    natural CubicSpline; the product's host design equals the oracle's."""
    import numpy as np
    from scipy.interpolate import CubicSpline

    from maua_b200.audiovisual.audioreactive import chroma as CH

    y, sr = G["audio_exact"], G["sr"]
    assert float(OA.estimate_tuning(y, sr, bins_per_octave=36)) == G["tuning"]
    xs, ys = OA.cens_quantiser_knots()
    x, a, b, c, d = OA.natural_cubic_coeffs(xs.numpy(), ys.numpy())
    ref = CubicSpline(xs.double().numpy(), ys.double().numpy(), bc_type="natural")
    t = np.linspace(-0.09, 1.09, 4001)
    i = np.clip(np.searchsorted(x, t, side="left") - 1, 0, len(a) - 1)
    f = t - x[i]
    assert np.abs(a[i] + (b[i] + (c[i] + d[i] * f) * f) * f - ref(t)).max() < 1e-9
    kx, coef = CH.cens_quantiser_design()
    assert torch.equal(kx, xs.float())
    assert np.abs(coef.numpy() - np.stack([a, b, c, d]).astype(np.float32)).max() == 0.0
    q = OA.spline_quantize(torch.tensor([[0.0, 0.03, 0.07, 0.15, 0.3, 0.6, 1.0]]))
    assert torch.allclose(q, torch.tensor([[0.0, 0.0, 0.25, 0.5, 0.75, 1.0, 1.0]]), atol=2e-3)   # the 4-step CENS staircase
