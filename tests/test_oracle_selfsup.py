"""oracle/selfsup.py against the committed fixture (tests/golden/selfsup.pt, written by make_selfsup_golden.py after
checking every function against the reference's own modules) and its natural spline against scipy."""
import os

import numpy as np
import pytest
import torch

from oracle import selfsup as OS

GOLD = os.path.join(os.path.dirname(__file__), "golden", "selfsup.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def test_envelope_ops_match_fixture(gold):
    assert torch.equal(OS.salience_weighted(gold["env"], 5, 40), gold["salience"])
    assert torch.equal(OS.gaussian_filter(gold["envs"], 3.0, mode="reflect"), gold["gauss_reflect"])
    assert torch.equal(OS.clamp_peaks_percentile(gold["envs"], 90), gold["clamp_peaks"])
    assert torch.equal(OS.emphasize(gold["envs"], 2.0, 75), gold["emphasize"])
    assert torch.equal(OS.drop_strength_from_rms(gold["rms_env"]), gold["drop_strength"])
    assert torch.equal(OS.tonnetz_from_chroma(gold["chroma"]), gold["tonnetz"])
    assert OS.gaussian_filter(gold["rms_env"], 3.0).shape == (len(gold["rms_env"]),)   # the reference squeezes [T,1] to [T]


def test_latent_patch_matches_fixture(gold):
    T = len(gold["base_latents"])
    assert torch.equal(OS.spline_loop_latents(gold["palette"][:5], T, 2.5), gold["spline_loop"])
    for rec in gold["latent_patch"]:
        kw = dict(palette=gold["palette"], segmentations=gold["segmentations"], features=gold["features"], tempo=120.0, fps=24,
                  segments=4, loop_bars=4, seq_feat_weight=0.8, mod_feat="rms", mod_feat_weight=0.6, **rec["case"])
        out = OS.latent_patch(torch.Generator().manual_seed(rec["seed"]), gold["base_latents"].clone(), **kw)
        assert torch.equal(out, rec["out"]), rec["case"]


def test_natural_spline_matches_scipy():
    from scipy.interpolate import CubicSpline

    torch.manual_seed(1)
    y = torch.randn(7, 3, 4)
    t_in = torch.linspace(0, 1, 7)
    t_out = torch.linspace(0, 3.3, 101) % 1
    got = OS.natural_spline_eval(t_in, y, t_out)
    want = CubicSpline(t_in.double().numpy(), y.reshape(7, -1).double().numpy(), bc_type="natural")(t_out.double().numpy())
    assert np.allclose(got.reshape(101, -1).numpy(), want, atol=1e-5)


def test_spectral_descriptors_match_fixture(gold):
    from oracle import audio as OA

    sig, sr = gold["spec_signal"], gold["spec_sr"]
    assert torch.equal(OA.mfcc(sig, sr), gold["mfcc"]) and gold["mfcc"].shape == (48, 20)
    assert torch.equal(OA.spectral_contrast(sig, sr), gold["spectral_contrast"]) and gold["spectral_contrast"].shape == (48, 7)
    assert torch.equal(OA.spectral_flatness(sig, sr), gold["spectral_flatness"]) and gold["spectral_flatness"].shape == (48, 1)


def test_oracle_dct_is_the_orthonormal_dct2():
    from scipy.fft import dct as sdct

    from oracle import audio as OA

    x = torch.randn(5, 128)
    assert np.allclose(OA.dct(x, norm="ortho").numpy(), sdct(x.numpy(), type=2, norm="ortho", axis=-1), atol=1e-4)


def test_contrast_band_design_matches_the_reference_loop():
    """Host band design of the device kernel (features.contrast_bands) selects the bins the oracle's boolean masks select."""
    from maua_b200.audiovisual.audioreactive.features import contrast_bands

    sr, n_bands, fmin, quantile = 1024 * 24, 6, 200.0, 0.02
    lo, hi, cnt = contrast_bands(sr)
    freq = torch.linspace(0, float(sr) / 2, 1025)
    octa = torch.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** torch.arange(0, n_bands + 1))
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        band = torch.logical_and(freq >= f_low, freq <= f_high)
        idx = band.flatten().nonzero()
        if k > 0:
            band[idx[0] - 1] = True
        if k == n_bands:
            band[idx[-1] + 1:] = True
        bins = band.nonzero().flatten().tolist()
        if k < n_bands:
            bins = bins[:-1]
        assert bins == list(range(lo[k], hi[k])), k
        assert cnt[k] == int(torch.maximum(torch.round(quantile * torch.sum(band)), torch.ones(())))
    assert hi[-1] == 1025
