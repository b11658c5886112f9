"""Device twins of the torch-native patch functions (maua_b200/audiovisual/audioreactive/selfsupervised.py) against the
CPU oracle (oracle/selfsup.py, pinned against the reference by tests/golden/make_selfsup_golden.py) on the fixture inputs.
fp32 both sides; tolerance 1e-5 absolute on O(1) values (summation order of the 1-D filters differs)."""
import os

import pytest
import torch

from oracle import selfsup as OS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "selfsup.pt")
TOL = 1e-5


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def close(a, b, tol=TOL):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float((a.cpu() - b).abs().max())
    assert err <= tol, err


def test_envelope_ops(cuda, gold):
    from maua_b200.audiovisual.audioreactive import selfsupervised as S

    env, envs = gold["env"].to(cuda), gold["envs"].to(cuda)
    for mode in ("circular", "reflect"):
        close(S.gaussian_filter(env, 3.0, mode=mode), OS.gaussian_filter(gold["env"], 3.0, mode=mode))
        close(S.gaussian_filter(envs, 3.0, mode=mode), OS.gaussian_filter(gold["envs"], 3.0, mode=mode))
    x4 = torch.randn(64, 2, 5, 6)
    close(S.gaussian_filter(x4.to(cuda), 2.0), OS.gaussian_filter(x4, 2.0))
    close(S.normalize(envs), OS.normalize(gold["envs"]))
    close(S.salience_weighted(env, 5, 40), gold["salience"], 1e-4)   # (short / long)^2 amplifies the filter rounding
    close(S.salience_weighted(env[:, None], 5, 40), gold["salience"], 1e-4)
    close(S.clamp_peaks_percentile(envs, 90), gold["clamp_peaks"])
    close(S.emphasize(envs, 2.0, 75), gold["emphasize"])
    close(S.clamp_upper_percentile(envs, 80), torch.clamp(gold["envs"], None, torch.quantile(gold["envs"], 0.8, dim=0)))
    close(S.clamp_lower_percentile(envs, 20), torch.clamp(gold["envs"], torch.quantile(gold["envs"], 0.2, dim=0), None))
    with pytest.raises(RuntimeError):
        S.gaussian_filter(gold["env"], 3.0)   # CPU tensor: no fallback
    close(S.gaussian_filter(gold["rms_env"].to(cuda), 3.0), OS.gaussian_filter(gold["rms_env"], 3.0))   # [T,1] -> [T]


def test_drop_strength_and_tonnetz(cuda, gold, monkeypatch):
    from maua_b200.audiovisual.audioreactive import features as F
    from maua_b200.audiovisual.audioreactive import selfsupervised as S

    monkeypatch.setattr(F, "rms", lambda audio, sr: gold["rms_env"].to(cuda))   # the device rms is pinned in test_audio_gpu.py
    close(S.drop_strength(None, 0), gold["drop_strength"], 2e-5)
    close(S.tonnetz(None, 0, chroma_fn=lambda a, sr: gold["chroma"].to(cuda)), gold["tonnetz"], 1e-5)


def test_spline_loop_latents(cuda, gold):
    from maua_b200.audiovisual.audioreactive import selfsupervised as S

    T = len(gold["base_latents"])
    close(S.spline_loop_latents(gold["palette"][:5].to(cuda), T, 2.5), gold["spline_loop"], 2e-5)
    for n_loops in (1, 3, 0.37):
        close(S.spline_loop_latents(gold["palette"][:4].to(cuda), 97, n_loops), OS.spline_loop_latents(gold["palette"][:4], 97, n_loops), 2e-5)


def test_latent_patch(cuda, gold):
    from maua_b200.audiovisual.audioreactive import selfsupervised as S

    feats = {k: v.to(cuda) for k, v in gold["features"].items()}
    segs = {k: v.to(cuda) for k, v in gold["segmentations"].items()}
    for rec in gold["latent_patch"]:
        kw = dict(palette=gold["palette"].to(cuda), segmentations=segs, features=feats, tempo=120.0, fps=24, segments=4, loop_bars=4,
                  seq_feat_weight=0.8, mod_feat="rms", mod_feat_weight=0.6, **rec["case"])
        lat = gold["base_latents"].to(cuda).clone()
        out = S.latent_patch(torch.Generator().manual_seed(rec["seed"]), lat, **kw)   # CPU generator: same permutation as the fixture
        assert out.data_ptr() == lat.data_ptr()
        close(out, rec["out"], 3e-5)


def test_noise_patch(cuda, gold):
    """noise.py:89-140: the wrapped sequencers evaluate to scale * merge(old, new) + bias with the oracle's formulas."""
    from maua_b200.audiovisual.audioreactive import noise as N
    from maua_b200.audiovisual.audioreactive import selfsupervised as S
    from oracle import noise as ON

    T = len(gold["env"])
    feats = {k: v.to(cuda) for k, v in gold["features"].items()}
    rng = torch.Generator(device="cuda").manual_seed(3)
    base = [N.Loop(rng=rng, length=T, size=(8, 8), n_loops=2) for _ in range(17)]
    before = [b(0, 4).clone() for b in base]
    rng2 = torch.Generator(device="cuda").manual_seed(4)
    out = S.noise_patch(rng2, list(base), feats, tempo=120.0, fps=24, patch_type="multiply", loop_bars=4, seq_feat="chromagram",
                        seq_feat_weight=0.5, mod_feat="rms", mod_feat_weight=0.7, merge_type="modulate", merge_depth="mid",
                        noise_mean=0.1, noise_std=2.0)
    for n in range(17):
        if n not in range(6, 12):
            assert out[n] is base[n]
            continue
        sb = out[n]
        assert isinstance(sb, N.ScaleBias) and isinstance(sb.base, N.Modulate) and isinstance(sb.base.right, N.Multiply)
        new = ON.multiply(sb.base.right.noise.cpu(), sb.base.right.modulator.cpu(), 0, 4)
        mod_mean = (0.7 * gold["features"]["rms"]).mean(1)
        want = ON.scale_bias(ON.modulate(before[n].cpu(), new, mod_mean, 0, 4), 2.0, 0.1)
        close(sb(0, 4), want, 1e-4)


def test_spectral_descriptors(cuda, gold):
    """mfcc / spectral_contrast / spectral_flatness on the device against the pinned oracle values.  The device FFT and the
    torch CPU FFT differ in rounding: dB-domain features are compared at 2e-3 dB-scale absolute, flatness relatively."""
    from maua_b200.audiovisual.audioreactive import features as F

    sig, sr = gold["spec_signal"].to(cuda), gold["spec_sr"]
    m = F.mfcc(sig, sr)
    assert m.shape == (48, 20)
    close(m, gold["mfcc"], 5e-3)                      # coefficients are O(100): 5e-5 relative
    c = F.spectral_contrast(sig, sr)
    assert c.shape == (48, 7)
    close(c, gold["spectral_contrast"], 5e-3)         # dB differences, O(10)
    f = F.spectral_flatness(sig, sr)
    assert f.shape == (48, 1)
    rel = float(((f.cpu() - gold["spectral_flatness"]).abs() / gold["spectral_flatness"]).max())
    assert rel < 1e-3, rel
    lin = F.spectral_contrast(sig, sr, linear=True)
    from oracle import audio as OA
    close(lin, OA.spectral_contrast(gold["spec_signal"], sr, linear=True), 1e-3)


def test_extract_features_end_to_end(cuda, gold):
    """All eight AFEATFNS (mir.py:9) on the device for a 16 s signal, post-processed as retrieve_music_information does
    (long_sigma 80 needs more than 320 frames); the chain against the oracle composition on the device's own raw features."""
    from maua_b200.audiovisual.audioreactive import selfsupervised as S

    sr = gold["spec_sr"]
    sig = torch.cat([gold["spec_signal"], gold["spec_signal"].flip(0)] * 4).to(cuda)   # 16 s = 384 frames
    raw = S.extract_features(sig, sr, postprocess=False)
    assert list(raw) == ["chromagram", "tonnetz", "mfcc", "spectral_contrast", "spectral_flatness", "rms", "drop_strength", "onsets"]
    assert {k: tuple(v.shape) for k, v in raw.items()} == {
        "chromagram": (384, 12), "tonnetz": (384, 6), "mfcc": (384, 20), "spectral_contrast": (384, 7), "spectral_flatness": (384, 1),
        "rms": (384, 1), "drop_strength": (384, 1), "onsets": (384, 1)}
    post = S.extract_features(sig, sr)
    for k, v in raw.items():
        assert torch.isfinite(post[k]).all(), k
        want = OS.normalize(OS.salience_weighted(OS.gaussian_filter(v.cpu(), sigma=2)))
        close(post[k].reshape(want.shape), want, 2e-4)
        assert float(post[k].min()) >= 0.0 and float(post[k].max()) <= 1.0 + 1e-6


def test_midpoint_quantile_on_the_device(cuda):
    """quantile / standardize / onset_envelope (efficient_quantile + processing.py:58-61,93-98): the device radix select
    against vectors from the reference's own compiled routine -- bit for bit (order statistics are exact, the mid point is
    formed in double as the reference does)."""
    import os

    from maua_b200.audiovisual.audioreactive import selfsupervised as ss
    from oracle import quantile as OQ

    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "quantile.pt"))
    for name, c in gold["cases"].items():
        x = c["x"].to(cuda)
        got = torch.stack([ss.quantile(x, q) for q in c["qs"]]).cpu()
        assert torch.equal(got, c["want"]), (name, got, c["want"])
    assert torch.equal(ss.standardize(gold["flow"].to(cuda)).cpu(), gold["standardize"])
    flux = ss.spectral_flux(gold["spec"].to(cuda))
    assert torch.equal(flux.cpu(), OQ.spectral_flux(gold["spec"]))
    env = ss.onset_envelope(flux).cpu()
    # the sum over bins is a device reduction (different order than the host's): values to 1e-6, the clamp bounds exact
    assert float((env - gold["onset_envelope"]).abs().max()) < 1e-6
    g = torch.Generator().manual_seed(8)
    big = torch.randn(300001, generator=g)
    big[::1000] = float("nan")
    for q in (0.0015, 0.5, 0.9985):
        assert torch.equal(ss.quantile(big.to(cuda), q).cpu(), OQ.quantile(big, q)), q
    with pytest.raises(RuntimeError):
        ss.quantile(big, 0.5)   # host tensor: no CPU fallback
