"""Device audio features (through mb_audio_onsets_rms) against the reference-generated golden vectors and the
oracle; onset-peak frame indices must be BIT-EXACT (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import audio as OA

pytestmark = pytest.mark.gpu
G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "audio.pt"))


def test_against_reference_golden_vectors(cuda):
    from maua_b200.audiovisual import audioreactive as ar

    y, sr = G["audio_exact"].to(cuda), G["sr"]
    on, rms = ar.onsets_rms(y, sr)
    assert on.shape == (len(G["onsets"]), 1) and rms.shape == on.shape
    assert float((on[:, 0].cpu() - G["onsets"]).abs().max()) < 2e-4      # fp32, different FFT / log10 rounding
    assert float((rms[:, 0].cpu() - G["rms"]).abs().max()) < 1e-6
    assert torch.equal(ar.onset_peaks(y, sr).cpu(), G["peaks"])          # bit-exact indices
    perc = ar.percussive(y).cpu()
    ref = OA.percussive(G["audio_exact"])
    assert float((perc - ref).abs().max()) < 1e-4 * float(ref.abs().max()) + 1e-6


@pytest.mark.parametrize("fps,dur", [(24, 30.0), (60, 12.0), (60, 180.0)])
def test_config_sized_sweep_matches_oracle(cuda, fps, dur):
    """BASELINE.json configs[1] audio (30 s @ 24 fps -> 720 frames), a 60 fps variant, and configs[2]'s full length
    (180 s @ 60 fps = 11 059 200 samples -> 10 800 frames), tremolo sweep."""
    from maua_b200.audiovisual import audioreactive as ar
    from maua_b200.workload import sine_sweep

    sr = 1024 * fps
    y48, _ = sine_sweep(dur, tremolo_hz=4.0)
    t = np.arange(int(dur * sr)) / sr
    y = torch.from_numpy(np.interp(t, np.arange(len(y48)) / 48000.0, y48).astype(np.float32))
    # a -40 dB noise floor, as any real recording has: without it the percussive residue of a pure sweep sits at the
    # numerical floor and the comparison measures FFT rounding noise amplified by the dB log, not the algorithm
    y = y + 0.01 * torch.randn(len(y), generator=torch.Generator().manual_seed(7))
    ref_on = OA.onsets(y, sr)[:, 0]
    ref_pk = OA.peak_indices(ref_on)
    on, rms = ar.onsets_rms(y.to(cuda), sr)
    assert on.shape[0] == int(dur * fps)
    err = (on[:, 0].cpu() - ref_on).abs()
    assert float(err.max()) < 1e-3 and float(err.mean()) < 5e-5, (float(err.max()), float(err.mean()))
    assert float((rms[:, 0].cpu() - OA.rms(y)[:, 0]).abs().max()) < 1e-6
    got_pk = ar.onset_peaks(y.to(cuda), sr).cpu()
    # how decisive every comparison of the peak picker is (signal.py:69-76: x[i] > x[i-1] and x[i] > x[i+1]): the smallest
    # |x[i] - x[i+1]| of the envelope against the device's deviation from the oracle
    gaps = (ref_on[1:] - ref_on[:-1]).abs()
    margins = torch.minimum(ref_on[ref_pk] - ref_on[(ref_pk - 1).clamp(0)], ref_on[ref_pk] - ref_on[(ref_pk + 1).clamp(max=len(ref_on) - 1)])
    hist = torch.histc(margins.log10().clamp(-8, 0), bins=8, min=-8, max=0).int().tolist()
    print(f"{fps} fps: {len(ref_pk)} peaks; peak margins per decade 1e-8..1 {hist}; smallest margin {float(margins.min()):.2e}, "
          f"smallest neighbour gap {float(gaps.min()):.2e}; envelope deviation max {float(err.max()):.2e}")
    assert torch.equal(got_pk, ref_pk), (sorted(set(got_pk.tolist()) ^ set(ref_pk.tolist())))   # unconditional: bit-exact indices


def test_errors(cuda):
    from maua_b200.audiovisual import audioreactive as ar

    with pytest.raises(ValueError):
        ar.onsets(torch.zeros(30000, device=cuda), 24576)
    with pytest.raises(RuntimeError):
        ar.onsets(torch.zeros(32768), 24576)


def test_signal_and_latent_ops_against_reference_vectors(cuda):
    """Envelope post-ops / latent sequencers vs vectors produced by the reference's own signal.py / latent.py."""
    from maua_b200.audiovisual import audioreactive as ar
    from oracle import signal as OS

    env = G["onsets"].to(cuda)
    assert float((ar.gaussian_filter(env, 2.0).cpu() - G["gauss2"]).abs().max()) < 1e-6
    assert float((ar.resample(env, 57).cpu() - G["resample57"]).abs().max()) < 1e-6
    assert float((ar.percentile_clip(env.clone(), 90)[:, 0].cpu() - G["pclip90"]).abs().max()) < 1e-6
    assert float((ar.multi_weighted(G["keys"].to(cuda), G["chroma"].to(cuda)).cpu() - G["multi_weighted"]).abs().max()) < 1e-5
    torch.manual_seed(3)
    lat = torch.randn(720, 16, 512)
    ref = OS.gaussian_filter(lat, 2.0, causal=0.3)
    assert float((ar.gaussian_filter(lat.to(cuda), 2.0, causal=0.3).cpu() - ref).abs().max()) < 1e-5
    e = torch.rand(720)
    assert float((ar.single_weighted(lat[0].to(cuda), lat[1].to(cuda), e.to(cuda)).cpu() - OS.single_weighted(lat[0], lat[1], e)).abs().max()) < 1e-6
    assert float((ar.normalize(lat.to(cuda)).cpu() - OS.normalize(lat)).abs().max()) < 1e-6
    assert float((ar.compress(e.to(cuda), 0.5, 0.5).cpu() - OS.compress(e, 0.5, 0.5)).abs().max()) < 1e-6


def test_constant_q_chroma_against_reference_golden_vectors(cuda):
    """chroma_cqt / |cqt| / harmonic on the device vs vectors produced by the reference's own functions."""
    from maua_b200.audiovisual import audioreactive as ar

    y, sr = G["audio_exact"].to(cuda), G["sr"]
    cq = ar.cqt_magnitude(y, sr, n_bins=252, bins_per_octave=36, tuning=0.0).cpu()
    assert cq.shape == G["cqt_abs"].shape
    assert float((cq - G["cqt_abs"]).abs().max()) < 2e-5 * float(G["cqt_abs"].max())     # fp32 FFT / summation order
    ch = ar.chroma_cqt(y, sr, tuning=0.0).cpu()
    assert ch.shape == (12, len(G["onsets"]))
    assert float((ch - G["chroma_cqt"]).abs().max()) < 5e-5 and abs(float(ch.max()) - 1.0) < 1e-6
    harm = ar.harmonic(y)
    ref_h = OA.harmonic(G["audio_exact"])
    assert float((harm.cpu() - ref_h).abs().max()) < 1e-4 * float(ref_h.abs().max()) + 1e-6
    chh = ar.chroma_cqt(harm, sr, tuning=0.0).cpu()     # the chromagram() front half: chroma of the harmonic signal
    assert float((chh - G["chroma_cqt_harmonic"]).abs().max()) < 2e-4


@pytest.mark.parametrize("fps,dur", [(24, 30.0), (60, 6.0)])
def test_config_sized_chroma_matches_oracle(cuda, fps, dur):
    """BASELINE.json configs[1] audio length (n_fft 1024 per octave) and a 60 fps rate (n_fft 2048 per octave)."""
    import warnings

    from maua_b200.audiovisual import audioreactive as ar
    from maua_b200.workload import sine_sweep

    warnings.filterwarnings("ignore")
    sr = 1024 * fps
    y48, _ = sine_sweep(dur, tremolo_hz=4.0)
    t = np.arange(int(dur * sr)) / sr
    y = torch.from_numpy(np.interp(t, np.arange(len(y48)) / 48000.0, y48).astype(np.float32))
    y = y + 0.01 * torch.randn(len(y), generator=torch.Generator().manual_seed(3))
    ref = OA.chroma_cqt(y.clone(), sr, tuning=0.0)
    got = ar.chroma_cqt(y.to(cuda), sr, tuning=0.0).cpu()
    assert got.shape == ref.shape == (12, int(dur * fps))
    assert float((got - ref).abs().max()) < 1e-4


def test_chroma_errors(cuda):
    from maua_b200.audiovisual import audioreactive as ar

    with pytest.raises(RuntimeError):
        ar.chroma_cqt(torch.zeros(1024 * 64), 24576)            # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ar.chroma_cqt(torch.zeros(1000, device=cuda), 24576)
    with pytest.raises(NotImplementedError):
        ar.chroma_cqt(torch.zeros(1024 * 64, device=cuda), 24576, tuning=None)
    with pytest.raises(RuntimeError):
        ar.chroma_cqt(torch.zeros(1024 * 64, device=cuda), 24576, hop_length=32)   # hop not a multiple of 2^6


def test_latent_sequencers_on_device(cuda):
    """slerp_loops / select_modulo vs vectors of the reference's own functions, spline_loops / tempo_loops vs the oracle."""
    from maua_b200.audiovisual import audioreactive as ar
    from oracle import signal as OS

    keys = G["keys"].to(cuda)
    sl = ar.slerp_loops(keys, 60, 2).cpu()
    assert sl.shape == G["slerp_loops"].shape and float((sl - G["slerp_loops"]).abs().max()) < 2e-5
    sm = ar.select_modulo(keys, G["env"].to(cuda), smooth=2).cpu()
    assert sm.shape == G["select_modulo"].shape and float((sm - G["select_modulo"]).abs().max()) < 1e-5
    torch.manual_seed(5)
    big = torch.randn(12, 16, 512)
    sp = ar.spline_loops(big.to(cuda), 720, 4).cpu()
    ref = OS.spline_loops(big, 720, 4)
    assert sp.shape == (720, 16, 512) and float((sp - ref).abs().max()) < 2e-4 * float(ref.abs().max())
    tl = ar.tempo_loops(big.to(cuda), 720, 24, 120.0, type="slerp").cpu()
    assert float((tl - OS.tempo_loops(big, 720, 24, 120.0, type="slerp")).abs().max()) < 5e-5


def test_noise_sequencers_on_device(cuda):
    """Blend / Multiply / Loop and the combinators vs vectors produced by the reference's own classes."""
    from maua_b200.audiovisual.audioreactive import noise as NS

    n = G["noise"]
    mod, i, b = n["mod"].to(cuda), n["i"], n["b"]
    rng = torch.Generator().manual_seed(0)
    nb = NS.Blend(rng, len(mod), (16, 24), mod, noise=n["blend_noise"])
    nm = NS.Multiply(rng, len(mod), (16, 24), mod, noise=n["mult_noise"])
    nl = NS.Loop(rng, len(mod), (16, 24), n_loops=2, sigma=5, noise=n["loop_noise"], device=cuda)
    assert float((nb(i, b).cpu() - n["blend"]).abs().max()) < 1e-5
    assert float((nm(i, b).cpu() - n["multiply"]).abs().max()) < 1e-5
    assert float((nl(i, b).cpu() - n["loop"]).abs().max()) < 2e-4      # sin(cos(.) * 10): fast-math free, fp32 argument rounding
    comb = NS.ScaleBias(NS.Modulate(NS.Average(nb, nl), nm, mod), 0.7, 0.1)
    out = comb(i, b)
    assert out.shape == (b, 16, 24) and float((out.cpu() - n["combined"]).abs().max()) < 2e-4


def test_tuning_and_chromagram_on_device(cuda):
    """estimate_tuning vs the reference's own values (golden), chroma_cqt with that tuning vs the reference's
    tuning=None result, chroma_cens / chromagram vs the oracle."""
    import warnings

    from maua_b200.audiovisual import audioreactive as ar

    warnings.filterwarnings("ignore")
    y, sr = G["audio_exact"], G["sr"]
    tun = float(ar.estimate_tuning(y.to(cuda), sr, bins_per_octave=36))
    assert abs(tun - G["tuning"]) < 1e-6
    harm = OA.harmonic(y)
    tun_h = float(ar.estimate_tuning(harm.to(cuda), sr, bins_per_octave=36))
    assert abs(tun_h - G["tuning_harmonic"]) < 1e-6
    raw = ar.chroma_cqt(harm.to(cuda), sr, tuning=tun_h, norm=False).cpu()
    assert float((raw - G["chroma_cqt_tuned_raw"]).abs().max()) < 2e-5 * float(G["chroma_cqt_tuned_raw"].max())
    ref = OA.chroma_cens(harm, sr)
    got = ar.chroma_cens(harm.to(cuda), sr).cpu()
    assert got.shape == ref.shape == (12, len(G["onsets"]))
    assert float((got - ref).abs().max()) < 2e-3        # the smooth step has slope ~10 per unit: 1e-4 in, 1e-3 out
    cg = ar.chromagram(y.to(cuda), sr).cpu()
    assert cg.shape == (len(G["onsets"]), 12) and float((cg - OA.chromagram(y, sr)).abs().max()) < 3e-3
    assert float((cg.norm(dim=1) - 1).abs().max()) < 1e-5
