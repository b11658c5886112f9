"""Pins oracle/audio.py and oracle/signal.py against the REFERENCE's own modules and writes tests/golden/audio.pt.

Runs only where /root/reference exists (this container).  The reference modules are imported unmodified; three
third-party imports they make at module level are absent in this image and are stubbed exactly as SURVEY
Appendix C.2 describes (librosa.note_to_hz, torchcubicspline, the compiled efficient_quantile, torchtyping).
Every oracle function is compared with the reference function on the same input (must be bit-identical: both
sides run the same torch CPU kernels) before the fixture is written.
    python tests/golden/make_audio_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
assert os.path.isdir(REF), "the reference checkout is needed to (re)generate the audio golden vectors"
sys.path.insert(0, REF)


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


# stubs for absent third-party modules
rosa = _pkg("librosa"); core = _pkg("librosa.core"); conv = types.ModuleType("librosa.core.convert")
conv.note_to_hz = lambda n: 32.70319566257483
core.convert = conv; rosa.core = core; sys.modules["librosa.core.convert"] = conv
tcs = types.ModuleType("torchcubicspline")
tcs.natural_cubic_spline_coeffs = lambda *a, **k: None
tcs.NaturalCubicSpline = object
sys.modules["torchcubicspline"] = tcs
tt = types.ModuleType("torchtyping")
class _TT:
    def __class_getitem__(cls, item):
        return torch.Tensor
tt.TensorType = _TT
sys.modules["torchtyping"] = tt
# package shells so the heavy package __init__ files (librosa/madmom/openunmix imports) are skipped
base = REF + "/maua/audiovisual/audioreactive"
_pkg("maua", REF + "/maua"); _pkg("maua.audiovisual", REF + "/maua/audiovisual")
_pkg("maua.audiovisual.audioreactive", base)
_pkg("maua.audiovisual.audioreactive.selfsupervised", base + "/selfsupervised")
_pkg("maua.audiovisual.audioreactive.selfsupervised.features", base + "/selfsupervised/features")
eq = types.ModuleType("maua.audiovisual.audioreactive.selfsupervised.features.efficient_quantile")
# the reference's own compiled efficient_quantile.cpp (oracle/_ref, built by oracle/build_ref.py) behind the wrapper of
# efficient_quantile/__init__.py:6-7 -- not a torch.quantile stand-in (that one interpolates linearly, the reference takes the
# mid point with a float32 q)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.build_ref import load_efficient_quantile  # noqa: E402

_eq = load_efficient_quantile()
assert _eq is not None, "run `python oracle/build_ref.py` first"
eq.quantile = lambda tensor, q: _eq(tensor.cpu().flatten(), torch.FloatTensor([q]), True, 3).squeeze().to(tensor.device)
sys.modules[eq.__name__] = eq

ref_audio = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.audio")
ref_spec = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.rosa.spectral")
ref_beat = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.rosa.beat")
ref_pitch = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.rosa.pitch")
ref_cq = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.rosa.constantq")
ref_signal = importlib.import_module("maua.audiovisual.audioreactive.signal")
ref_latent = importlib.import_module("maua.audiovisual.audioreactive.latent")
ref_noise = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.noise")

from maua_b200.workload import sine_sweep  # noqa: E402
from oracle import audio as OA, noise as ON, signal as OS  # noqa: E402

torch.manual_seed(0)
fps, dur = 24, 4.0
sr = 1024 * fps
y48, _ = sine_sweep(dur, tremolo_hz=4.0)
# resample 48 kHz -> 1024*fps by linear interpolation (the fixture only needs a deterministic band-limited signal)
t = np.arange(int(dur * sr)) / sr
y = torch.from_numpy(np.interp(t, np.arange(len(y48)) / 48000.0, y48).astype(np.float32))
y = y + 0.05 * torch.randn(len(y))


def same(a, b, name):
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert torch.equal(a, b), f"oracle != reference for {name}: max diff {(a - b).abs().max()}"


with torch.inference_mode():
    d = ref_spec.stft(y)
    same(OA.stft(y), d, "stft")
    same(OA.spectrogram(y, 2.0), ref_spec.spectrogram(y, power=2.0), "spectrogram")
    same(OA.mel_filterbank(sr, fmax=11025.0), ref_spec.mel(sr, 2048, fmax=11025.0), "mel")
    hr, pr = ref_spec.hpss(d, margin=8.0)
    ho, po = OA.hpss(d, margin=8.0)
    same(ho, hr, "hpss harmonic"); same(po, pr, "hpss percussive")
    same(OA.percussive(y), ref_audio.percussive(y), "percussive")
    same(OA.onset_strength(OA.percussive(y), sr), ref_beat.onset_strength(ref_audio.percussive(y), sr), "onset_strength")
    on_ref = ref_audio.onsets(y, sr)
    same(OA.onsets(y, sr), on_ref, "onsets")
    rms_ref = ref_audio.rms(y, sr)
    same(OA.rms(y), rms_ref, "rms")
    pulse_ref = ref_audio.pulse(y, sr)
    same(OA.pulse(y, sr), pulse_ref, "pulse (plp)")

    import warnings
    warnings.filterwarnings("ignore")  # torchaudio's deprecation notice for the "kaiser_window" method name
    harm_ref = ref_audio.harmonic(y)
    same(OA.harmonic(y), harm_ref, "harmonic")
    cq_ref = ref_cq.cqt(y.clone(), sr, n_bins=252, bins_per_octave=36, tuning=0.0)
    same(OA.cqt(y.clone(), sr, n_bins=252, bins_per_octave=36, tuning=0.0), cq_ref, "cqt")
    chroma_ref = ref_spec.chroma_cqt(y.clone(), sr, tuning=0.0)
    same(OA.chroma_cqt(y.clone(), sr, tuning=0.0), chroma_ref, "chroma_cqt")
    chroma_h_ref = ref_spec.chroma_cqt(harm_ref.clone(), sr, tuning=0.0)
    same(OA.chroma_cqt(OA.harmonic(y), sr, tuning=0.0), chroma_h_ref, "chroma_cqt(harmonic)")

    pit_ref, mag_ref = ref_pitch.piptrack(y, sr)
    pit_o, mag_o = OA.piptrack(y, sr)
    same(pit_o, pit_ref, "piptrack pitches"); same(mag_o, mag_ref, "piptrack mags")
    tun_ref = ref_pitch.estimate_tuning(y, sr, bins_per_octave=36)
    same(torch.as_tensor(OA.estimate_tuning(y, sr, bins_per_octave=36)), torch.as_tensor(tun_ref), "estimate_tuning")
    tun_h_ref = ref_pitch.estimate_tuning(harm_ref, sr, bins_per_octave=36)
    same(torch.as_tensor(OA.estimate_tuning(harm_ref, sr, bins_per_octave=36)), torch.as_tensor(tun_h_ref), "estimate_tuning(harmonic)")
    chroma_tuned_ref = ref_spec.chroma_cqt(harm_ref.clone(), sr, tuning=None, norm=False)   # the reference's own tuning default
    same(OA.chroma_cqt(harm_ref.clone(), sr, tuning=float(tun_h_ref), norm=False), chroma_tuned_ref, "chroma_cqt(tuning=None)")

    env = on_ref[:, 0].clone()
    same(OS.normalize(env), ref_signal.normalize(env), "signal.normalize")
    same(OS.resample(env, 57), ref_signal.resample(env, 57), "signal.resample")
    same(OS.gaussian_filter(env, 2.0), ref_signal.gaussian_filter(env, 2.0), "gaussian_filter 1d")
    lat = torch.randn(len(env), 4, 8)
    same(OS.gaussian_filter(lat, 3.0, causal=0.3), ref_signal.gaussian_filter(lat, 3.0, causal=0.3), "gaussian_filter 3d causal")
    same(OS.percentile_clip(env.clone(), 90), ref_signal.percentile_clip(env.clone(), 90), "percentile_clip")
    same(OS.compress(env.clone(), 0.5, 0.5), ref_signal.compress(env.clone(), 0.5, 0.5), "compress")
    keys = torch.randn(5, 4, 8)
    same(OS.single_weighted(keys[0], keys[1], env), ref_latent.single_weighted(keys[0], keys[1], env), "single_weighted")
    chroma = torch.rand(len(env), 5)
    mw_ref = ref_latent.multi_weighted(keys, chroma.clone())
    assert torch.allclose(OS.multi_weighted(keys, chroma.clone()), mw_ref, atol=1e-6), "multi_weighted"
    same(OS.slerp_loops(keys, 60, 2), ref_latent.slerp_loops(keys, 60, 2), "slerp_loops")

    sel_ref = ref_latent.select_modulo(keys, env.clone(), smooth=2)
    same(OS.select_modulo(keys, env.clone(), smooth=2), sel_ref, "select_modulo")

    # noise sequencers: the reference's own classes on a seeded CPU generator
    rng = torch.Generator().manual_seed(7)
    mod = torch.rand(len(env), 3, generator=rng)
    nb = ref_noise.Blend(rng, len(env), (16, 24), mod)
    nm = ref_noise.Multiply(rng, len(env), (16, 24), mod)
    nl = ref_noise.Loop(rng, len(env), (16, 24), n_loops=2, sigma=5)
    i0, bsz = 5, 7
    same(ON.blend(nb.noise, mod, i0, bsz), nb(i0, bsz), "noise Blend")
    same(ON.multiply(nm.noise, mod, i0, bsz), nm(i0, bsz), "noise Multiply")
    same(ON.loop(nl.noise, nl.idx, nl.sigma, i0, bsz), nl(i0, bsz), "noise Loop")
    comb = ref_noise.ScaleBias(ref_noise.Modulate(ref_noise.Average(nb, nl), nm, mod), 0.7, 0.1)
    comb_ref = comb(i0, bsz)
    same(ON.scale_bias(ON.modulate(ON.average(nb(i0, bsz), nl(i0, bsz)), nm(i0, bsz), mod.mean(1), i0, bsz), 0.7, 0.1), comb_ref, "noise combinators")

    peaks = OA.peak_indices(on_ref)
    margins = torch.minimum(on_ref[peaks, 0] - on_ref[(peaks - 1).clamp(0), 0], on_ref[peaks, 0] - on_ref[(peaks + 1).clamp(max=len(on_ref) - 1), 0])
    out = dict(sr=sr, fps=fps, audio=y.half(), audio_exact=y, stft_abs=d.abs()[:, ::16].half(), perc_abs=pr.abs()[:, ::16].half(),
               onsets=on_ref[:, 0], rms=rms_ref[:, 0], pulse=pulse_ref[:, 0], peaks=peaks, peak_margins=margins,
               gauss2=ref_signal.gaussian_filter(env, 2.0), pclip90=ref_signal.percentile_clip(env.clone(), 90)[:, 0],
               resample57=ref_signal.resample(env, 57), lat=lat[:, :2, :4].clone(), lat_gauss=ref_signal.gaussian_filter(lat, 3.0, causal=0.3)[:, :2, :4].clone(),
               harmonic=harm_ref.half(), cqt_abs=cq_ref.abs(), chroma_cqt=chroma_ref, chroma_cqt_harmonic=chroma_h_ref,
               select_modulo=sel_ref, env=env.clone(),
               noise=dict(mod=mod, blend_noise=nb.noise.clone(), mult_noise=nm.noise.clone(), loop_noise=nl.noise.clone(), loop_idx=nl.idx.clone(),
                          i=i0, b=bsz, blend=nb(i0, bsz), multiply=nm(i0, bsz), loop=nl(i0, bsz), combined=comb_ref),
               tuning=float(tun_ref), tuning_harmonic=float(tun_h_ref), chroma_cqt_tuned_raw=chroma_tuned_ref,
               keys=keys, chroma=chroma, multi_weighted=mw_ref, slerp_loops=ref_latent.slerp_loops(keys, 60, 2))
torch.save(out, os.path.join(ROOT, "tests", "golden", "audio.pt"))
print("oracle == reference on every pinned function; wrote tests/golden/audio.pt",
      {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})
print("peaks", len(peaks), "min margin", float(margins.min()))
