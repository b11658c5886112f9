"""Pins oracle/selfsup.py against the REFERENCE's own modules and writes tests/golden/selfsup.pt.

Runs only where /root/reference exists (this container).  Reference modules are imported unmodified with the absent
third-party imports stubbed: torchaudio.functional's biquads are not touched by the functions
pinned here; torchcubicspline is replaced by the oracle's natural spline so that latent_patch / spline_loop_latents are
pinned AROUND the spline (the spline itself is parity-unpinned, see oracle/selfsup.py).
    python tests/golden/make_selfsup_golden.py
"""
import importlib
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
assert os.path.isdir(REF), "the reference checkout is needed to (re)generate the golden vectors"
sys.path.insert(0, REF)

from oracle import selfsup as OS  # noqa: E402


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


class _Spline:
    def __init__(self, coeffs):
        self.t, self.y = coeffs

    def evaluate(self, t_out):  # y is [W, K+1, L] (the reference permutes before and after, latent.py:11-13)
        return OS.natural_spline_eval(self.t, self.y.permute(1, 0, 2), t_out).permute(1, 0, 2)


tcs = types.ModuleType("torchcubicspline")
tcs.natural_cubic_spline_coeffs = lambda t, y: (t, y)
tcs.NaturalCubicSpline = _Spline
sys.modules["torchcubicspline"] = tcs
try:
    import torchaudio  # noqa: F401
except Exception:
    ta = _pkg("torchaudio"); taf = types.ModuleType("torchaudio.functional")
    taf.contrast = taf.highpass_biquad = taf.lowpass_biquad = None
    ta.functional = taf; sys.modules["torchaudio.functional"] = taf
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

rosa = _pkg("librosa"); core = _pkg("librosa.core"); conv = types.ModuleType("librosa.core.convert")
conv.note_to_hz = lambda n: 32.70319566257483
core.convert = conv; rosa.core = core; sys.modules["librosa.core.convert"] = conv
ref_proc = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.processing")
ref_audio = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.audio")
ref_lat = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.latent")


def ref_salience_weighted(envelope, short_sigma=5, long_sigma=80):
    """mir.py:13-21 verbatim in behaviour (mir.py itself imports librosa / scipy.stats / the segmenters at module level)."""
    if envelope.dim() > 1:
        envelope = envelope.squeeze(1)
    short = ref_proc.gaussian_filter(envelope, short_sigma, mode="reflect", causal=0)
    long = ref_proc.gaussian_filter(envelope, long_sigma, mode="reflect", causal=0)
    weighted = (short / long) ** 2 * envelope
    return weighted.unsqueeze(1) if weighted.dim() < 2 else weighted


def same(a, b, name, exact=True):
    assert a.shape == b.shape, (name, a.shape, b.shape)
    ok = torch.equal(a, b) if exact else torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    assert ok, f"oracle != reference for {name}: max diff {(a - b).abs().max()}"


torch.manual_seed(0)
T = 240
env = torch.rand(T).pow(3) + 0.05
envs = torch.rand(T, 3)
lat4 = torch.randn(T, 2, 5, 6)
gold = {"env": env, "envs": envs}
for mode in ("circular", "reflect"):
    for x, nm in ((env, "1d"), (envs, "2d"), (lat4, "4d")):
        same(OS.gaussian_filter(x, 3.0, mode=mode), ref_proc.gaussian_filter(x, 3.0, mode=mode), f"gaussian_filter {mode} {nm}")
same(OS.normalize(envs), ref_proc.normalize(envs), "normalize")
same(OS.salience_weighted(env, 5, 40), ref_salience_weighted(env, 5, 40), "salience_weighted")
same(OS.clamp_peaks_percentile(envs, 90), ref_proc.clamp_peaks_percentile(envs, 90), "clamp_peaks_percentile")
same(OS.emphasize(envs, 2.0, 75), ref_proc.emphasize(envs, 2.0, 75), "emphasize")
gold["salience"] = OS.salience_weighted(env, 5, 40)
gold["gauss_reflect"] = OS.gaussian_filter(envs, 3.0, mode="reflect")
gold["clamp_peaks"] = OS.clamp_peaks_percentile(envs, 90)
gold["emphasize"] = OS.emphasize(envs, 2.0, 75)

# drop_strength / tonnetz of features/audio.py with their inputs injected (rms and the chromagram are pinned elsewhere)
rms_env = torch.rand(T, 1)
chroma = torch.rand(12, T) + 0.01
_rms = ref_audio.rms
ref_audio.rms = lambda audio, sr: rms_env
same(OS.drop_strength_from_rms(rms_env), ref_audio.drop_strength(None, 0), "drop_strength")
ref_audio.rms = _rms
same(OS.tonnetz_from_chroma(chroma), ref_audio.tonnetz(torch.zeros(1), 0, chroma_fn=lambda a, sr: chroma), "tonnetz")
same(OS.gaussian_filter(rms_env, 3.0), ref_proc.gaussian_filter(rms_env, 3.0), "gaussian_filter [T,1] -> [T]")
gold["rms_env"], gold["chroma"] = rms_env, chroma
gold["drop_strength"], gold["tonnetz"] = OS.drop_strength_from_rms(rms_env), OS.tonnetz_from_chroma(chroma)

# spectral descriptors of features/audio.py:59-133 on a 2 s test signal at sr = 1024 * 24
from oracle import audio as OA  # noqa: E402
sr = 1024 * 24
tt = torch.arange(2 * sr) / sr
sig = 0.5 * torch.sin(2 * torch.pi * (200 * tt + 900 * tt ** 2)) * (0.6 + 0.4 * torch.sin(2 * torch.pi * 3 * tt)) + 0.05 * torch.randn(2 * sr)
same(OA.mfcc(sig, sr), ref_audio.mfcc(sig, sr), "mfcc")
same(OA.spectral_contrast(sig, sr), ref_audio.spectral_contrast(sig, sr), "spectral_contrast")
same(OA.spectral_flatness(sig, sr), ref_audio.spectral_flatness(sig, sr), "spectral_flatness")
gold["spec_signal"], gold["spec_sr"] = sig, sr
gold["mfcc"], gold["spectral_contrast"], gold["spectral_flatness"] = OA.mfcc(sig, sr), OA.spectral_contrast(sig, sr), OA.spectral_flatness(sig, sr)

palette = torch.randn(12, 18, 16)
same(OS.spline_loop_latents(palette[:5], T, 2.5), ref_lat.spline_loop_latents(palette[:5], T, 2.5), "spline_loop_latents")
gold["palette"] = palette
gold["spline_loop"] = OS.spline_loop_latents(palette[:5], T, 2.5)

features = {"onsets": torch.rand(T, 1), "chromagram": torch.rand(T, 12), "rms": torch.rand(T, 1)}
segmentations = {("onsets", 4): torch.randint(0, 4, (T,)), ("chromagram", 4): torch.randint(0, 4, (T,)), ("rms", 4): torch.randint(0, 4, (T,))}
gold["features"], gold["segmentations"] = features, segmentations
cases = [
    dict(patch_type="segmentation", seq_feat="onsets", merge_type="average", merge_depth="low"),
    dict(patch_type="feature", seq_feat="onsets", merge_type="modulate", merge_depth="mid"),
    dict(patch_type="feature", seq_feat="chromagram", merge_type="overwrite", merge_depth="high"),
    dict(patch_type="loop", seq_feat="rms", merge_type="modulate", merge_depth="lowmid"),
    dict(patch_type="loop", seq_feat="rms", merge_type="average", merge_depth="all"),
]
gold["latent_patch"] = []
base_lat = torch.randn(T, 18, 16)
gold["base_latents"] = base_lat
for i, c in enumerate(cases):
    kw = dict(palette=palette, segmentations=segmentations, features=features, tempo=120.0, fps=24, segments=4, loop_bars=4,
              seq_feat_weight=0.8, mod_feat="rms", mod_feat_weight=0.6, **c)
    a = OS.latent_patch(torch.Generator().manual_seed(10 + i), base_lat.clone(), **kw)
    b = ref_lat.latent_patch(torch.Generator().manual_seed(10 + i), base_lat.clone(), **kw)
    same(a, b, f"latent_patch {c}")
    perm = torch.randperm(len(palette), generator=torch.Generator().manual_seed(10 + i))
    gold["latent_patch"].append(dict(case=c, seed=10 + i, permutation=perm, out=a))
torch.save(gold, os.path.join(ROOT, "tests", "golden", "selfsup.pt"))
print("oracle/selfsup.py == reference on every pinned function; wrote tests/golden/selfsup.pt")
