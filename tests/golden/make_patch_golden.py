"""Pins the random sub-patch tables of maua_b200's Patch against the REFERENCE's own Patch class on a CPU generator and
writes tests/golden/patch.json.  Runs only where /root/reference exists.  The reference's patch.py is imported unmodified;
its heavy sibling mir.py (librosa, scipy.stats, torch_geometric through the segmenters) is replaced by a module carrying
the two feature-name lists parsed out of mir.py's source, torchcubicspline by a dummy (no spline is evaluated here).
    python tests/golden/make_patch_golden.py
"""
import ast
import importlib
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
assert os.path.isdir(REF)
sys.path.insert(0, REF)


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


tcs = types.ModuleType("torchcubicspline")
tcs.natural_cubic_spline_coeffs = lambda *a, **k: None
tcs.NaturalCubicSpline = object
sys.modules["torchcubicspline"] = tcs
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
mir = types.ModuleType("maua.audiovisual.audioreactive.selfsupervised.mir")
consts = {}
for node in ast.parse(open(base + "/selfsupervised/mir.py").read()).body:
    if isinstance(node, ast.Assign) and node.targets[0].id in ("UNITFEATS", "ALLFEATS"):
        consts[node.targets[0].id] = eval(compile(ast.Expression(node.value), "mir", "eval"), dict(consts))
mir.UNITFEATS, mir.ALLFEATS = consts["UNITFEATS"], consts["ALLFEATS"]
sys.modules[mir.__name__] = mir
ref_patch = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.patch")

from maua_b200.audiovisual.audioreactive import patch as P  # noqa: E402

assert (P.UNITFEATS, P.ALLFEATS) == (mir.UNITFEATS, mir.ALLFEATS)
T = 64
features = {k: torch.zeros(T, 1) for k in mir.ALLFEATS}
segmentations = {(name, k): torch.zeros(T, dtype=torch.long) for name in mir.ALLFEATS for k in (2, 4, 6, 8, 12, 16)}
gold = []
for seed in (42, 7, 123456):
    a = ref_patch.Patch(features, segmentations, tempo=120.0, fps=24, seed=seed, device="cpu")
    b = P.Patch(features, segmentations, tempo=120.0, fps=24, seed=seed, device="cpu")
    for key in ("n_base_latents", "sigma_base_noise", "loops_base_noise", "latent_patches", "noise_patches", "ks", "length"):
        assert getattr(a, key) == getattr(b, key), (seed, key, getattr(a, key), getattr(b, key))
    assert repr(a) == repr(b)
    a.update_intensity(0.7); b.update_intensity(0.7)
    assert a.latent_patches == b.latent_patches and a.noise_patches == b.noise_patches
    gold.append(dict(seed=seed, n_base_latents=b.n_base_latents, sigma_base_noise=b.sigma_base_noise,
                     loops_base_noise=b.loops_base_noise, latent_patches=b.latent_patches, noise_patches=b.noise_patches, repr=repr(b)))
# skewnorm / random_choice draw for draw
g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
assert torch.equal(ref_patch.skewnorm(g1, 5, 0.6, 0.5, (16,)), P.skewnorm(g2, 5, 0.6, 0.5, (16,)))
json.dump(gold, open(os.path.join(ROOT, "tests", "golden", "patch.json"), "w"))
print("maua_b200 Patch == reference Patch (CPU generator) for seeds 42, 7, 123456; wrote tests/golden/patch.json")
