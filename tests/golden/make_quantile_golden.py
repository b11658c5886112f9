"""Golden vectors for the quantile path, produced by the REFERENCE's own code: its compiled efficient_quantile.cpp
(oracle/_ref/efficient_quantile.so, built by oracle/build_ref.py) behind its own ``quantile`` wrapper, and its own
``standardize`` / ``onset_envelope`` / ``spectral_flux`` (features/processing.py) running on top of it.

run in the build container:  python oracle/build_ref.py && python tests/golden/make_quantile_golden.py
"""
import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
from oracle.build_ref import load_efficient_quantile  # noqa: E402
from oracle import quantile as OQ  # noqa: E402

_eq = load_efficient_quantile()
assert _eq is not None, "run oracle/build_ref.py first"


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


base = REF + "/maua/audiovisual/audioreactive"
_pkg("maua", REF + "/maua"); _pkg("maua.audiovisual", REF + "/maua/audiovisual")
_pkg("maua.audiovisual.audioreactive", base)
_pkg("maua.audiovisual.audioreactive.selfsupervised", base + "/selfsupervised")
_pkg("maua.audiovisual.audioreactive.selfsupervised.features", base + "/selfsupervised/features")
# the package's __init__ (efficient_quantile/__init__.py:1-7) restated around the compiled routine: its own file imports the
# extension by a relative name that only exists after the reference's setup.py ran
eq = types.ModuleType("maua.audiovisual.audioreactive.selfsupervised.features.efficient_quantile")
eq.quantile = lambda tensor, q: _eq(tensor.cpu().flatten(), torch.FloatTensor([q]), True, 3).squeeze().to(tensor.device)
sys.modules[eq.__name__] = eq
ref_proc = importlib.import_module("maua.audiovisual.audioreactive.selfsupervised.features.processing")

g = torch.Generator().manual_seed(11)
cases = {}
for name, n in [("n1", 1), ("n2", 2), ("n7", 7), ("n720", 720), ("n1001", 1001), ("n10800", 10800), ("n20001", 20001)]:
    x = torch.randn(n, generator=g)
    if n >= 720:
        x[::97] = x[3]          # ties
    qs = [0.0, 0.0015, 0.025, 0.25, 0.5, 0.75, 0.975, 0.9985, 1.0]
    want = torch.stack([eq.quantile(x, q) for q in qs])
    mine = torch.stack([OQ.quantile(x, q) for q in qs])
    assert torch.equal(want, mine), (name, want, mine)
    cases[name] = dict(x=x, qs=qs, want=want)
xn = torch.randn(500, generator=g)
xn[::7] = float("nan")
cases["nan"] = dict(x=xn, qs=[0.25, 0.75], want=torch.stack([eq.quantile(xn, q) for q in (0.25, 0.75)]))
assert torch.equal(cases["nan"]["want"], torch.stack([OQ.quantile(xn, q) for q in (0.25, 0.75)]))

flow = torch.randn(720, generator=g) * 3 + 1
spec = torch.rand(720, 24, generator=g)
flux = ref_proc.spectral_flux(spec)
out = dict(cases=cases, flow=flow, standardize=ref_proc.standardize(flow.clone()), spec=spec,
           onset_envelope=ref_proc.onset_envelope(flux.clone()))
assert torch.equal(out["standardize"], OQ.standardize(flow)), "standardize"
assert torch.equal(flux, OQ.spectral_flux(spec)), "spectral_flux"
assert torch.equal(out["onset_envelope"], OQ.onset_envelope(flux)), "onset_envelope"
torch.save(out, os.path.join(HERE, "quantile.pt"))
print("oracle.quantile == reference (compiled efficient_quantile + processing.py), bit for bit; wrote tests/golden/quantile.pt")
