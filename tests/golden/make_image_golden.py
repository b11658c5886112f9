"""Pins oracle/image.py against the REFERENCE's own maua/ops/image.py::resample and writes tests/golden/image.pt.

Runs only where /root/reference exists.  maua/ops/image.py imports two third-party modules that are absent here and
unrelated to resample (medpy's noise estimator, resize_right): they are stubbed; everything resample executes is the
reference's own code.      python tests/golden/make_image_golden.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
assert os.path.isdir(REF), "the reference checkout is needed to (re)generate the image golden vectors"
sys.path.insert(0, REF)
for name in ("medpy", "medpy.filter", "medpy.filter.noise", "resize_right"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["medpy.filter.noise"].immerkaer = lambda *a, **k: None
sys.modules["resize_right"].resize = lambda *a, **k: None
# skip the heavy package __init__ files: load maua/ops/image.py as a plain module
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_image", REF + "/maua/ops/image.py")
ref_image = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_image)

from oracle import image as OI  # noqa: E402

torch.manual_seed(0)
x = torch.rand(2, 3, 48, 64)
cases = {"down": (20, 36), "down_h_up_w": (30, 100), "up": (96, 80), "short_side": 24}
out = {"x": x}
for name, size in cases.items():
    r = ref_image.resample(x.clone(), size)
    assert torch.equal(OI.resample(x.clone(), size), r), name
    out[name] = r
noise = torch.randn(3, 1, 32, 32)
out["noise"] = noise
out["pyr_8"] = OI.noise_pyramid_level(noise, (8, 8))
out["pyr_128"] = OI.noise_pyramid_level(noise, (128, 128))
torch.save(out, os.path.join(ROOT, "tests", "golden", "image.pt"))
print("oracle == reference for resample; wrote tests/golden/image.pt", {k: tuple(v.shape) for k, v in out.items()})
