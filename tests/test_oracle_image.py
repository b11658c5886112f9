"""oracle/image.py against the vectors produced by the reference's own resample (tests/golden/make_image_golden.py),
and the product's host-side Lanczos tap design against the oracle's."""
import os

import torch

from oracle import image as OI

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "image.pt"))


def test_resample_matches_reference_vectors():
    for name, size in {"down": (20, 36), "down_h_up_w": (30, 100), "up": (96, 80), "short_side": 24}.items():
        assert torch.equal(OI.resample(G["x"].clone(), size), G[name]), name


def test_lanczos_design_of_the_host_facade_matches_oracle():
    from maua_b200 import ops

    for ratio in (20 / 48, 36 / 64, 0.5, 0.9):
        assert torch.equal(ops._lanczos_taps(ratio), OI.lanczos(OI.ramp(ratio, 2), 2))
