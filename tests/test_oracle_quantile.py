"""The quantile restatement (oracle/quantile.py) against vectors produced by the reference's own compiled
efficient_quantile.cpp + processing.py (tests/golden/make_quantile_golden.py), and against the compiled routine itself when
oracle/_ref holds it (built by oracle/build_ref.py in the build container; it travels to the GPU box)."""
import os

import pytest
import torch

from oracle import quantile as OQ
from oracle.build_ref import load_efficient_quantile

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "quantile.pt"))


def test_quantile_matches_reference_vectors_bit_for_bit():
    for name, c in G["cases"].items():
        got = torch.stack([OQ.quantile(c["x"], q) for q in c["qs"]])
        assert torch.equal(got, c["want"]), name


def test_midpoint_and_float32_q_quirks():
    x = torch.arange(10, dtype=torch.float32)
    assert float(OQ.quantile(x, 0.5)) == 4.5                     # mid point of the two neighbours, never a linear blend
    assert float(OQ.quantile(x, 0.26)) == 2.5 and float(torch.quantile(x, 0.26)) != 2.5
    # q is rounded to float32 first: 0.7 -> 0.699999988..., so 0.7 * 10 lands below 7 and the mid point of (6, 7) comes
    # out where a double q would hit the order statistic 7 exactly
    x11 = torch.arange(11, dtype=torch.float32)
    assert float(OQ.quantile(x11, 0.7)) == 6.5
    assert float(OQ.quantile(x, 0.0)) == 0.0 and float(OQ.quantile(x, 1.0)) == 9.0
    assert torch.isnan(OQ.quantile(torch.tensor([float("nan")]), 0.5))


def test_standardize_and_onset_envelope_match_reference_vectors():
    assert torch.equal(OQ.standardize(G["flow"]), G["standardize"])
    assert torch.equal(OQ.onset_envelope(OQ.spectral_flux(G["spec"])), G["onset_envelope"])


def test_against_the_compiled_reference_routine():
    ref = load_efficient_quantile()
    if ref is None:
        pytest.skip("oracle/_ref/efficient_quantile.so not built (python oracle/build_ref.py)")
    g = torch.Generator().manual_seed(5)
    for n in (1, 3, 10, 999, 4096):
        x = torch.randn(n, generator=g)
        for q in (0.0, 0.1, 0.25, 0.5, 0.7, 0.975, 1.0):
            want = ref(x, torch.FloatTensor([q]), True, 3).squeeze()
            assert torch.equal(OQ.quantile(x, q), want), (n, q)
