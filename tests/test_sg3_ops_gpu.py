"""Parity of the two hot ops (through the C ABI) against the CPU oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import sg3 as O

pytestmark = pytest.mark.gpu


def _rel_err(got, ref):
    ref = ref.double()
    return float((got.double().cpu() - ref).abs().max() / ref.square().mean().sqrt())


CONV_CASES = [
    # B, Cin, Cout, H, W, k, demod
    (2, 64, 128, 20, 20, 3, True),
    (2, 81, 51, 70, 66, 3, True),
    (1, 512, 512, 38, 38, 3, True),
    (2, 128, 96, 40, 24, 1, True),
    (1, 203, 130, 150, 150, 3, True),
    (1, 32, 3, 64, 64, 1, False),
    (2, 51, 32, 100, 90, 3, True),
    (1, 32, 32, 70, 70, 3, True),
    (1, 96, 64, 33, 41, 3, True),
    (2, 128, 81, 60, 50, 3, True),
    (1, 40, 17, 45, 37, 3, True),
    (2, 64, 96, 20, 24, 2, True),      # 2x2 kernels: the parity kernels of the polyphase transposed conv (csrc/sg2.cu)
    (1, 48, 256, 33, 33, 2, True),
]


@pytest.mark.parametrize("case", CONV_CASES)
# 1 CUDA cores, 0 tcgen05 product dispatch (narrow layers: resident weight tiles), 2 the same with 16-wide tiles,
# 3 full weight tiles streamed per stage, 4 pixel-major tile up to 128 couts, 7 cout-major tile with one patch load per chunk,
# 8 cout-major tile with three epilogue warp groups (column-split, double-buffered tcgen05.ld), 9 the same with one group
# 10 cout-major tile, one 34-pixel-wide patch load per chunk (7-row tiles, N = 240), three epilogue groups
# 11 pixel-major tile with the three kw taps stacked along N (one instruction per (kh, 16 channels), shuffle-add epilogue)
# 12 cout-major tile with the kw taps stacked along M for layers with <= 32 couts (TMEM column shifts, cross-quadrant exchange)
@pytest.mark.parametrize("impl", [1, 0, 2, 3, 4, 7, 8, 9, 10, 11, 12])
def test_modulated_conv2d(cuda, case, impl):
    from maua_b200 import ops

    B, Cin, Cout, H, W, k, demod = case
    g = torch.Generator().manual_seed(1234 + Cin + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g)
    s = torch.randn(B, Cin, generator=g) + 1.0
    gain = 0.7
    ref = O.modulated_conv2d_ref(x, w, s, demodulate=demod, padding=k - 1, input_gain=torch.tensor(gain))
    got = ops.modulated_conv2d(x.to(cuda), w.to(cuda), s.to(cuda), demodulate=demod, input_gain=gain, impl=impl)
    assert got.shape == ref.shape
    # fp16 operands (x*s, W) and fp16 output, fp32 accumulation: a few 1e-3 of the output RMS
    assert _rel_err(got, ref) < 8e-3


FL_CASES = [
    # C, H, W, up, down, ut, dt, pad(lo,hi), radial
    (3, 38, 38, 2, 2, 12, 12, (9, 8), False),
    (2, 54, 54, 4, 2, 24, 12, (-6, -9), False),
    (2, 150, 130, 2, 2, 12, 12, (-11, -12), False),
    (2, 86, 86, 4, 2, 24, 12, (-6, -9), True),
    (4, 33, 47, 1, 1, 1, 1, (0, 0), False),
    (2, 40, 40, 2, 1, 12, 1, (5, 6), False),
    (2, 40, 40, 1, 2, 1, 12, (5, 6), False),
    (1, 300, 280, 2, 2, 12, 12, (9, 8), False),
    (1, 200, 180, 4, 2, 24, 12, (-7, -8), False),
    (1, 90, 70, 2, 2, 12, 12, (-10, -13), False),
]


@pytest.mark.parametrize("case", FL_CASES)
@pytest.mark.parametrize("impl", ["tensorcore", "generic", "cudacore"])
def test_filtered_lrelu(cuda, case, impl, monkeypatch):
    from maua_b200 import ops

    C, H, W, up, down, ut, dt, (lo, hi), radial = case
    monkeypatch.setenv("MB_FLRELU_IMPL", {"tensorcore": "0", "generic": "1", "cudacore": "2"}[impl])
    g = torch.Generator().manual_seed(99 + H)
    x = (torch.randn(2, C, H, W, generator=g) * 2).half().float()
    b = torch.randn(C, generator=g)
    fu = O.design_lowpass_filter(ut, 8.0, 9.0, 64.0) if ut > 1 else None
    fd = O.design_lowpass_filter(dt, 8.0, 9.0, 64.0, radial=radial) if dt > 1 else None
    pad = [lo, hi, lo, hi]
    ref = O.filtered_lrelu_ref(x, fu=fu, fd=fd, b=b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256)
    got = ops.filtered_lrelu(x.to(cuda), None if fu is None else fu.to(cuda), None if fd is None else fd.to(cuda),
                             b.to(cuda), up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256)
    assert got.shape == ref.shape
    # max error relative to the RMS (values reach ~4-5 RMS).  CUDA-core kernels: fp32 arithmetic, fp16 output
    # rounding only.  Tensor-core chain: fp16 taps (DC-corrected) and fp16 rounding between the four passes.
    assert _rel_err(got, ref) < (1e-2 if impl == "tensorcore" else 5e-3)


# Streaming kernels (flrelu_stream.cuh) against the tile kernels (flrelu_mma.cu) they replace: the MMA sequence per output
# element is the same, so the results must be bit-identical; and against the oracle, in both output layouts.
# C covers full 16-channel groups, partial last groups (v = 1, 3, 11 channels) and fewer than 16 channels;
# heights cover columns of one block up to columns the planner splits into several segments.
STREAM_CASES = [
    # B, C, H, W, up, ut, pad(lo,hi), radial
    (2, 16, 38, 38, 2, 12, (9, 8), False),
    (1, 17, 54, 54, 4, 24, (-6, -9), False),
    (2, 19, 150, 130, 2, 12, (-11, -12), False),
    (1, 35, 86, 86, 4, 24, (-6, -9), True),
    (1, 3, 300, 90, 2, 12, (9, 8), False),
    (1, 27, 200, 180, 4, 24, (-7, -8), False),
    (3, 32, 90, 70, 2, 12, (-10, -13), False),
    (1, 51, 276, 70, 2, 12, (-10, -13), False),
    (1, 20, 120, 100, 2, 12, (-10, -13), True),
    (1, 81, 140, 64, 4, 24, (-6, -9), False),
]


@pytest.mark.parametrize("case", STREAM_CASES)
@pytest.mark.parametrize("layout", ["planar", "nhwc"])
def test_filtered_lrelu_stream(cuda, case, layout, monkeypatch):
    from maua_b200 import ops

    B, C, H, W, up, ut, (lo, hi), radial = case
    monkeypatch.setenv("MB_FLRELU_IMPL", "0")
    monkeypatch.setenv("MB_FLRELU_TEST_NHWC", "1" if layout == "nhwc" else "0")
    g = torch.Generator().manual_seed(7 + H + C)
    x = (torch.randn(B, C, H, W, generator=g) * 2).half().float()
    b = torch.randn(C, generator=g)
    fu = O.design_lowpass_filter(ut, 8.0, 9.0, 64.0)
    fd = O.design_lowpass_filter(12, 8.0, 9.0, 64.0, radial=radial)
    pad = [lo, hi, lo, hi]
    kw = dict(up=up, down=2, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256)
    args = (x.to(cuda), fu.to(cuda), fd.to(cuda), b.to(cuda))
    monkeypatch.setenv("MB_FLRELU_STREAM", "1")
    got = ops.filtered_lrelu(*args, **kw)
    monkeypatch.setenv("MB_FLRELU_STREAM", "0")
    tile = ops.filtered_lrelu(*args, **kw)
    ref = O.filtered_lrelu_ref(x, fu=fu, fd=fd, b=b, **kw)
    assert got.shape == ref.shape
    assert torch.isfinite(got).all()
    assert _rel_err(got, ref) < 1e-2
    assert torch.equal(got, tile), f"stream vs tile kernels differ: max {float((got - tile).abs().max()):.3e}"


def test_filtered_lrelu_clamp_guard(cuda, monkeypatch):
    """The streaming kernel runs a clamp-free, single-instruction activation (t + k |t|, its factor in the output scale) when the
    producer's recorded max |x| proves the clamp inactive.  On inputs far below the clamp the guarded and the clamping variants
    must agree to fp16 rounding and both match the oracle; on inputs that do reach the clamp, the clamping variant (what an
    unknown or large maximum selects) must match the oracle."""
    from maua_b200 import ops

    g = torch.Generator().manual_seed(21)
    fu = O.design_lowpass_filter(12, 8.0, 9.0, 64.0)
    fd = O.design_lowpass_filter(12, 8.0, 9.0, 64.0)
    kw = dict(up=2, down=2, padding=[9, 8, 9, 8], gain=np.sqrt(2), slope=0.2, clamp=256)
    monkeypatch.setenv("MB_FLRELU_IMPL", "0")
    monkeypatch.setenv("MB_FLRELU_STREAM", "1")
    for layout in ("0", "1"):
        monkeypatch.setenv("MB_FLRELU_TEST_NHWC", layout)
        x = (torch.randn(2, 32, 70, 50, generator=g) * 2).half().float()
        b = torch.randn(32, generator=g)
        args = (x.to(cuda), fu.to(cuda), fd.to(cuda), b.to(cuda))
        monkeypatch.setenv("MB_FLRELU_ASSUME_SAFE", "1")
        fast = ops.filtered_lrelu(*args, **kw)
        monkeypatch.setenv("MB_FLRELU_ASSUME_SAFE", "0")
        safe = ops.filtered_lrelu(*args, **kw)
        ref = O.filtered_lrelu_ref(x, fu=fu, fd=fd, b=b, **kw)
        assert not torch.equal(fast, safe)                     # the fast path really is a different instruction sequence
        assert _rel_err(fast, safe.cpu()) < 6e-3 and _rel_err(fast, ref) < 1e-2 and _rel_err(safe, ref) < 1e-2
        print("clamp guard, layout", layout, "fast vs clamping", _rel_err(fast, safe.cpu()), "fast vs oracle", _rel_err(fast, ref), "clamping vs oracle", _rel_err(safe, ref))
    big = (torch.randn(1, 16, 40, 40, generator=g) * 150).half().float()      # reaches +-256 / sqrt(2) after up-sampling
    ref = O.filtered_lrelu_ref(big, fu=fu, fd=fd, b=None, **kw)
    got = ops.filtered_lrelu(big.to(cuda), fu.to(cuda), fd.to(cuda), None, **kw)
    assert float(ref.abs().max()) > 100 and _rel_err(got, ref) < 1e-2
