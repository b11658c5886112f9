"""Probe: pixel-major conv with ONE patch load and shifted A descriptors (impl 5: plain start offset, impl 6: with the
descriptor base-offset field) against the oracle; wrapped in a timeout by the caller."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maua_b200 import ops
from oracle import sg3 as O
cases = [(2, 51, 32, 100, 90, 3), (1, 32, 32, 70, 70, 3), (1, 96, 64, 33, 41, 3), (2, 81, 51, 60, 50, 3), (1, 40, 17, 45, 37, 3)]
for impl in (4, 5, 6):
    for (B, Cin, Cout, H, W, k) in cases:
        g = torch.Generator().manual_seed(1234 + Cin + H)
        x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g); s = torch.randn(B, Cin, generator=g) + 1.0
        ref = O.modulated_conv2d_ref(x, w, s, demodulate=True, padding=k - 1, input_gain=torch.tensor(0.7))
        got = ops.modulated_conv2d(x.cuda(), w.cuda(), s.cuda(), demodulate=True, input_gain=0.7, impl=impl).cpu()
        err = float((got - ref).norm() / ref.norm())
        print(f"impl {impl} case {(B, Cin, Cout, H, W, k)} rel err {err:.3e}", flush=True)
