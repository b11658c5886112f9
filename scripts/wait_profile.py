"""Where the warp roles of a conv kernel wait (needs a -DMB_WAIT_PROFILE build and MB_DEBUG=1): wait_profile.py impl Cin Cout"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maua_b200 import ops, _lib
impl, Cin, Cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B, H = 4, 1042
x = torch.randn(B, Cin, H, H, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda")
s = torch.randn(B, Cin, device="cuda")
ops.modulated_conv2d(x, w, s, impl=impl)
torch.cuda.synchronize()
a = _lib.debug_words(64)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.modulated_conv2d(x, w, s, impl=impl)
torch.cuda.synchronize()
b = _lib.debug_words(64)
names = {1: "producer waits empty[s]", 2: "MMA waits tempty[acc]", 3: "MMA waits full[s]", 4: "epilogue waits tfull[acc]", 5: "MMA waits weights"}
tiles = {12: -(-(H + 2) // 7) * -(-(H + 2) // 32) * B, 11: -(-(H + 2) // 8) * -(-(H + 2) // 30) * B}.get(impl, 0)
print(f"impl {impl} Cin {Cin} Cout {Cout}: ~{tiles / 148:.0f} tiles per CTA")
for tag, name in names.items():
    cyc, n = (b[32 + tag] - a[32 + tag]) * 16, b[48 + tag] - a[48 + tag]
    print(f"  {name:28s} {n:6d} waits, {cyc:10d} cycles total, {cyc / max(n, 1):8.0f} per wait")
