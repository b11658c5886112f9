"""One warm-up + one profiled full-size forward (for ncu --launch-skip): one_forward.py [B] [T|R]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maua_b200.GAN.networks import stylegan3 as N
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = sys.argv[2] if len(sys.argv) > 2 else "T"
torch.manual_seed(0)
net = N.SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3, **(N.SG3_R_KWARGS if cfg == "R" else {}))
for k, v in os.environ.items():
    if k.startswith("MBOPT_"):
        net.set_option(k[6:].lower(), int(v))
ws = torch.randn(B, net.num_ws, 512, device="cuda")
out = torch.empty(B, 1024, 1024, 3, device="cuda", dtype=torch.uint8)
for _ in range(2):
    net(ws, out_fmt="u8", out=out)
torch.cuda.synchronize()
