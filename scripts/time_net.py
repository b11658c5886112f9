"""Time the full-size StyleGAN3 synthesis forward (random init) on cuda:0: usage time_net.py B warmup iters [T|R]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maua_b200.GAN.networks import stylegan3 as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = sys.argv[4] if len(sys.argv) > 4 else "T"
res = int(sys.argv[5]) if len(sys.argv) > 5 else 1024
torch.manual_seed(0)
net = N.SynthesisNetwork(w_dim=512, img_resolution=res, img_channels=3, **(N.SG3_R_KWARGS if cfg == "R" else {}))
for k, v in os.environ.items():
    if k.startswith("MBOPT_"):
        net.set_option(k[6:].lower(), int(v))
ws = torch.randn(B, net.num_ws, 512, device="cuda")
out = torch.empty(B, res, res, 3, device="cuda", dtype=torch.uint8)
for _ in range(warm):
    net(ws, out_fmt="u8", out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    net(ws, out_fmt="u8", out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"cfg {cfg} res {res} B {B}: {ms:.3f} ms/batch  {ms / B:.3f} ms/frame  {1000 * B / ms:.1f} frames/s  launches {net.last_launch_count()}  out mean {out.float().mean().item():.2f}")
