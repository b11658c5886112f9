"""Per-layer device times (CUDA events recorded by the library) of the full-size forward: layer_times.py [B] [T|R]"""
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
for _ in range(3):
    net(ws, out_fmt="u8", out=out)
net.set_option("profile", 2); net.set_option("profile_reset", 1)
iters = 5
for _ in range(iters):
    net(ws, out_fmt="u8", out=out)
torch.cuda.synchronize()
acc = {}
for kind, layer, ms in net.profile_read():
    acc[(kind, layer)] = acc.get((kind, layer), 0.0) + ms / iters
names = {0: "styles", 1: "input", 2: "conv", 3: "flrelu", 4: "transpose", 5: "torgb"}
geo = net.geometry["layers"]
tot = sum(acc.values())
print(f"B={B} cfg={cfg}: total {tot:.3f} ms/batch = {tot / B:.3f} ms/frame = {1000 * B / tot:.1f} frames/s")
for i, g in enumerate(geo):
    c, f = acc.get((2, i), 0.0), acc.get((3, i), 0.0)
    hc = g["in_size"] + g["conv_kernel"] - 1
    gf = 2.0 * g["in_channels"] * g["out_channels"] * g["conv_kernel"] ** 2 * hc * hc * B / 1e9
    mb = g["out_channels"] * (hc * hc + g["out_size"] ** 2) * 2.0 * B / 1e6
    print(f"{g['name']:14s} conv {c:7.3f} ms ({gf / c if c else 0:7.0f} TFLOP/s... GF={gf:7.1f})   flrelu {f:7.3f} ms ({mb / f / 1e3 if f else 0:6.2f} TB/s alg)  up{g['up']}")
print({names[k]: round(sum(v for (kk, _), v in acc.items() if kk == k), 3) for k in names})
