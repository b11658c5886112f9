import os, sys, torch
sys.path.insert(0, os.getcwd())
from maua_b200.GAN.networks import stylegan3 as N
from oracle import sg3 as O
for cfg in ("T", "R"):
    onet = O.make_synthesis(cfg, 1024, seed=0)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3, **(N.SG3_R_KWARGS if cfg == "R" else {}))
    net.load_state_dict(onet.state_dict())
    torch.manual_seed(1)
    ws = torch.randn(1, net.num_ws, 512)
    ref = ((onet(ws) + 1) / 2).clamp(0, 1)
    for guard in ("1", "0"):
        os.environ["MB_FLRELU_GUARD"] = guard
        out = ((net(ws.cuda()).cpu() + 1) / 2).clamp(0, 1)
        print(cfg, "guard", guard, "max-abs pixel error vs oracle", float((out - ref).abs().max()), "rms", float((out - ref).pow(2).mean().sqrt()))
