"""Generates tests/golden/sg3_tiny.pt from the oracle restatement (oracle/sg3.py).

PARITY UNPINNED: the reference's StyleGAN3 network source is an un-vendored submodule (maua/GAN/nv), so these
vectors pin the ORACLE against regressions and give the CUDA path a fixed target; they are not outputs of the
reference itself.  Run from the repo root:  python tests/golden/make_sg3_golden.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sg3 as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "T64": dict(config="T", img_resolution=64, channel_base=1024, channel_max=32),
    "R64": dict(config="R", img_resolution=64, channel_base=2048, channel_max=48),
}

out = {}
for name, kw in CASES.items():
    kw = dict(kw)
    net = O.make_synthesis(kw.pop("config"), seed=3, **kw)
    torch.manual_seed(11)
    ws = torch.randn(2, net.num_ws, 512)
    img, acts = net(ws, return_activations=True)
    out[name] = dict(ws=ws, img=img.half(), act_rms=torch.tensor([a.square().mean().sqrt() for a in acts]),
                     act_shapes=[tuple(a.shape) for a in acts])
torch.save(out, os.path.join(HERE, "sg3_tiny.pt"))
print({k: (v["img"].shape, v["act_rms"][-1].item()) for k, v in out.items()})
