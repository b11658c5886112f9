"""Pins oracle/sg2.py against the REFERENCE's in-tree StyleGAN2 inference network and writes tests/golden/sg2.pt.

The reference modules (maua/GAN/wrappers/inference/{ops,stylegan2}.py) are imported unmodified; their forward raises
as written, so ``conv2d_resample`` is repaired in place with the three mechanical fixes the survey lists
(python ints for the paddings, a transpose for groups == 1, python max/min for the transposed-conv padding) -- every
other line that runs is the reference's own.  Needs /root/reference.
    python tests/golden/make_sg2_golden.py
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
INF = "/root/reference/maua/GAN/wrappers/inference"
assert os.path.isdir(INF), "the reference checkout is needed to (re)generate the StyleGAN2 golden vectors"


def _load(name, path, package):
    spec = importlib.util.spec_from_file_location(f"{package}.{name}", path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


import types
pkg = types.ModuleType("refinf"); pkg.__path__ = [INF]; sys.modules["refinf"] = pkg
ops = _load("ops", INF + "/ops.py", "refinf")
sg2 = _load("stylegan2", INF + "/stylegan2.py", "refinf")


def conv2d_resample_repaired(x, w, f=None, up=1, down=1, padding=0, groups=1):
    up, down, padding, groups = int(up), int(down), int(padding), int(groups)
    out_channels, in_per_group, kh, kw = w.shape
    fw, fh = ops._get_filter_size(f)
    px0 = px1 = py0 = py1 = padding                                    # repair 1 (ops.py:200)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
        if groups == 1:
            w = w.transpose(0, 1)                                       # repair 2 (ops.py:213)
        else:
            w = w.reshape(groups, out_channels // groups, in_per_group, kh, kw).permute(0, 2, 1, 3, 4)
            w = w.reshape(groups * in_per_group, out_channels // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0); pyt = max(min(-py0, -py1), 0)   # repair 3 (ops.py:222-223)
        x = torch.nn.functional.conv_transpose2d(x, w, stride=up, padding=(pyt, pxt), groups=groups)
        return ops.upfirdn2d(x=x, f=f, padding=torch.tensor([px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt]), gain=torch.tensor(up ** 2))
    return torch.nn.functional.conv2d(x, w, padding=(py0, px0), groups=groups)


ops.conv2d_resample = conv2d_resample_repaired
sg2.conv2d_resample = conv2d_resample_repaired

from oracle import sg2 as O  # noqa: E402

out = {}
with torch.inference_mode():
    for res, kw in [(32, dict(channel_base=512, channel_max=32)), (64, dict(channel_base=2048, channel_max=64))]:
        torch.manual_seed(2)
        ref = sg2.SynthesisNetwork(w_dim=512, img_resolution=res, img_channels=3, **kw).eval()
        mine = O.make_synthesis(res, seed=2, **kw)
        missing = mine.load_state_dict(ref.state_dict(), strict=True)
        torch.manual_seed(4)
        ws = torch.randn(2, ref.num_ws, 512)
        a, b = ref(ws), mine(ws)
        assert a.shape == b.shape == (2, 3, res, res)
        err = float((a - b).abs().max())
        assert err <= 1e-4 * float(a.abs().max()), f"oracle != reference at {res}: {err}"
        out[f"sg2_{res}"] = dict(kw=kw, seed=2, ws=ws, img=a.clone(), state={k: v.clone() for k, v in ref.state_dict().items()} if res == 32 else None)
        print(f"res {res}: oracle vs reference max abs diff {err:.3e} (|img| max {float(a.abs().max()):.2f}), num_ws {ref.num_ws}")
torch.save(out, os.path.join(ROOT, "tests", "golden", "sg2.pt"))
print("wrote tests/golden/sg2.pt")
