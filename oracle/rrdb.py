"""Oracle: RRDBNet (the RealESRGAN x4 generator) in plain fp32 PyTorch on CPU.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference instantiates basicsr.archs.rrdbnet_arch.RRDBNet and realesrgan.RealESRGANer
(maua/super/image/models/realesrgan.py:22-49) from an absent submodule (maua/submodules/RealESRGAN) and absent packages;
no source, test or golden vector for them exists under /root/reference.  This file restates the published architecture
(xinntao/BasicSR rrdbnet_arch.py, xinntao/Real-ESRGAN utils.py) and is anchored on the reference's constructor call
(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23 | 6, num_grow_ch=32, scale=4) and on the state-dict key names of the
checkpoints it downloads.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class ResidualDenseBlock(nn.Module):
    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_grow_ch, 3, 1, 1)
        self.conv2 = nn.Conv2d(num_feat + num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv3 = nn.Conv2d(num_feat + 2 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv4 = nn.Conv2d(num_feat + 3 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)

    def forward(self, x):
        lrelu = lambda t: F.leaky_relu(t, 0.2)  # noqa: E731
        x1 = lrelu(self.conv1(x))
        x2 = lrelu(self.conv2(torch.cat((x, x1), 1)))
        x3 = lrelu(self.conv3(torch.cat((x, x1, x2), 1)))
        x4 = lrelu(self.conv4(torch.cat((x, x1, x2, x3), 1)))
        x5 = self.conv5(torch.cat((x, x1, x2, x3, x4), 1))
        return x5 * 0.2 + x


class RRDB(nn.Module):
    def __init__(self, num_feat, num_grow_ch=32):
        super().__init__()
        self.rdb1, self.rdb2, self.rdb3 = (ResidualDenseBlock(num_feat, num_grow_ch) for _ in range(3))

    def forward(self, x):
        return self.rdb3(self.rdb2(self.rdb1(x))) * 0.2 + x


class RRDBNet(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        assert scale == 4
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)

    def forward(self, x):
        feat = self.conv_first(x)
        feat = feat + self.conv_body(self.body(feat))
        feat = F.leaky_relu(self.conv_up1(F.interpolate(feat, scale_factor=2, mode="nearest")), 0.2)
        feat = F.leaky_relu(self.conv_up2(F.interpolate(feat, scale_factor=2, mode="nearest")), 0.2)
        return self.conv_last(F.leaky_relu(self.conv_hr(feat), 0.2))


def make(num_block=23, seed=0, rdb_scale=0.1):
    """Random-init network: torch's Conv2d defaults, the dense blocks' convs scaled like basicsr's default_init_weights(0.1)."""
    torch.manual_seed(seed)
    net = RRDBNet(num_block=num_block)
    for m in net.body.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)
            m.weight.data.mul_(rdb_scale)
            m.bias.data.normal_(0, 0.02)     # non-zero biases so the bias path is exercised
    return net.eval().requires_grad_(False)


def enhance(model, img, pre_pad=10, scale=4):
    """RealESRGANer.enhance(img) with tile=0 (Real-ESRGAN utils.py): HWC 0..255 BGR image -> HWC uint8 at 4x."""
    img = np.asarray(img).astype(np.float32) / 255.0
    x = torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1)))[None]
    x = F.pad(x, (0, pre_pad, 0, pre_pad), "reflect")
    y = model(x)
    _, _, h, w = y.shape
    y = y[:, :, 0: h - pre_pad * scale, 0: w - pre_pad * scale]
    out = y[0].float().clamp_(0, 1).numpy()
    out = np.transpose(out[[2, 1, 0], :, :], (1, 2, 0))
    return (out * 255.0).round().astype(np.uint8)
