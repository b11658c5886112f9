"""Oracle: the reference's in-tree StyleGAN2 inference network, fp32 PyTorch on CPU.  TEST INFRASTRUCTURE ONLY.

Restates maua/GAN/wrappers/inference/ops.py:65-256 and inference/stylegan2.py:29-436 (the only synthesis
arithmetic that ships inside the reference tree).  As written the reference forward raises (SURVEY F4:
``padding.repeat(4)`` on an int at ops.py:200, a 5-D permute of a 4-D tensor at :213, ``torch.max(...,0)``
namedtuples as conv padding at :222-224); this restatement is what that code computes once those three
mechanical defects are repaired.
PINNED: tests/golden/make_sg2_golden.py imports the reference modules, repairs ``conv2d_resample`` in place and
checks this oracle against the reference network on the same state dict before writing tests/golden/sg2.pt.
Parameter names = the reference's state-dict keys (bs.{i}.conv0.affine.weight, ...).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def setup_filter(f=(1, 3, 3, 1)):
    """ops.py:236-256: 1-D taps -> normalised outer-product 2-D filter (non-separable below 8 taps)."""
    f = torch.as_tensor(f, dtype=torch.float32)
    f = torch.outer(f, f)
    return f / f.sum()


def upfirdn2d(x, f, up=1, down=1, padding=(0, 0, 0, 0), gain=1.0):
    """ops.py:87-114 (2-D filter branch; the filter is NOT flipped)."""
    B, C, H, W = x.shape
    px0, px1, py0, py1 = [int(p) for p in padding]
    x = x.reshape(B, C, H, 1, W, 1)
    x = F.pad(x, [0, up - 1, 0, 0, 0, up - 1]).reshape(B, C, H * up, W * up)
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    f = f * (gain ** (f.ndim / 2))
    x = F.conv2d(x, f[None, None].repeat(C, 1, 1, 1), groups=C)
    return x[:, :, ::down, ::down]


def upsample2d(x, f, up=2):
    """ops.py:117-133."""
    fw = fh = f.shape[-1]
    p = ((fw + up - 1) // 2, (fw - up) // 2, (fh + up - 1) // 2, (fh - up) // 2)
    return upfirdn2d(x, f, up=up, padding=p, gain=up * up)


def bias_act(x, b=None, act="linear", gain=None, clamp=None):
    """ops.py:65-84 (linear / lrelu 0.2)."""
    if b is not None:
        x = x + b.reshape([-1 if i == 1 else 1 for i in range(x.ndim)])
    if act == "lrelu":
        x = F.leaky_relu(x, 0.2)
    g = (math.sqrt(2) if act == "lrelu" else 1.0) if gain is None else gain
    if g != 1:
        x = x * g
    if clamp is not None and clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def conv2d_resample(x, w, f, up, padding, groups):
    """ops.py:189-233 with the three repairs; only the up in {1, 2}, down == 1 branches the network uses."""
    O, I, kh, kw = w.shape
    fw = f.shape[-1]
    px0 = px1 = py0 = py1 = int(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2
        py0 += (fw + up - 1) // 2; py1 += (fw - up) // 2
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, O // groups, I, kh, kw).transpose(1, 2).reshape(groups * I, O // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = F.conv_transpose2d(x, w, stride=up, padding=(pyt, pxt), groups=groups)
        return upfirdn2d(x, f, padding=(px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt), gain=up ** 2)
    return F.conv2d(x, w, padding=(py0, px0), groups=groups)


def modulated_conv2d(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None, demodulate=True):
    """ops.py:146-186 (fp32 path: no pre-normalisation)."""
    B, C, H, W = x.shape
    O, I, kh, kw = weight.shape
    w = weight.unsqueeze(0) * styles.reshape(B, 1, I, 1, 1)
    if demodulate:
        w = w / ((w * w).sum((2, 3, 4)) + 1e-8).sqrt().reshape(B, O, 1, 1, 1)
    x = conv2d_resample(x.reshape(1, B * C, H, W), w.reshape(B * O, I, kh, kw), resample_filter, up, padding, B)
    x = x.reshape(B, O, H * up, W * up)
    return x if noise is None else x + noise


class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_features, out_features, bias_init=0.0):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]))
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init)))
        self.weight_gain = 1 / math.sqrt(in_features)

    def forward(self, x):
        return F.linear(x, self.weight * self.weight_gain, self.bias)


class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, up=1, conv_clamp=256.0):
        super().__init__()
        self.in_channels, self.out_channels, self.resolution, self.up, self.conv_clamp = in_channels, out_channels, resolution, up, conv_clamp
        self.register_buffer("resample_filter", setup_filter())
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, 3, 3]))
        self.register_buffer("noise_const", torch.randn([resolution, resolution]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w):
        x = modulated_conv2d(x, self.weight, self.affine(w), noise=self.noise_const, up=self.up, padding=1,
                             resample_filter=self.resample_filter)
        return bias_act(x, self.bias, act="lrelu", clamp=self.conv_clamp)


class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, conv_clamp=256.0):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, 1, 1]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / math.sqrt(in_channels)

    def forward(self, x, w):
        x = modulated_conv2d(x, self.weight, self.affine(w) * self.weight_gain, demodulate=False,
                             resample_filter=setup_filter())
        return bias_act(x, self.bias, clamp=self.conv_clamp)


class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels):
        super().__init__()
        self.in_channels, self.resolution = in_channels, resolution
        self.register_buffer("resample_filter", setup_filter())
        self.num_conv, self.num_torgb = (1 if in_channels == 0 else 2), 1
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim, resolution, up=2)
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim, resolution)
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim)

    def forward(self, x, img, ws):
        i = 0
        if self.in_channels == 0:
            x = self.const.unsqueeze(0).repeat(ws.shape[0], 1, 1, 1)
        else:
            x = self.conv0(x, ws[:, i]); i += 1
        x = self.conv1(x, ws[:, i]); i += 1
        if img is not None:
            img = upsample2d(img, self.resample_filter)
        y = self.torgb(x, ws[:, i])
        return x, (img + y if img is not None else y)


class SynthesisNetwork(torch.nn.Module):
    """stylegan2.py:385-436 (architecture 'skip', fp32)."""

    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512):
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self.block_resolutions = [2 ** i for i in range(2, int(np.log2(img_resolution)) + 1)]
        ch = {r: min(channel_base // r, channel_max) for r in self.block_resolutions}
        self.num_ws = 0
        bs = []
        for r in self.block_resolutions:
            blk = SynthesisBlock(ch[r // 2] if r > 4 else 0, ch[r], w_dim, r, img_channels)
            self.num_ws += blk.num_conv + (blk.num_torgb if r == img_resolution else 0)
            bs.append(blk)
        self.bs = torch.nn.ModuleList(bs)

    def forward(self, ws):
        x = img = None
        i = 0
        for blk in self.bs:
            x, img = blk(x, img, ws.narrow(1, i, blk.num_conv + blk.num_torgb))
            i += blk.num_conv
        return img


def make_synthesis(img_resolution=256, seed=0, **kw):
    torch.manual_seed(seed)
    return SynthesisNetwork(w_dim=512, img_resolution=img_resolution, img_channels=3, **kw).eval().requires_grad_(False)
