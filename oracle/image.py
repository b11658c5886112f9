"""Oracle: image-space helpers of the reference, fp32 on CPU.  TEST INFRASTRUCTURE ONLY.

Restates maua/ops/image.py:190-240 (sinc, lanczos, ramp, resample), maua/ops/noise.py:23-88 (perlin_noise, with the
gradient angles passed in so the host RNG draw is shared with the device path) and the bicubic + unit-std step of
maua/GAN/wrappers/stylegan2.py:196-213 (make_noise_pyramid).  PINNED: tests/golden/make_image_golden.py imports the
reference's own maua/ops/image.py (medpy / resize_right stubbed: absent third-party imports unrelated to resample) and
requires torch.equal for resample before writing tests/golden/image.pt.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def sinc(x):
    return torch.where(x != 0, torch.sin(math.pi * x) / (math.pi * x), x.new_ones([]))


def lanczos(x, a):
    cond = torch.logical_and(-a < x, x < a)
    out = torch.where(cond, sinc(x) * sinc(x / a), x.new_zeros([]))
    return out / out.sum()


def ramp(ratio, width):
    n = math.ceil(width / ratio + 1)
    out = torch.empty([n])
    cur = 0
    for i in range(out.shape[0]):
        out[i] = cur
        cur += ratio
    return torch.cat([-out[1:].flip([0]), out])[1:-1]


def resample(input, size, align_corners=True):
    """image.py:214-240."""
    n, c, h, w = input.shape
    if isinstance(size, (int, float)):
        short, long = (w, h) if w <= h else (h, w)
        new_short, new_long = round(size), round(size * long / short)
        dw, dh = (new_short, new_long) if w <= h else (new_long, new_short)
    else:
        dh, dw = size
    input = input.view([n * c, 1, h, w])
    if dh < h:
        kernel_h = lanczos(ramp(dh / h, 2), 2).to(input.device, input.dtype)
        pad_h = (kernel_h.shape[0] - 1) // 2
        input = F.pad(input, (0, 0, pad_h, pad_h), "reflect")
        input = F.conv2d(input, kernel_h[None, None, :, None])
    if dw < w:
        kernel_w = lanczos(ramp(dw / w, 2), 2).to(input.device, input.dtype)
        pad_w = (kernel_w.shape[0] - 1) // 2
        input = F.pad(input, (pad_w, pad_w, 0, 0), "reflect")
        input = F.conv2d(input, kernel_w[None, None, None, :])
    input = input.view([n, c, h, w])
    return F.interpolate(input, (dh, dw), mode="bicubic", align_corners=align_corners)


def noise_pyramid_level(noise, size):
    """stylegan2.py:203-212: bicubic (align_corners=False) resize, divided by the per-sample std."""
    out = F.interpolate(noise, size, mode="bicubic", align_corners=False)
    return out / out.std((1, 2, 3), keepdim=True)


def perlin_noise(shape, res, theta, phi, tileable=(True, False, False)):
    """ops/noise.py:27-88 with the random angles theta / phi [res+1]^3 given (np.random draws in the reference)."""
    delta = (res[0] / shape[0], res[1] / shape[1], res[2] / shape[2])
    d = (shape[0] // res[0], shape[1] // res[1], shape[2] // res[2])
    grid = np.mgrid[0: res[0]: delta[0], 0: res[1]: delta[1], 0: res[2]: delta[2]].astype(np.float32)
    grid = torch.from_numpy(grid.transpose(1, 2, 3, 0) % 1)
    gradients = np.stack((np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)), axis=3)
    if tileable[0]:
        gradients[-1, :, :] = gradients[0, :, :]
    if tileable[1]:
        gradients[:, -1, :] = gradients[:, 0, :]
    if tileable[2]:
        gradients[:, :, -1] = gradients[:, :, 0]
    gradients = torch.from_numpy(gradients.repeat(d[0], 0).repeat(d[1], 1).repeat(d[2], 2))
    g = {}
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                sl = (slice(d[0], None) if a else slice(None, -d[0]), slice(d[1], None) if b else slice(None, -d[1]),
                      slice(d[2], None) if c else slice(None, -d[2]))
                off = torch.stack((grid[..., 0] - a, grid[..., 1] - b, grid[..., 2] - c), axis=3)
                g[(a, b, c)] = torch.sum(off * gradients[sl], 3)
    t = grid * grid * grid * (grid * (grid * 6 - 15) + 10)
    n00 = g[(0, 0, 0)] * (1 - t[..., 0]) + t[..., 0] * g[(1, 0, 0)]
    n10 = g[(0, 1, 0)] * (1 - t[..., 0]) + t[..., 0] * g[(1, 1, 0)]
    n01 = g[(0, 0, 1)] * (1 - t[..., 0]) + t[..., 0] * g[(1, 0, 1)]
    n11 = g[(0, 1, 1)] * (1 - t[..., 0]) + t[..., 0] * g[(1, 1, 1)]
    n0 = (1 - t[..., 1]) * n00 + t[..., 1] * n10
    n1 = (1 - t[..., 1]) * n01 + t[..., 1] * n11
    return ((1 - t[..., 2]) * n0 + t[..., 2] * n1) * 2 - 1
