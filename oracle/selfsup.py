"""Oracle: the torch-native ("selfsupervised") envelope / latent patch functions of the reference, fp32 on CPU.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates maua/audiovisual/audioreactive/selfsupervised/features/processing.py:11-56,102-139, mir.py:13-21 and
latent.py:7-80.  PINNED by tests/golden/make_selfsup_golden.py, which imports the reference's own modules (absent
third-party imports stubbed as SURVEY Appendix C.2) and requires torch.equal / allclose against these functions before
writing tests/golden/selfsup.pt.  Exception: the natural cubic spline of spline_loop_latents belongs to the absent,
un-pinned torchcubicspline (setup.py:104): restated (natural_spline_eval below, checked against scipy) -- PARITY
UNPINNED for that stage; the golden script pins everything AROUND it by handing the reference the same spline.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.nn.functional import conv1d, pad


def gaussian_filter(x, sigma, mode="circular", causal=1):
    """features/processing.py:11-50 (radius <= n_frames branch)."""
    dim = len(x.shape)
    while len(x.shape) < 3:
        x = x[:, None]
    radius = min(int(sigma * 4), 3 * len(x))
    channels = x.shape[1]
    kernel = torch.arange(-radius, radius + 1, dtype=torch.float32)
    kernel = torch.exp(-0.5 / sigma ** 2 * kernel ** 2)
    kernel = kernel / kernel.sum()
    kernel = kernel.view(1, 1, len(kernel)).repeat(channels, 1, 1)
    if dim == 4:
        t, c, h, w = x.shape
        x = x.view(t, c, h * w)
    x = x.transpose(0, 2)
    x = pad(x, (radius, radius), mode=mode)
    x = conv1d(x, weight=kernel, groups=channels)
    x = x.transpose(0, 2)
    if dim == 4:
        x = x.view(t, c, h, w)
    if len(x.shape) > dim:
        x = x.squeeze()
    return x


def normalize(array):
    array = array - array.min()
    return array / (array.max() + 1e-8)


def salience_weighted(envelope, short_sigma=5, long_sigma=80):
    """mir.py:13-21."""
    if envelope.dim() > 1:
        envelope = envelope.squeeze(1)
    short = gaussian_filter(envelope, short_sigma, mode="reflect", causal=0)
    long = gaussian_filter(envelope, long_sigma, mode="reflect", causal=0)
    weighted = (short / long) ** 2 * envelope
    if weighted.dim() < 2:
        weighted = weighted.unsqueeze(1)
    return weighted


def clamp_peaks_percentile(signal, percent):
    """features/processing.py:102-122."""
    if len(signal.shape) < 2:
        signal = signal.unsqueeze(1)
    result = []
    for sig in signal.unbind(1):
        locs = torch.arange(0, sig.shape[0])
        peaks = torch.gt(sig, sig[(locs + 1).clamp(0, sig.shape[0] - 1)]) & torch.gt(sig, sig[(locs - 1).clamp(0, sig.shape[0] - 1)])
        result.append(torch.clamp(sig, None, torch.quantile(sig[peaks], percent / 100)))
    return torch.stack(result, dim=1)


def emphasize(envs, strength, percentile):
    """features/processing.py:133-139."""
    lo = envs.min(dim=0).values
    x = envs - lo
    hi = x.max(dim=0).values
    x = x / hi
    x = x * (1 + torch.tanh(strength * (x - torch.quantile(x, q=percentile / 100, dim=0))))
    return (x * hi) + lo


def drop_strength_from_rms(rms):
    """features/audio.py:38-39 given rms(audio, sr) [T, 1]."""
    return emphasize(gaussian_filter(rms, 10), strength=10, percentile=50).unsqueeze(1)


def tonnetz_from_chroma(chroma):
    """features/audio.py:46-56 given chroma_fn(y, sr) [12, T] -> [T, 6]."""
    dim_map = torch.linspace(0, 12, chroma.shape[0])
    scale = torch.tensor([7.0 / 6, 7.0 / 6, 3.0 / 2, 3.0 / 2, 2.0 / 3, 2.0 / 3])
    V = scale.reshape(-1, 1) * dim_map
    V[::2] -= 0.5
    R = torch.tensor([1, 1, 1, 1, 0.5, 0.5])
    phi = R[:, None] * torch.cos(torch.pi * V)
    return (phi @ (chroma / chroma.norm(p=1, dim=0))).T


def natural_spline_eval(t_in, y, t_out):
    """Natural cubic spline through (t_in[m], y[m, ...]) with uniform knots, evaluated at t_out (float64 inside)."""
    m = len(y)
    yy = y.reshape(m, -1).double().numpy()
    h = float(t_in[1] - t_in[0])
    z = np.zeros_like(yy)
    if m > 2:
        a = np.zeros((m - 2, m - 2))
        np.fill_diagonal(a, 4.0)
        idx = np.arange(m - 3)
        a[idx, idx + 1] = 1.0
        a[idx + 1, idx] = 1.0
        z[1:-1] = np.linalg.solve(a, 6.0 / h ** 2 * (yy[:-2] - 2 * yy[1:-1] + yy[2:]))
    u = (t_out.double().numpy() - float(t_in[0])) / h
    i = np.clip(np.floor(u).astype(np.int64), 0, m - 2)
    f = (u - i)[:, None]
    g = 1.0 - f
    out = g * yy[i] + f * yy[i + 1] + h * h / 6.0 * ((g ** 3 - g) * z[i] + (f ** 3 - f) * z[i + 1])
    return torch.from_numpy(out).to(y.dtype).reshape(len(t_out), *y.shape[1:])


def spline_loop_latents(y, size, n_loops=1):
    """latent.py:7-13."""
    y = torch.cat((y, y[[0]]))
    t_in = torch.linspace(0, 1, len(y)).to(y)
    t_out = torch.linspace(0, n_loops, size).to(y) % 1
    return natural_spline_eval(t_in, y, t_out)


def latent_patch(rng, latents, palette, segmentations, features, tempo, fps, patch_type, segments, loop_bars, seq_feat,
                 seq_feat_weight, mod_feat, mod_feat_weight, merge_type, merge_depth):
    """latent.py:16-80."""
    feature = seq_feat_weight * features[seq_feat]
    segmentation = segmentations[(seq_feat, segments)]
    permutation = torch.randperm(len(palette), generator=rng, device=rng.device)
    if patch_type == "segmentation":
        selection = permutation[:segments]
        selectseq = selection[segmentation.cpu().numpy()]
        sequence = gaussian_filter(palette[selectseq], 5)
    elif patch_type == "feature":
        n_select = feature.shape[1]
        if n_select == 1:
            selection = permutation[:2]
            sequence = feature[..., None] * palette[selection][[0]] + (1 - feature[..., None]) * palette[selection][[1]]
        else:
            selection = permutation[:n_select]
            sequence = torch.einsum("TN,NWL->TWL", feature, palette[selection])
    elif patch_type == "loop":
        selection = permutation[:segments]
        n_loops = len(latents) / fps / 60 / tempo / 4 / loop_bars
        sequence = spline_loop_latents(palette[selection], len(latents), n_loops=n_loops)
    sequence = gaussian_filter(sequence, 1)
    lays = {"low": slice(0, 6), "mid": slice(6, 12), "high": slice(12, 18), "lowmid": slice(0, 12), "midhigh": slice(6, 18),
            "all": slice(0, 18)}[merge_depth]
    if merge_type == "average":
        latents[:, lays] += sequence[:, lays]
        latents[:, lays] /= 2
    elif merge_type == "modulate":
        modulation = mod_feat_weight * features[mod_feat][..., None]
        latents[:, lays] *= 1 - modulation
        latents[:, lays] += modulation * sequence[:, lays]
    else:
        latents[:, lays] = sequence[:, lays]
    return latents
