"""Constant-Q chroma on the device: host wrapper over mb_chroma_cqt (csrc/chroma.cu).

Mirrors maua/audiovisual/audioreactive/selfsupervised/features/rosa/spectral.py:286-325 ``chroma_cqt`` and
rosa/constantq.py:13-116 ``cqt`` (same names, arguments and output shapes).  The filter-bank, decimation-kernel and
fold-matrix DESIGN below runs once per (sr, geometry) on the host -- constants of the transform, like the mel
filterbank in features.py -- and follows the reference's formulas so the device kernel multiplies the same numbers;
every per-sample operation (decimation, framing, FFT, filter bank, magnitude, fold, normalisation) is CUDA.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ... import _lib

C1_HZ = 32.70319566257483  # note_to_hz("C1"): the reference's fmin default (rosa/convert.py:129-130)


def kaiser_decimation_kernel(lowpass_filter_width=6, rolloff=0.99, beta=14.769656459379492):
    """The 2:1 `sinc_interp_kaiser` kernel torchaudio.functional.resample builds for my_sr -> my_sr/2
    (constantq.py:83; torchaudio _get_sinc_resample_kernel with orig_freq=2, new_freq=1 after the gcd), float32 [28],
    and its left padding `width` (13).  Same expression order and dtypes as torchaudio so the taps are identical."""
    orig_freq, new_freq = 2, 1
    base_freq = min(orig_freq, new_freq) * rolloff
    width = math.ceil(lowpass_filter_width * orig_freq / base_freq)
    idx = torch.arange(-width, width + orig_freq, dtype=torch.float64)[None, None] / orig_freq
    t = torch.arange(0, -new_freq, -1)[:, None, None] / new_freq + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    beta_tensor = torch.tensor(float(beta))
    window = torch.i0(beta_tensor * torch.sqrt(1 - (t / lowpass_filter_width) ** 2)) / torch.i0(beta_tensor)
    t *= math.pi
    scale = base_freq / orig_freq
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.to(torch.float32).reshape(-1).contiguous(), width


def _constant_q_lengths(sr, fmin, n_bins, bins_per_octave, filter_scale=1, gamma=0.0):
    alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
    q = float(filter_scale) / alpha
    freq = fmin * (2.0 ** (torch.arange(n_bins, dtype=torch.float) / bins_per_octave))
    return q * sr / (freq + gamma / alpha)


def octave_filter_bank(sr, fmin_top, bins_per_octave, filter_scale=1, sparsity=0.01):
    """One-sided FFT of the top octave's hann-windowed complex exponentials, sparsified per row at the `sparsity`
    magnitude quantile (constantq.py:119-163, 219-262) -> (rowptr int32 [bpo+1], col int32 [nnz], val complex64 [nnz], n_fft)."""
    lengths = _constant_q_lengths(sr, fmin_top, bins_per_octave, bins_per_octave, filter_scale)
    freqs = fmin_top * (2.0 ** (torch.arange(bins_per_octave, dtype=torch.float) / bins_per_octave))
    max_len = int(2.0 ** (torch.ceil(torch.log2(max(lengths)))))
    rows = []
    for ilen, freq in zip(lengths, freqs):
        half = torch.div(ilen, 2, rounding_mode="floor")
        sig = torch.exp(torch.arange(-half, half, dtype=torch.float) * 1j * 2 * torch.pi * freq / sr)
        sig = sig * torch.hann_window(len(sig))
        sig = sig / sig.norm(p=1, dim=0)
        lpad = int((max_len - sig.shape[-1]) // 2)
        rows.append(torch.nn.functional.pad(sig, (lpad, int(max_len - sig.shape[-1] - lpad)), mode="constant"))
    basis = torch.stack(rows) * (lengths[:, None] / float(max_len))
    fft_basis = torch.fft.fft(basis, n=max_len, axis=1)[:, : (max_len // 2) + 1]
    mags = torch.abs(fft_basis)
    norms = torch.sum(mags, axis=1, keepdims=True)
    mag_sort = torch.sort(mags, axis=1).values
    cumulative = torch.cumsum(mag_sort / norms, axis=1)
    thr_idx = torch.argmin((cumulative < sparsity).to(torch.uint8), axis=1)
    keep = mags >= mag_sort[torch.arange(fft_basis.shape[0]), thr_idx][:, None]
    rowptr = torch.zeros(bins_per_octave + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(keep.sum(1), 0).to(torch.int32)
    col = keep.nonzero()[:, 1].to(torch.int32).contiguous()
    val = fft_basis[keep].to(torch.complex64).contiguous()
    return rowptr, col, val, max_len


def cq_to_chroma(n_input, bins_per_octave=12, n_chroma=12, fmin=None):
    """rosa/convert.py:69-117 (base_c=True, no window) -> float32 [n_chroma, n_input]."""
    n_merge = float(bins_per_octave) / n_chroma
    fmin = C1_HZ if fmin is None else float(fmin)
    m = torch.repeat_interleave(torch.eye(n_chroma), round(n_merge), dim=1)
    m = torch.roll(m, -int(n_merge // 2), dims=1)
    n_octaves = int(np.ceil(float(n_input) / bins_per_octave))
    m = torch.tile(m, (1, n_octaves))[:, :n_input]
    midi_0 = (12 * (np.log2(np.float32(fmin)) - np.log2(440.0)) + 69) % 12
    roll = int(torch.round(torch.tensor(midi_0 * (n_chroma / 12.0))))
    return torch.roll(m, roll, dims=0).to(torch.float).contiguous()


_design_cache = {}
_ws_cache = {}


def _design(sr, hop_length, fmin, n_bins, bins_per_octave, tuning, n_chroma, device):
    key = (float(sr), int(hop_length), None if fmin is None else float(fmin), n_bins, bins_per_octave, float(tuning), n_chroma, str(device))
    d = _design_cache.get(key)
    if d is None:
        f0 = torch.tensor(C1_HZ if fmin is None else float(fmin)).float() * 2.0 ** (tuning / bins_per_octave)
        top = (f0 * 2.0 ** (torch.arange(0, n_bins, dtype=torch.float) / bins_per_octave))[-bins_per_octave:]
        rowptr, col, val, n_fft = octave_filter_bank(sr, torch.min(top), bins_per_octave)
        lengths = _constant_q_lengths(sr, f0, n_bins, bins_per_octave)
        kern, width = kaiser_decimation_kernel()
        d = dict(rowptr=rowptr.to(device), col=col.to(device), val=torch.view_as_real(val).contiguous().to(device), n_fft=n_fft,
                 inv_sqrt_len=(1.0 / torch.sqrt(lengths)).contiguous().to(device), kern=kern.to(device), width=width,
                 fold=cq_to_chroma(n_bins, bins_per_octave, n_chroma, fmin).to(device))
        _design_cache[key] = d
    return d


def _run(y, sr, hop_length, fmin, n_bins, bins_per_octave, tuning, n_chroma, threshold, norm, want_cqt):
    if not y.is_cuda:
        raise RuntimeError("maua_b200 chroma features need a CUDA tensor (no CPU fallback)")
    if tuning is None:
        raise NotImplementedError("chroma_cqt: tuning=None runs rosa/pitch.py estimate_tuning in the reference; pass tuning explicitly")
    lib = _lib.load()
    y = y.detach().to(torch.float32).contiguous().reshape(-1)
    n = y.numel()
    if n % hop_length:
        raise ValueError(f"audio length must be a multiple of hop_length={hop_length}")
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    dev = y.device
    d = _design(sr, hop_length, fmin, n_bins, bins_per_octave, tuning, n_chroma, dev)
    T = n // hop_length
    with torch.cuda.device(dev):
        chroma = torch.empty(n_chroma, T, device=dev)
        cq = torch.empty(n_bins, T, device=dev) if want_cqt else None
        nbytes = lib.mb_chroma_workspace_bytes(n, n_bins, hop_length)
        ws = _ws_cache.get((nbytes, str(dev)))
        if ws is None:
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            _ws_cache.clear()
            _ws_cache[(nbytes, str(dev))] = ws
        off = (-ws.data_ptr()) % 256
        _lib.check(lib.mb_chroma_cqt(_lib.ptr(y), n, int(hop_length), int(d["n_fft"]), n_octaves, min(bins_per_octave, n_bins),
                                     _lib.ptr(d["kern"]), d["kern"].numel(), int(d["width"]), _lib.ptr(d["rowptr"]), _lib.ptr(d["col"]),
                                     _lib.ptr(d["val"]), _lib.ptr(d["inv_sqrt_len"]), _lib.ptr(d["fold"]), n_chroma,
                                     float("-inf") if threshold is None else float(threshold), int(bool(norm)), _lib.ptr(cq),
                                     _lib.ptr(chroma), C.c_void_p(ws.data_ptr() + off), nbytes, _lib.stream_ptr()))
    return chroma, cq


def chroma_cqt(y, sr, hop_length=1024, fmin=None, threshold=0.0, tuning=0.0, n_chroma=12, n_octaves=7, window=None,
               bins_per_octave=36, norm=True):
    """rosa/spectral.py:286-325 -> [n_chroma, T] (max-normalised).  `tuning` must be given (the reference's None default
    estimates it with rosa/pitch.py first); `window` (chroma smoothing) is not built."""
    if window is not None:
        raise NotImplementedError("chroma_cqt: window is not supported")
    return _run(y, sr, hop_length, fmin, n_octaves * bins_per_octave, bins_per_octave, tuning, n_chroma, threshold, norm, False)[0]


def cqt_magnitude(y, sr, hop_length=1024, fmin=None, n_bins=84, bins_per_octave=12, tuning=0.0):
    """|cqt(y, sr, ...)| of rosa/constantq.py:13-27 -> [n_bins, T] (n_bins a multiple of bins_per_octave)."""
    if n_bins % bins_per_octave:
        raise NotImplementedError("cqt_magnitude: n_bins must be a whole number of octaves")
    return _run(y, sr, hop_length, fmin, n_bins, bins_per_octave, tuning, 12, None, False, True)[1]


# ---- tuning estimate and CENS chroma: the rest of chromagram() ------------------------------------------------------
def estimate_tuning(y, sr, n_fft=2048, resolution=0.01, bins_per_octave=12):
    """rosa/pitch.py:9-24 (piptrack defaults of :27-37) -> 0-dim CUDA tensor, tuning in fractions of a bin."""
    if not y.is_cuda:
        raise RuntimeError("maua_b200 chroma features need a CUDA tensor (no CPU fallback)")
    if n_fft != 2048:
        raise NotImplementedError("estimate_tuning: only n_fft=2048 (the reference default)")
    lib = _lib.load()
    y = y.detach().to(torch.float32).contiguous().reshape(-1)
    n = y.numel()
    if n % 512:
        raise ValueError("audio length must be a multiple of 512 (n_fft // 4, the reference's piptrack hop)")
    with torch.cuda.device(y.device):
        out = torch.empty(1, device=y.device)
        nbytes = lib.mb_tuning_workspace_bytes(n)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=y.device)
        off = (-ws.data_ptr()) % 256
        _lib.check(lib.mb_estimate_tuning(_lib.ptr(y), n, float(sr), int(bins_per_octave), int(np.ceil(1.0 / resolution)), _lib.ptr(out),
                                          C.c_void_p(ws.data_ptr() + off), nbytes, _lib.stream_ptr()))
    return out[0]


def _natural_cubic_coeffs(x, y):
    """Natural cubic spline through (x_i, y_i) in float64: per interval S(t) = a + b f + c f^2 + d f^3, f = t - x_i --
    the quantities torchcubicspline.natural_cubic_spline_coeffs hands to the reference's spline_eval (spectral.py:189-203)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    h = np.diff(x)
    m = np.zeros(n)
    if n > 2:
        a_ = np.zeros((n - 2, n - 2))
        rhs = 6.0 * ((y[2:] - y[1:-1]) / h[1:] - (y[1:-1] - y[:-2]) / h[:-1])
        i = np.arange(n - 2)
        a_[i, i] = 2.0 * (h[:-1] + h[1:])
        a_[i[1:], i[1:] - 1] = h[1:-1]
        a_[i[:-1], i[:-1] + 1] = h[1:-1]
        m[1:-1] = np.linalg.solve(a_, rhs)
    a = y[:-1]
    b = (y[1:] - y[:-1]) / h - h * (2.0 * m[:-1] + m[1:]) / 6.0
    c = m[:-1] / 2.0
    d = (m[1:] - m[:-1]) / (6.0 * h)
    return a, b, c, d


def cens_quantiser_design():
    """Knots of the smooth 4-step CENS quantisation curve (spectral.py:164-188) and its natural-spline coefficients
    -> (x float32 [240], coef float32 [4, 239])."""
    steps = [0.4, 0.2, 0.1, 0.05]
    p1, p2, p3, p4 = np.diff(list(reversed(steps + [0])))
    xs = [torch.linspace(-0.1, 0.025, 101)[:-1], torch.linspace(0.025, p1, 11)[:-1], torch.linspace(p1, p1 + p2, 11)[:-1],
          torch.linspace(p1 + p2, p1 + p2 + p3, 11)[:-1], torch.linspace(p1 + p2 + p3, 0.5, 11)[:-1], torch.linspace(0.5, 1.1, 100)]
    ys = torch.cat((0.5 * torch.ones(len(xs[0])), xs[1] / p1, (xs[2] - p1) / p2 + 1, (xs[3] - p1 - p2) / p3 + 2,
                    (xs[4] - p1 - p2 - p3) / p4 + 3, 4.5 * torch.ones(len(xs[5]))))
    x = torch.cat(xs)
    coef = np.stack(_natural_cubic_coeffs(x.numpy(), ys.numpy()))
    return x.float().contiguous(), torch.from_numpy(coef).float().contiguous()


_cens_cache = {}


def chroma_cens(y, sr, hop_length=1024, fmin=None, tuning=None, n_chroma=12, n_octaves=7, bins_per_octave=36, window=None,
                win_len_smooth=41, smoothing_window=torch.hann_window):
    """rosa/spectral.py:239-280 -> [n_chroma, T]: chroma_cqt(norm=False) -> L1 -> smooth quantiser -> temporal
    smoothing -> L2.  tuning=None estimates it on the device first, as the reference does (one scalar read-back: the
    filter bank is designed for that tuning)."""
    if tuning is None:
        tuning = float(estimate_tuning(y, sr, bins_per_octave=bins_per_octave).item())
    raw = chroma_cqt(y, sr, hop_length=hop_length, fmin=fmin, threshold=0.0, tuning=tuning, n_chroma=n_chroma, n_octaves=n_octaves,
                     window=window, bins_per_octave=bins_per_octave, norm=False)
    dev = raw.device
    key = (str(dev), int(win_len_smooth), smoothing_window)
    d = _cens_cache.get(key)
    if d is None:
        kx, coef = cens_quantiser_design()
        if win_len_smooth:
            win = smoothing_window(win_len_smooth + 2)
            win = win / torch.sum(win)
        else:
            win = torch.ones(1)
        d = (kx.to(dev), coef.to(dev), win.float().contiguous().to(dev))
        _cens_cache[key] = d
    kx, coef, win = d
    T = raw.shape[1]
    with torch.cuda.device(dev):
        scratch, out = torch.empty_like(raw), torch.empty_like(raw)
        _lib.check(_lib.load().mb_chroma_cens_post(_lib.ptr(raw), n_chroma, T, _lib.ptr(kx), _lib.ptr(coef), kx.numel(), _lib.ptr(win),
                                                   win.numel(), _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr()))
    return out


def chromagram(audio, sr):
    """features/audio.py:44: chroma_cens(harmonic(audio), sr).T -> [T, 12]."""
    from .features import harmonic

    return chroma_cens(harmonic(audio), sr).T
