"""Torch-native ("selfsupervised") envelope / latent / noise patch functions on the device: mirror of
maua/audiovisual/audioreactive/selfsupervised/{features/processing.py, mir.py, latent.py, noise.py} (rows a7, a10, a11,
a12 of SURVEY §8a), same names and arguments.  CUDA tensors only; the per-element arithmetic runs in the library's
kernels (csrc/signal_ops.cu, csrc/sequencers.cu), order statistics (quantile / sort / randperm) are torch device ops."""
from __future__ import annotations

import torch

from ... import _lib
from . import noise as _noise
from .latent import single_weighted
from .signal import _cuda32

_PAD = {"circular": 0, "reflect": 1}


def gaussian_filter(x, sigma, mode: str = "circular", causal: float = 1):
    """features/processing.py:11-50: temporal Gaussian, `mode` padding; the causal factor is commented out upstream (:24),
    so `causal` is accepted and ignored exactly like there.  Any trailing shape ([T], [T,C], [T,C,L], [T,C,H,W]); like the
    reference (:47-48) a 2-D input comes back squeezed ([T,1] -> [T])."""
    if mode not in _PAD:
        raise NotImplementedError(f"gaussian_filter: padding mode '{mode}' (built: circular, reflect)")
    x = _cuda32(x, "x")
    T = x.shape[0]
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_gaussian_filter_ex(_lib.ptr(x), _lib.ptr(y), T, x.numel() // T, float(sigma), 0, 0.0, _PAD[mode],
                                                     _lib.stream_ptr()))
    return y.squeeze() if y.dim() == 2 else y


def normalize(array):
    """features/processing.py:53-56: (x - min) / (max(x - min) + 1e-8)."""
    from .signal import normalize as _normalize

    return _normalize(array, eps=1e-8)


def quantile(tensor, q):
    """features/efficient_quantile/__init__.py:6-7: mid-point quantile of the flattened tensor (NaNs ignored, q rounded to
    float32 as the reference's FloatTensor([q]) does) -> 0-d tensor.  The reference copies to the host and partially sorts
    there; here a radix select on the device (csrc/signal_ops.cu::quantile_mid_kernel)."""
    x = _cuda32(tensor, "tensor").flatten().contiguous()
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_quantile_mid(_lib.ptr(x), x.numel(), float(q), _lib.ptr(out), _lib.stream_ptr()))
    return out[0].to(tensor.dtype)


def standardize(array):
    """features/processing.py:58-61: clamp to the inter-quartile range (+1e-10 on the upper bound), then normalize."""
    array = _cuda32(array, "array")
    return normalize(torch.clamp(array, quantile(array, 0.25), quantile(array, 0.75) + 1e-10))


def spectral_flux(spec):
    """features/processing.py:88-89: forward difference along time, last row against zero."""
    spec = _cuda32(spec, "spec")
    return torch.diff(spec, dim=0, append=torch.zeros((1, spec.shape[1]), device=spec.device))


def onset_envelope(flux):
    """features/processing.py:93-98: half-wave rectified flux summed over bins, clamped to its 2.5 % .. 97.5 % quantiles,
    shifted and scaled to [0, 1]."""
    flux = _cuda32(flux, "flux")
    u = torch.sum(0.5 * (flux + torch.abs(flux)), dim=1)
    u = torch.clamp(u, quantile(u, 0.025), quantile(u, 0.975))
    u = u - u.min()
    return u / u.max()


def salience_weighted(envelope, short_sigma=5, long_sigma=80):
    """mir.py:13-21: (gaussian(short) / gaussian(long))^2 * envelope, reflect padding -> [T, 1]."""
    env = _cuda32(envelope, "envelope")
    if env.dim() > 1:
        env = env.squeeze(1)
    env = env.contiguous()
    short = gaussian_filter(env, short_sigma, mode="reflect", causal=0)
    long = gaussian_filter(env, long_sigma, mode="reflect", causal=0)
    out = torch.empty_like(env)
    with torch.cuda.device(env.device):
        _lib.check(_lib.load().mb_salience(_lib.ptr(short), _lib.ptr(long), _lib.ptr(env), _lib.ptr(out), env.numel(), _lib.stream_ptr()))
    return out.unsqueeze(1) if out.dim() < 2 else out


def _peaks(sig):
    n = sig.shape[0]
    i = torch.arange(n, device=sig.device)
    return (sig > sig[(i + 1).clamp(0, n - 1)]) & (sig > sig[(i - 1).clamp(0, n - 1)])


def clamp_peaks_percentile(signal, percent):
    """features/processing.py:102-122: clamp each column to the `percent` quantile of its strict local maxima."""
    signal = _cuda32(signal, "signal")
    if signal.dim() < 2:
        signal = signal.unsqueeze(1)
    cols = [torch.clamp(sig, None, torch.quantile(sig[_peaks(sig)], percent / 100)) for sig in signal.unbind(1)]
    return torch.stack(cols, dim=1)


def clamp_upper_percentile(signal, percentile):
    signal = _cuda32(signal, "signal")
    return torch.clamp(signal, None, torch.quantile(signal, percentile / 100, dim=0))


def clamp_lower_percentile(signal, percentile):
    signal = _cuda32(signal, "signal")
    return torch.clamp(signal, torch.quantile(signal, percentile / 100, dim=0), None)


def emphasize(envs, strength, percentile):
    """features/processing.py:133-139."""
    envs = _cuda32(envs, "envs")
    lo = envs.min(dim=0).values
    x = envs - lo
    hi = x.max(dim=0).values
    x = x / hi
    x = x * (1 + torch.tanh(strength * (x - torch.quantile(x, q=percentile / 100, dim=0))))
    return (x * hi) + lo


def drop_strength(audio, sr):
    """features/audio.py:38-39: emphasize(gaussian_filter(rms(audio, sr), 10), strength=10, percentile=50) -> [T, 1]."""
    from .features import rms

    return emphasize(gaussian_filter(rms(audio, sr), 10), strength=10, percentile=50).unsqueeze(1)


def tonnetz(y, sr, chroma_fn=None):
    """features/audio.py:46-56: tonal centroid features, 6 x 12 projection of the L1-normalised chromagram -> [T, 6].
    chroma_fn(y, sr) -> [12, T] (default: the device chromagram, transposed)."""
    if chroma_fn is None:
        from .chroma import chromagram

        chroma_fn = lambda a, sr: chromagram(a, sr).T
    chroma = _cuda32(chroma_fn(y, sr), "chroma")
    n = chroma.shape[0]
    dim_map = torch.linspace(0, 12, n, device=chroma.device)
    scale = torch.tensor([7.0 / 6, 7.0 / 6, 3.0 / 2, 3.0 / 2, 2.0 / 3, 2.0 / 3], device=chroma.device)
    V = scale.reshape(-1, 1) * dim_map
    V[::2] -= 0.5
    R = torch.tensor([1, 1, 1, 1, 0.5, 0.5], device=chroma.device)
    phi = (R[:, None] * torch.cos(torch.pi * V)).t().contiguous()          # [12, 6]: host-sized design matrix
    frames = (chroma / chroma.norm(p=1, dim=0)).t().contiguous()            # [T, 12]
    out = torch.empty(frames.shape[0], 6, device=chroma.device)
    with torch.cuda.device(chroma.device):  # out[T, 6] = frames[T, 12] @ phi[12, 6]: the mixing kernel of the noise sequencers
        _lib.check(_lib.load().mb_noise_mix(_lib.ptr(phi), _lib.ptr(frames), frames.shape[0], n, 6, 0, _lib.ptr(out), _lib.stream_ptr()))
    return out


def plp(y, sr, hop_length=1024, win_length=1024, tempo_min=60, tempo_max=180):
    """Predominant local pulse (features/rosa/beat.py:41-75) -> [T] in [0, 1].  The mel power spectrogram comes from the fused
    STFT kernel; the tempogram is a Fourier transform over a [T] envelope with hop 1 (T x 513 bins, once per track): torch's
    device stft / istft, as in the reference."""
    from .features import _spectrogram

    if hop_length != 1024:
        raise NotImplementedError("plp: the device STFT is built for hop_length = 1024 (one hop per video frame)")
    mel = _spectrogram(y, sr, False, True, mel_fmax=11025.0)[1]                       # [T, 128] power
    s = 10.0 * torch.log10(torch.clamp(mel.t(), min=1e-10))
    s = torch.maximum(s, s.max() - 80.0)
    env = torch.clamp(s[:, 1:] - s[:, :-1], min=0).median(dim=0).values              # onset_strength, median over the bands
    env = torch.nn.functional.pad(env, (2, 0))[: s.shape[1]]
    n = min(len(env), win_length)
    win = torch.hann_window(n, device=env.device)
    ft = torch.stft(env, n_fft=n, hop_length=1, center=True, window=win, pad_mode="reflect", return_complex=True)
    freqs = torch.linspace(0, float(sr * 60 / float(hop_length)) / 2, int(1 + n // 2), device=env.device)
    if tempo_min is not None:
        ft[freqs < tempo_min] = 0
    if tempo_max is not None:
        ft[freqs > tempo_max] = 0
    mag = torch.log1p(1e6 * torch.abs(ft))
    ft[mag < mag.max(dim=0, keepdim=True).values] = 0
    ft = ft / (torch.finfo(ft.dtype).tiny ** 0.5 + torch.abs(ft.abs().max(dim=0, keepdim=True).values))
    pulse = torch.istft(ft, n_fft=n, hop_length=1, center=True, window=win, length=len(env))
    pulse = torch.clamp(pulse, torch.zeros((), device=pulse.device), pulse.max())
    return normalize(pulse)


def pulse(audio, sr):
    """features/audio.py:67-68: plp(percussive(audio), sr) -> [T, 1]."""
    from .features import percussive

    return plp(percussive(audio), sr).unsqueeze(-1)


def spline_loop_latents(y, size, n_loops=1):
    """latent.py:7-13: natural cubic spline through cat(y, y[0]) walked n_loops (may be fractional) times."""
    lat = _cuda32(y, "y")
    K, Cc = lat.shape[0], lat[0].numel()
    out = torch.empty((size,) + tuple(lat.shape[1:]), device=lat.device)
    ws = torch.empty((K + 1) * Cc, device=lat.device)
    with torch.cuda.device(lat.device):
        _lib.check(_lib.load().mb_spline_loop_latents(_lib.ptr(lat), K, Cc, float(n_loops), int(size), _lib.ptr(out), _lib.ptr(ws),
                                                      _lib.stream_ptr()))
    return out


_LATENT_LAYERS = {"low": (0, 6), "mid": (6, 12), "high": (12, 18), "lowmid": (0, 12), "midhigh": (6, 18), "all": (0, 18)}


def latent_patch(rng, latents, palette, segmentations, features, tempo, fps, patch_type, segments, loop_bars, seq_feat,
                 seq_feat_weight, mod_feat, mod_feat_weight, merge_type, merge_depth):
    """latent.py:16-80: one random-patch step on the W+ sequence `latents` [T, num_ws, 512] (modified in place and
    returned, as in the reference)."""
    if not latents.is_cuda or latents.dtype != torch.float32 or not latents.is_contiguous():
        raise RuntimeError("latent_patch: latents must be a contiguous float32 CUDA tensor (it is updated in place)")
    palette = _cuda32(palette, "palette")
    feature = seq_feat_weight * _cuda32(features[seq_feat], seq_feat)
    segmentation = segmentations[(seq_feat, segments)]
    permutation = torch.randperm(len(palette), generator=rng, device=rng.device).to(palette.device)
    lib = _lib.load()
    T, L, D = latents.shape

    if patch_type == "segmentation":
        selection = permutation[:segments]
        selectseq = selection[segmentation.to(selection.device)]
        sequence = gaussian_filter(palette[selectseq], 5)
    elif patch_type == "feature":
        n_select = feature.shape[1]
        if n_select == 1:
            selection = permutation[:2]
            sequence = single_weighted(palette[selection[1]], palette[selection[0]], feature.reshape(-1))
        else:
            selection = permutation[:n_select]
            pal = palette[selection].reshape(n_select, -1).contiguous()
            sequence = torch.empty((T,) + tuple(palette.shape[1:]), device=palette.device)
            with torch.cuda.device(palette.device):  # einsum("TN,NWL->TWL") = the mixing kernel of the noise sequencers
                _lib.check(lib.mb_noise_mix(_lib.ptr(pal), _lib.ptr(feature.contiguous()), T, n_select, pal.shape[1], 0,
                                            _lib.ptr(sequence), _lib.stream_ptr()))
    elif patch_type == "loop":
        selection = permutation[:segments]
        n_loops = len(latents) / fps / 60 / tempo / 4 / loop_bars
        sequence = spline_loop_latents(palette[selection], len(latents), n_loops=n_loops)
    else:
        raise ValueError(f"latent_patch: unknown patch_type '{patch_type}'")
    sequence = gaussian_filter(sequence, 1)

    lay0, lay1 = _LATENT_LAYERS[merge_depth]
    mode = {"average": 0, "modulate": 1}.get(merge_type, 2)
    modulation = None
    if mode == 1:
        modulation = (mod_feat_weight * _cuda32(features[mod_feat], mod_feat)).reshape(-1).contiguous()
        if modulation.numel() != T:
            raise ValueError("latent_patch: a modulation feature must have one value per frame")
    with torch.cuda.device(latents.device):
        _lib.check(lib.mb_latent_merge(_lib.ptr(latents), _lib.ptr(sequence.contiguous()), _lib.ptr(modulation), mode, lay0, lay1, T, L, D,
                                       _lib.stream_ptr()))
    return latents


_NOISE_LAYERS = {"low": range(0, 6), "mid": range(6, 12), "high": range(12, 17), "lowmid": range(0, 12), "midhigh": range(6, 17),
                 "all": range(0, 17)}


def noise_patch(rng, noise, features, tempo, fps, patch_type, loop_bars, seq_feat, seq_feat_weight, mod_feat, mod_feat_weight,
                merge_type, merge_depth, noise_mean, noise_std):
    """noise.py:89-140: wrap the per-layer noise sequencers of `noise` (list) with one random-patch step."""
    lays = _NOISE_LAYERS[merge_depth]
    feature = seq_feat_weight * features[seq_feat]
    for n in lays:
        if patch_type == "blend":
            new_noise = _noise.Blend(rng=rng, length=len(feature), size=noise[n].size, modulator=feature)
        elif patch_type == "multiply":
            new_noise = _noise.Multiply(rng=rng, length=len(feature), size=noise[n].size, modulator=feature)
        elif patch_type == "loop":
            n_loops = len(feature) / fps / 60 / tempo / 4 / loop_bars
            new_noise = _noise.Loop(rng=rng, length=len(feature), size=noise[n].size, n_loops=n_loops, device=feature.device)
        else:
            raise ValueError(f"noise_patch: unknown patch_type '{patch_type}'")
        if merge_type == "average":
            noise[n] = _noise.Average(left=noise[n], right=new_noise)
        elif merge_type == "modulate":
            noise[n] = _noise.Modulate(left=noise[n], right=new_noise, modulator=mod_feat_weight * features[mod_feat])
        else:  # overwrite
            noise[n] = new_noise
        noise[n] = _noise.ScaleBias(noise[n], scale=noise_std, bias=noise_mean)
    return noise


# ---- the feature half of retrieve_music_information (mir.py:9-11, 24-25, 43) ------------------------------------------------
UNITFEATS = ["rms", "drop_strength", "onsets", "spectral_flatness"]
ALLFEATS = ["chromagram", "tonnetz", "mfcc", "spectral_contrast"] + UNITFEATS


def audio_feature_functions():
    """AFEATFNS of mir.py:9 as device functions, in the reference's order."""
    from .chroma import chromagram
    from .features import mfcc, onsets, rms, spectral_contrast, spectral_flatness

    return [chromagram, tonnetz, mfcc, spectral_contrast, spectral_flatness, rms, drop_strength, onsets]


def postprocess_feature(af):
    """mir.py:43: normalize(salience_weighted(gaussian_filter(af, sigma=2)))."""
    return normalize(salience_weighted(gaussian_filter(af, sigma=2)))


def extract_features(audio, sr, postprocess=True):
    """{name: [T, C]} for the eight audio features of mir.py:9 (raw, or post-processed as retrieve_music_information returns
    them).  The segmentations and the tempo estimate of that function (Laplacian segmentation, librosa beat tracking) are
    not built: pass them to Patch from elsewhere."""
    feats = {fn.__name__: fn(audio, sr) for fn in audio_feature_functions()}
    return {k: postprocess_feature(v) for k, v in feats.items()} if postprocess else feats
