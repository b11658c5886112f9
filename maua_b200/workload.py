"""Synthetic workloads of BASELINE.json: sine-sweep audio and the audio-reactive latent
sequence that drives the generator.  Host-side setup code (numpy); nothing here is timed by bench.py.
"""
from __future__ import annotations

import numpy as np
import torch


def sine_sweep(duration_s: float, sr: int = 48000, f0: float = 20.0, f1: float = 20000.0, tremolo_hz: float = 0.0):
    """y[n] = 0.8 sin(2 pi (f0 t + (f1-f0) t^2 / (2 D))), optional amplitude tremolo."""
    t = np.arange(int(round(duration_s * sr)), dtype=np.float64) / sr
    y = 0.8 * np.sin(2 * np.pi * (f0 * t + (f1 - f0) * t * t / (2 * duration_s)))
    if tremolo_hz > 0:
        y = y * (0.55 + 0.45 * np.sin(2 * np.pi * tremolo_hz * t))
    return y.astype(np.float32), sr


def key_latents(num_ws: int, seeds=range(1, 13), z_dim: int = 512):
    """Key latents as maua/GAN/wrappers/stylegan.py:58-69 draws them (RandomState(seed).randn), broadcast
    over num_ws (random-init mappers are meaningless, SURVEY §8d uses w = z)."""
    z = np.concatenate([np.random.RandomState(s).randn(1, z_dim) for s in seeds]).astype(np.float32)
    return torch.from_numpy(z)[:, None, :].repeat(1, num_ws, 1)


def envelope_latents(w_keys: torch.Tensor, chroma: torch.Tensor, onsets: torch.Tensor, n_loops: int = 4):
    """[T,num_ws,512] sequence: chroma-weighted mix of the key latents (latent.py:21-31 multi_weighted)
    blended by the onset envelope with a looping interpolation through the keys."""
    T = chroma.shape[0]
    K = w_keys.shape[0]
    wts = chroma / (chroma.sum(1, keepdim=True) + 1e-8)
    tonal = torch.einsum("tn,nwl->twl", wts, w_keys[: chroma.shape[1]])
    pos = torch.linspace(0, n_loops * K, T + 1)[:-1]
    i0 = pos.floor().long() % K
    i1 = (i0 + 1) % K
    frac = (pos - pos.floor())[:, None, None]
    loop = w_keys[i0] * (1 - frac) + w_keys[i1] * frac
    o = onsets.reshape(T, 1, 1)
    return (o * tonal + (1 - o) * loop).contiguous()


def host_envelopes(audio: np.ndarray, sr: int, n_frames: int):
    """Cheap host envelopes for workloads that only need *some* audio-reactive drive: per-frame RMS
    (normalised) as 'onsets' and a 12-bin pseudo-chroma from the sweep's instantaneous position."""
    hop = len(audio) // n_frames
    fr = audio[: hop * n_frames].reshape(n_frames, hop)
    rms = np.sqrt((fr.astype(np.float64) ** 2).mean(1))
    rms = (rms - rms.min()) / (rms.max() - rms.min() + 1e-8)
    zc = (np.diff(np.signbit(fr), axis=1) != 0).sum(1) * sr / (2.0 * hop)  # zero-crossing frequency estimate
    pitch = 12 * np.log2(np.maximum(zc, 1.0) / 440.0)
    chroma = np.zeros((n_frames, 12), dtype=np.float32)
    for k in range(12):
        d = np.abs(((pitch - k + 6) % 12) - 6)
        chroma[:, k] = np.exp(-0.5 * (d / 0.8) ** 2)
    return torch.from_numpy(rms.astype(np.float32)), torch.from_numpy(chroma)


def c2_latents(num_ws: int = 16, duration_s: float = 30.0, fps: int = 24):
    """Config 2 of BASELINE.json: 30 s @ 24 fps = 720 audio-reactive W+ latents."""
    audio, sr = sine_sweep(duration_s, tremolo_hz=4.0)
    T = int(round(duration_s * fps))
    onsets, chroma = host_envelopes(audio, sr, T)
    return envelope_latents(key_latents(num_ws), chroma, onsets), (audio, sr)


def resampled_sweep(duration_s: float, fps: int, device):
    """The job's audio as the torch-native entry point prepares it (selfsupervised/sample.py:16-32): 48 kHz tremolo sweep,
    resampled to sr = 1024 * fps with torchaudio.functional.resample (one hop per video frame) -> (float32 [T * 1024] on
    `device`, sr)."""
    import torchaudio

    audio, sr_file = sine_sweep(duration_s, tremolo_hz=4.0)
    sr = 1024 * fps
    T = int(round(duration_s * fps))
    y = torchaudio.functional.resample(torch.from_numpy(audio).to(device), sr_file, sr)
    n = T * 1024
    y = y[:n] if y.numel() >= n else torch.nn.functional.pad(y, (0, n - y.numel()))
    return y.contiguous(), sr


def audio_reactive_latents(y, sr, num_ws: int, n_loops: int = 4):
    """SURVEY §8d's latent recipe on the device, from the resampled track `y`: onsets / rms (mb_audio_onsets_rms) and the
    chromagram (harmonic -> tuning estimate -> constant-Q -> CENS) -> ``multi_weighted(w, chroma)`` mixed by the onset envelope
    with ``spline_loops(w, T, n_loops)``, then ``gaussian_filter(sigma=2)`` over time -> [T, num_ws, 512]."""
    from .audiovisual import audioreactive as ar

    onsets, _ = ar.onsets_rms(y, sr)
    chroma = ar.chromagram(y, sr).contiguous()                 # [T, 12]
    T = onsets.shape[0]
    keys = key_latents(num_ws).to(y.device)
    drive = ar.normalize(ar.gaussian_filter(onsets[:, 0], 2.0), eps=1e-8).reshape(T, 1, 1)
    tonal = ar.multi_weighted(keys, chroma + 1e-6)
    loops = ar.spline_loops(keys, T, n_loops)
    return ar.gaussian_filter(drive * tonal + (1 - drive) * loops, 2.0).contiguous()


def job_latents_device(num_ws: int, device, duration_s: float = 30.0, fps: int = 24):
    """Latents of a BASELINE.json job (configs[1]: 30 s @ 24 fps = 720 frames; configs[2]: 180 s @ 60 fps = 10 800 frames) with the
    whole audio-reactive part computed ON THE DEVICE.  Returns (latents [T,num_ws,512] on `device`, info dict with the device
    time of the feature + sequencing pass)."""
    y, sr = resampled_sweep(duration_s, fps, device)
    audio_reactive_latents(y, sr, num_ws)                      # warm-up (allocations, module load, filter design)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lat = audio_reactive_latents(y, sr, num_ws)
    e1.record()
    torch.cuda.synchronize()
    return lat, {"audio_features_ms": e0.elapsed_time(e1), "audio_samples": int(y.numel()), "sr": sr}


def c2_latents_device(num_ws: int, device, duration_s: float = 30.0, fps: int = 24):
    return job_latents_device(num_ws, device, duration_s, fps)
