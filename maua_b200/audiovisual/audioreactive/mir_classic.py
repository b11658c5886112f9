"""The classic ``ar.*`` music-information functions: mirror of maua/audiovisual/audioreactive/mir.py:17-176 (onsets, volume,
chroma, tonnetz, spectral_max, pitch_dominance, pulse, tempo) -- same names, arguments and return containers.

The reference computes these on the host with librosa and madmom, two absent, un-pinned third-party packages
(setup.py:60,62), so there is no reference arithmetic to reproduce bit for bit: PARITY UNPINNED for this module.  Each
function runs the corresponding IN-TREE torch-native arithmetic of the reference (selfsupervised/features/audio.py and
rosa/*, which this build pins bit for bit or to fp32 rounding) on the device, with the classic function's post-processing
around it (percentile clip, min-max scaling, nearest-neighbour filtering, note selection).  Framing is the in-tree twin's
(n_fft 2048, hop 1024) instead of librosa's hop of 512: envelopes come out at sr / 1024 frames per second; patch files
resample them to the video's frame count with ``ar.resample``, as they do with librosa's.

Host arrays go in and come out (numpy / CPU tensors as the reference returns them); CUDA tensors stay on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import chroma as _chroma
from . import features as _f
from . import signal as _signal
from .audio import _pad_to_hop, _to_device


def _frames(audio):
    y, restore = _to_device(audio)
    return _pad_to_hop(y.float()), restore


def _host(t, like):
    """Return container of the reference for a device result: CUDA input -> CUDA tensor, host input -> CPU tensor."""
    return t if (torch.is_tensor(like) and like.is_cuda) else t.cpu()


def onsets(audio, sr, type="mm", prepercussive=4):
    """mir.py:17-61 -> float32 [n_frames] onset envelope, percentile-clipped at 95 %.  Both detector families of the reference
    (librosa's onset_strength, madmom's five spectral detectors) are absent; the in-tree mel-flux onset strength
    (rosa/beat.py:10-23) of the percussive component runs instead, for either ``type``."""
    y, _ = _frames(audio)
    if prepercussive:
        y = _f.percussive(y, margin=float(prepercussive))
    mel = _f._spectrogram(y, sr, False, True, mel_fmax=11025.0)[1]
    s = 10.0 * torch.log10(torch.clamp(mel.t(), min=1e-10))
    s = torch.maximum(s, s.max() - 80.0)
    env = torch.clamp(s[:, 1:] - s[:, :-1], min=0).mean(dim=0)
    env = torch.nn.functional.pad(env, (2, 0))[: s.shape[1]]
    env = _signal.percentile_clip(env, 95)
    return _host(env.reshape(-1).float(), audio)


def volume(audio, sr):
    """mir.py:65-77: RMS envelope scaled to [0, 1] -> float32 [n_frames]."""
    y, _ = _frames(audio)
    vol = _f.rms(y, sr)[:, 0]
    vol = vol - vol.min()
    return _host(vol / vol.max(), audio)


def _nn_filter_median(ch):
    """librosa.decompose.nn_filter(ch, aggregate=np.median, metric="cosine") on [T, 12] frames (mir.py:117): every frame
    becomes the element-wise median of its nearest neighbours in cosine distance (k = 2 * ceil(sqrt(T - 1)), itself excluded,
    librosa's recurrence_matrix defaults restated)."""
    T = ch.shape[0]
    if T < 4:
        return ch
    unit = ch / ch.norm(dim=1, keepdim=True).clamp_min(1e-12)
    dist = 1.0 - unit @ unit.t()
    dist.fill_diagonal_(float("inf"))
    k = int(min(T - 1, 2 * np.ceil(np.sqrt(T - 1))))
    idx = dist.topk(k, dim=1, largest=False).indices                    # [T, k]
    return ch[idx].median(dim=1).values                                 # [T, k, 12] -> [T, 12]


def chroma(audio, sr, type="cens", nearest_neighbor=True, preharmonic=4, notes=12):
    """mir.py:81-122 -> float32 numpy [n_frames, notes] (the reference returns numpy here), min-max scaled.  ``type``: "cens"
    and "cqt" are the in-tree constant-Q chromagrams; "stft" / "deep" / "clp" fall back to "cens" with the reference's own
    message for an unknown type."""
    y, _ = _frames(audio)
    if preharmonic:
        y = _f.harmonic(y, margin=float(preharmonic))
    if type == "cqt":
        ch = _chroma.chroma_cqt(y, sr, tuning=None).t()
    else:
        if type != "cens":
            print("chroma type not recognized, options are: [cens, cqt, deep, clp, or stft]. defaulting to cens...")
        ch = _chroma.chroma_cens(y, sr).t()
    ch = ch.contiguous()
    if nearest_neighbor:
        ch = torch.minimum(ch, _nn_filter_median(ch))
    if notes < 12:
        ch = ch[:, torch.argsort(-ch.sum(0))[:notes]]
    ch = ch - ch.min()
    ch = ch / (ch.max() + 1e-8)
    out = ch.float()
    return out if (torch.is_tensor(audio) and audio.is_cuda) else out.cpu().numpy()


def _tonnetz_matrix(device):
    """The 6 x 12 tonal-centroid projection (features/audio.py:46-56, librosa.feature.tonnetz)."""
    dim_map = torch.linspace(0, 12, 12, device=device)
    scale = torch.tensor([7.0 / 6, 7.0 / 6, 3.0 / 2, 3.0 / 2, 2.0 / 3, 2.0 / 3], device=device)
    V = scale.reshape(-1, 1) * dim_map
    V[::2] -= 0.5
    R = torch.tensor([1, 1, 1, 1, 0.5, 0.5], device=device)
    return R[:, None] * torch.cos(torch.pi * V)


def tonnetz(audio, sr, type="cens", nearest_neighbor=True, preharmonic=4):
    """mir.py:126-132 -> float32 [n_frames, 6], min-max scaled."""
    ch = chroma(audio, sr, type=type, nearest_neighbor=nearest_neighbor, preharmonic=preharmonic)
    ch = torch.as_tensor(ch).to(_to_device(audio)[0].device)
    frames = ch / ch.norm(p=1, dim=1, keepdim=True).clamp_min(1e-12)
    ton = frames @ _tonnetz_matrix(ch.device).t()
    ton = ton - ton.min()
    return _host((ton / ton.max()).float(), audio)


def spectral_max(audio, sr, n_mels=512):
    """mir.py:144-150: per-frame maximum of the mel power spectrogram, min-max scaled -> float32 [n_frames]."""
    y, _ = _frames(audio)
    mag = _f._spectrogram(y, sr, True, False)[0]                         # [T, 1025]
    fb = _f.mel_filterbank(sr, n_mels=n_mels).to(mag.device)             # [n_mels, 1025]
    spectrum = ((mag * mag) @ fb.t()).amax(dim=1)
    spectrum = spectrum - spectrum.min()
    return _host(spectrum / spectrum.max(), audio)


def pitch_dominance(audio, sr, type="cens", nearest_neighbor=True, preharmonic=4):
    """mir.py:153-159: pitch classes sorted by their share of the chromagram, most dominant first -> int64 [12]."""
    ch = torch.as_tensor(chroma(audio, sr, type=type, nearest_neighbor=nearest_neighbor, preharmonic=preharmonic))
    norm = ch / ch.sum(dim=1, keepdim=True).clamp_min(1e-12)
    return torch.argsort(norm.sum(dim=0), descending=True).cpu()


def pulse(audio, sr, prior="lognorm", type="mm", prepercussive=4):
    """mir.py:162-176: predominant local pulse of the onset envelope, max-normalised -> float32 [n_frames].  The in-tree plp
    (rosa/beat.py:41-75) restricts tempi to 60..180 bpm instead of weighting them with a scipy.stats prior."""
    from .selfsupervised import plp

    y, _ = _frames(audio)
    if prepercussive:
        y = _f.percussive(y, margin=float(prepercussive))
    pul = plp(y, sr)
    return _host((pul / pul.abs().max().clamp_min(1e-12)).float(), audio)


def round_to_nearest_half(number):
    return round(number * 2) / 2


def tempo(audio, sr, prior="uniform", type="mm", prepercussive=4):
    """mir.py:183-210 -> [bpm, ...]: the global tempo estimate first, then the tempi of the strongest autocorrelation peaks of
    the onset envelope folded into 80..200 bpm, all rounded to the nearest half."""
    from . import beat

    env = onsets(audio, sr, type=type, prepercussive=prepercussive)
    env = env.cpu().numpy().astype(np.float64)
    fps = sr / 1024.0
    ac = np.correlate(env, env, mode="full")[len(env) - 1:][:512]
    ac = ac / (np.abs(ac).max() + 1e-12)
    peaks = np.argsort(-ac)[:10]
    peaks = peaks[(peaks > 3) & (peaks < len(ac))]
    tempos_ac = 60.0 * fps / peaks
    for t in range(len(tempos_ac)):
        while tempos_ac[t] < 80:
            tempos_ac[t] *= 2
        while tempos_ac[t] > 200:
            tempos_ac[t] /= 2
    bpm = float(np.squeeze(beat.tempo(env, sr=sr, hop_length=1024)))
    return [round_to_nearest_half(b) for b in (bpm, *tempos_ac)]


def pitch_track(audio, sr, preharmonic=4):
    raise NotImplementedError("pitch_track (librosa.piptrack averaged per frame, mir.py:135-141) is not built; "
                              "maua_b200.audiovisual.audioreactive.chroma.estimate_tuning holds the device piptrack")
