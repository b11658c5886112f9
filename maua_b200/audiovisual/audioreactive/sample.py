"""Random-patch audio-reactive render: mirror of maua/audiovisual/audioreactive/selfsupervised/sample.py:16-101 (`generate`).

audio file -> mono, cropped, resampled to 1024 * fps -> retrieve_music_information (device features, host tempo / beats,
Laplacian segmentations) -> random or loaded Patch -> latent palette from the StyleGAN2 mapper -> latent + per-layer noise
sequences -> batches through the synthesizer with that batch's noise maps -> uint8 frames -> frame sink (ffmpeg's stdin, or
the raw rgb24 file when no ffmpeg binary exists).  Differences from the reference: WAV input through the standard library
(no audio backend in this image), and downscale_factor / aspect_ratio must be 1 because the StyleGAN2 output-size hooks
are not built; the reference's habit of dropping the last partial batch (:85) is kept.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch

from ..patches.base import load_audio as _load_wav
from ..render._sink import RingWriter
from ..render.ffmpeg import FFMPEG
from .mir import retrieve_music_information
from .patch import Patch


def load_audio(audio_file, offset, duration, fps, device="cuda"):
    """sample.py:16-32: mono float32 at sr = 1024 * fps (one STFT hop = one video frame), trimmed to whole frames."""
    from torchaudio.functional import resample

    audio, sr, _ = _load_wav(audio_file, offset, -1 if duration is None else duration)
    new_sr = int(round(1024 * fps))
    audio = resample(audio.to(device), sr, new_sr)
    return audio[: (len(audio) // 1024) * 1024].contiguous(), new_sr


@torch.inference_mode()
def generate(audio_file: str, stylegan2_checkpoint: Optional[str] = None, patch_file: Optional[str] = None, seed: Optional[int] = None,
             latent_seeds: Optional[str] = None, fps: float = 30, audio_offset: float = 0, audio_duration: Optional[float] = None,
             downscale_factor: float = 4, aspect_ratio: float = 1, batch_size: int = 32, device: str = "cuda",
             out_file: Optional[str] = None, sink=None):
    """Returns (out_file, frames written, patch).  `sink`: optional object with write(bytes-like) that receives the rgb24
    stream instead of ffmpeg / the raw file (tests, custom encoders)."""
    from ...GAN.wrappers.stylegan2 import StyleGAN2

    if seed is None:
        seed = torch.randint(0, 2 ** 32, size=(), device=device).item()
    audio, sr = load_audio(audio_file, audio_offset, audio_duration, fps, device)
    features, segmentations, tempo = retrieve_music_information(audio, sr, device=device)

    if patch_file is None:
        patch = Patch(features=features, segmentations=segmentations, tempo=tempo, seed=seed, fps=fps, device=device)
    else:
        patch = Patch.load(patch_file, features=features, segmentations=segmentations, tempo=tempo, fps=fps, device=device)

    # sample.py:53,70: the output size follows from the two factors and the wrapper's output-size hook (default: "stretch" on
    # layer 0, i.e. a resized constant input) renders it
    out_size = (round(aspect_ratio * 1024 / downscale_factor), round(1024 / downscale_factor))
    G = StyleGAN2(model_file=stylegan2_checkpoint, output_size=out_size).to(device)
    if out_file is None:
        out_file = f"output/{Path(audio_file).stem}_RandomPatches++_seed{seed}_{out_size[0]}x{out_size[1]}.mp4"
    if latent_seeds is None:
        z = torch.randn((180, 512), device=device, generator=torch.Generator(device).manual_seed(seed))
        latent_palette = G.mapper(z)
    else:
        latent_palette = G.get_w_latents(latent_seeds)
    latents, noise = patch.forward(latent_palette.to(device).float(), downscale_factor=downscale_factor, aspect_ratio=aspect_ratio)
    n_layers = len(G.synthesizer.layer_names) - 1   # per-frame noise maps the network takes (bs.0.conv1 ... last conv1)

    renderer = FFMPEG(out_file, fps=fps, audio_file=audio_file, audio_offset=audio_offset, audio_duration=audio_duration,
                      batch_size=batch_size)
    proc, writer, written = None, None, 0
    try:
        for i in range(0, len(latents) - batch_size, batch_size):
            L = latents[i: i + batch_size]
            N = {f"noise{j}": module.forward(i, batch_size)[:, None] for j, module in enumerate(noise[:n_layers])}
            u8 = G.synthesizer(latents=L, out_fmt="u8", **N)            # uint8 NHWC: the rgb24 wire format, fused in the last kernel
            if writer is None:
                if sink is None:
                    Path(out_file).parent.mkdir(parents=True, exist_ok=True)
                    sink_obj, proc = renderer._open_sink(u8.shape[2], u8.shape[1])
                else:
                    sink_obj = sink
                ring = [torch.empty((batch_size,) + tuple(u8.shape[1:]), dtype=torch.uint8).pin_memory() for _ in range(3)]
                writer = RingWriter(sink_obj, ring)
            k = writer.acquire()
            writer.ring[k][: u8.shape[0]].copy_(u8, non_blocking=True)
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream())
            writer.submit(k, u8.shape[0], event)
            written += u8.shape[0]
            if i == 0 and sink is None:
                patch.save(out_file.replace(".mp4", ".json"))
    finally:
        try:
            if writer is not None:
                writer.close()
        finally:
            if writer is not None and sink is None:
                writer.sink.close()
            if proc is not None:
                proc.wait()
    return out_file, written, patch
