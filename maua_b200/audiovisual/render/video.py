"""``VideoWriter`` / ``write_video``: the frame-by-frame sink API of maua/ops/video.py:107-155 on the ring
writer of this package.

    with VideoWriter(output_file, output_size=(w, h), fps=24, audio_file=...) as video:
        video.write(frame)          # [1, C, H, W] (or [B, C, H, W]) in `value_range`

The reference queues tensors to a WriteWorker thread that runs tensor2bytes (ops/io.py:47-70) and writes to ffmpeg's stdin
through the absent ffmpeg-python package.  Here `write` converts to the rgb24 wire format where the frame lives (on the
device for CUDA tensors: clamp, rescale, round, uint8, HWC), copies into a ring of host buffers (pinned + asynchronous for
CUDA frames) and a writer thread feeds the sink: the `ffmpeg` binary when present, else `output_file + ".rgb24"`, or a
caller-supplied object with write().  Frames with an odd height or width are resampled to even sizes like the reference
(:89-92) -- on the device; host tensors with odd sizes are rejected (no CPU resampler in this package).
"""
from __future__ import annotations

import shutil
import subprocess
from math import ceil

import numpy as np
import torch

from ._sink import RingWriter


def tensor2bytes_device(tensor, value_range=(0, 1)):
    """tensor2bytes of ops/io.py:47-70 up to the host copy: [B, C, H, W] float -> uint8 [B, H, W, C] on the same device."""
    mn, mx = value_range
    return tensor.permute(0, 2, 3, 1).clamp(mn, mx).sub(mn).div(mx - mn).mul(255).round().to(torch.uint8).contiguous()


class VideoWriter:
    def __init__(self, output_file, output_size, fps, audio_file=None, audio_offset=0, audio_duration=None, ffmpeg_preset="slow",
                 debug=False, value_range=(0, 1), sink=None, ring_depth=4):
        self.output_file, self.fps, self.value_range, self.debug = output_file, fps, value_range, debug
        self.size = (2 * ceil(output_size[0] / 2), 2 * ceil(output_size[1] / 2))      # (w, h), even like the reference (:35)
        self.audio = (audio_file, audio_offset, audio_duration)
        self.ffmpeg_preset, self.ring_depth = ffmpeg_preset, ring_depth
        self._sink, self._own_sink, self._proc, self._writer, self.frames_written = sink, sink is None, None, None, 0

    # ---- sink -------------------------------------------------------------------------------------------------------------
    def _open(self):
        if self._sink is not None:
            return
        exe = shutil.which("ffmpeg")
        if exe is None:
            self._sink = open(self.output_file + ".rgb24", "wb")
            return
        audio_file, offset, duration = self.audio
        cmd = [exe, "-hide_banner", "-y", "-v", "warning", "-f", "rawvideo", "-pix_fmt", "rgb24", "-framerate", str(self.fps),
               "-s", f"{self.size[0]}x{self.size[1]}", "-i", "pipe:"]
        if audio_file is not None:
            cmd += ["-ss", str(offset), "-guess_layout_max", "0"] + (["-t", str(duration)] if duration is not None else []) + ["-i", audio_file]
        cmd += ["-framerate", str(self.fps), "-pix_fmt", "yuv420p", "-preset", self.ffmpeg_preset]
        if audio_file is not None:
            cmd += ["-b:a", "320K", "-ac", "2"]
        self._proc = subprocess.Popen(cmd + [self.output_file], stdin=subprocess.PIPE, stderr=None if self.debug else subprocess.DEVNULL)
        self._sink = self._proc.stdin

    def __enter__(self):
        self._open()
        return self

    def write(self, tensor):
        if tensor.dim() == 3:
            tensor = tensor[None]
        b, _, h, w = tensor.shape
        if h % 2 or w % 2:
            if not tensor.is_cuda:
                raise RuntimeError("VideoWriter: odd frame sizes are resampled on the device only (pass CUDA frames)")
            from ... import ops

            tensor = ops.resample(tensor.float(), (2 * ceil(h / 2), 2 * ceil(w / 2)))
        if tensor.is_cuda:
            from ._loop import frames_to_rgb24

            u8 = frames_to_rgb24(tensor, value_range=self.value_range)      # one kernel
        else:
            u8 = tensor2bytes_device(tensor, self.value_range)
        if self._writer is None or tuple(self._writer.ring[0].shape[1:]) != tuple(u8.shape[1:]) or self._writer.ring[0].shape[0] < b:
            if self._writer is not None:
                self._writer.close()
            make = (lambda: torch.empty(u8.shape, dtype=torch.uint8).pin_memory()) if u8.is_cuda else (lambda: np.empty(tuple(u8.shape), dtype=np.uint8))
            self._open()
            self._writer = RingWriter(self._sink, [make() for _ in range(self.ring_depth)])
        k = self._writer.acquire()
        slot = self._writer.ring[k]
        if u8.is_cuda:
            slot[:b].copy_(u8, non_blocking=True)
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream())
        else:
            slot[:b] = u8.numpy()
            event = None
        self._writer.submit(k, b, event)
        self.frames_written += b

    def __exit__(self, exc_type, exc, tb):
        try:
            if self._writer is not None:
                self._writer.close()
        finally:
            if self._own_sink and self._sink is not None:
                self._sink.close()
            if self._proc is not None:
                self._proc.wait()
        return False


def write_video(tensor, output_file, fps=24, audio_file=None, audio_offset=0, audio_duration=None, ffmpeg_preset="slow",
                debug=False, value_range=(0, 1), sink=None):
    """ops/video.py:131-155: write a [T, C, H, W] tensor (or numpy array) frame by frame."""
    _, _, h, w = tensor[[0]].shape
    with VideoWriter(output_file, (w, h), fps, audio_file, audio_offset, audio_duration, ffmpeg_preset, debug, value_range, sink=sink) as video:
        for frame in tensor:
            frame = frame if isinstance(frame, torch.Tensor) else torch.from_numpy(frame.copy())
            video.write(frame.squeeze().unsqueeze(0))
