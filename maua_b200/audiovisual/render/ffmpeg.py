"""FFMPEG renderer: mirror of maua/audiovisual/render/ffmpeg.py:21-77.

Same constructor / call signature.  Frames are converted to the rgb24 wire format on the device and streamed
to ``ffmpeg`` on stdin when the binary exists; without it (this image has none) the raw rgb24 stream is written
to ``output_file + ".rgb24"`` so the sink stays testable.  The device->host copies go into a ring of pinned buffers and
a writer thread feeds the pipe (render/_sink.py), so copy and pipe write overlap the synthesis of the next batches
(the x264 encode itself stays with ffmpeg).
"""
import shutil
import subprocess

import torch

from . import Renderer
from ._loop import frame_batches, to_uint8
from ._sink import RingWriter


class FFMPEG(Renderer):
    def __init__(self, output_file, fps=24, audio_file=None, audio_offset=0, audio_duration=None, ffmpeg_preset="medium",
                 batch_size=16):
        super().__init__()
        self.output_file, self.fps, self.ffmpeg_preset = output_file, fps, ffmpeg_preset
        self.audio_file, self.audio_offset, self.audio_duration = audio_file, audio_offset, audio_duration
        self.batch_size = batch_size

    def _open_sink(self, w, h):
        exe = shutil.which("ffmpeg")
        if exe is None:
            return open(self.output_file + ".rgb24", "wb"), None
        cmd = [exe, "-y", "-f", "rawvideo", "-pix_fmt", "rgb24", "-s", f"{w}x{h}", "-r", str(self.fps), "-i", "-"]
        if self.audio_file is not None:
            cmd += ["-ss", str(self.audio_offset)] + (["-t", str(self.audio_duration)] if self.audio_duration else []) + ["-i", self.audio_file]
        cmd += ["-c:v", "libx264", "-preset", self.ffmpeg_preset, "-pix_fmt", "yuv420p", self.output_file]
        proc = subprocess.Popen(cmd, stdin=subprocess.PIPE, stderr=subprocess.DEVNULL)
        return proc.stdin, proc

    ring_depth = 3

    def __call__(self, synthesizer, inputs, postprocess, fp16=True):
        sink, proc, writer = None, None, None
        try:
            for start, frames in frame_batches(synthesizer, inputs, self.batch_size, self.device):
                frame_batch = postprocess(frames.add(1).div(2))
                u8 = to_uint8(frame_batch).permute(0, 2, 3, 1).contiguous()  # rgb24: H, W, 3 per frame
                if writer is None:
                    sink, proc = self._open_sink(u8.shape[2], u8.shape[1])
                    ring = [torch.empty((self.batch_size,) + tuple(u8.shape[1:]), dtype=torch.uint8).pin_memory()
                            for _ in range(self.ring_depth)]
                    writer = RingWriter(sink, ring)
                k = writer.acquire()
                writer.ring[k][: u8.shape[0]].copy_(u8, non_blocking=True)
                event = torch.cuda.Event()
                event.record(torch.cuda.current_stream())
                writer.submit(k, u8.shape[0], event)
        finally:
            try:
                if writer is not None:
                    writer.close()
            finally:
                if sink is not None:
                    sink.close()
                if proc is not None:
                    proc.wait()
