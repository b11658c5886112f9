"""Shared frame loop of the renderers: pinned host staging -> async H2D -> synthesizer -> uint8 frames."""
from __future__ import annotations

import torch


def frame_batches(synthesizer, inputs, batch_size, device):
    """Yield (start, frames_f32[B,3,H,W] in ~[-1,1]) per batch; inputs is the reference's dict of [T,...] tensors."""
    keys = list(inputs.keys())
    T = len(next(iter(inputs.values())))
    staged = {k: (v.detach() if v.is_cuda else v.detach().cpu().contiguous().pin_memory()) for k, v in inputs.items()}
    synthesizer = synthesizer.to(device)
    for i in range(0, T, batch_size):
        batch = {k: staged[k][i:i + batch_size].to(device, non_blocking=True) for k in keys}
        yield i, synthesizer(**batch)


def to_uint8(frames01):
    """clamp -> *255 -> round -> uint8 (tensor2bytes, maua/ops/io.py:47-70)."""
    return frames01.clamp(0, 1).mul(255).round().to(torch.uint8)


class AsyncFrameDownloader:
    """Finished uint8 frame batches -> pinned host ring on a side stream, so the device->host copy of batch i
    overlaps the synthesis of batch i+1 (the reference copies each batch synchronously, render/memmap.py:31-33).

        buf = dl.device_buffer(i)          # waits (on the compute stream) until copy i-depth has drained
        net(ws, out_fmt="u8", out=buf)
        dl.download(i)                     # enqueue the copy behind the kernels of batch i
        frames = dl.host(i)                # blocks the host until batch i is in pinned memory
    """

    def __init__(self, batch_shape, device, depth=2):
        self.depth = depth
        self.dev = [torch.empty(batch_shape, dtype=torch.uint8, device=device) for _ in range(depth)]
        self.pinned = [torch.empty(batch_shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]   # kernels of the batch are done
        self.done = [None] * depth                                 # copy of the batch is done

    def device_buffer(self, i):
        k = i % self.depth
        if self.done[k] is not None:
            torch.cuda.current_stream().wait_event(self.done[k])
        return self.dev[k]

    def download(self, i):
        k = i % self.depth
        self.ready[k].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ready[k])
            self.pinned[k].copy_(self.dev[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.done[k] = ev

    def host(self, i):
        k = i % self.depth
        self.done[k].synchronize()
        return self.pinned[k]

    def synchronize(self):
        self.copy_stream.synchronize()
